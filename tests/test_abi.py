"""The C-ABI library loads without a GPU and exports every symbol include/b200_t1.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b200_t1.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from booster_gym_b200 import _build, _lib

    _build.build()
    lib = C.CDLL(_build.LIB)
    names = declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    bound = set(_lib.PROTOTYPES)
    assert set(names) == bound, (sorted(set(names) - bound), sorted(bound - set(names)))


def test_struct_layouts_match_the_header():
    from booster_gym_b200 import _abi, _lib

    lib = _lib.load()
    assert lib.b200_sizeof(0) == C.sizeof(_abi.ModelF)
    assert lib.b200_sizeof(1) == C.sizeof(_abi.T1Config)
    assert lib.b200_sizeof(2) == C.sizeof(_abi.PpoConfig)
    assert lib.b200_sizeof(3) == C.sizeof(_abi.ModelD)
    assert lib.b200_sizeof(99) == -1
    assert lib.b200_version() >= 100
    assert lib.b200_ppo_num_params() == _abi.NPARAMS_PADDED
    assert lib.b200_t1_rng_slots() == 21
    ff = _lib.field_table(0)
    assert ff["root_states"] == (0, 13) and ff["env_origins"][1] == 3
    assert lib.b200_t1_num_float_rows() == max(r + c for r, c in ff.values())
    assert sum(r * c for _, _, r, c in _lib.param_table()) == _abi.NPARAMS == 177945


def test_error_codes_without_a_device():
    """argument validation happens before any CUDA call: bad arguments -> B200_ERR_ARG and the reference's exception types"""
    from booster_gym_b200 import _abi, _lib

    lib = _lib.load()
    assert lib.b200_t1_create(None, None, None, 0, 0, 0, 0, 0, None) == _abi.ERR_ARG
    assert b"bad argument" in lib.b200_last_error()
    assert lib.b200_ppo_workspace_bytes(0, 0) == 0
    assert lib.b200_ppo_workspace_bytes(24, 4096) > 500e6
    with pytest.raises(ValueError):
        _lib.check(lib.b200_ppo_param_info(99, None, None, None, None), "param_info")
    assert lib.b200_t1_step(None, None, None, None, None, None, None, None, 0, None) == _abi.ERR_ARG
    assert lib.b200_gae(None, None, None, None, None, 0.99, 0.95, None, None, None, 1, 1, None) == _abi.ERR_ARG


def test_no_cpu_fallback_in_the_product_path():
    """the package must fail loudly without its CUDA library / device: no oracle import, no CPU route"""
    import subprocess
    import sys

    src = ""
    for dp, _, files in os.walk(os.path.join(ROOT, "booster_gym_b200")):
        for f in files:
            if f.endswith(".py"):
                src += open(os.path.join(dp, f)).read()
    assert "import oracle" not in src and "from oracle" not in src
    code = ("import yaml, sys; sys.path.insert(0, %r); from booster_gym_b200.envs import T1;"
            "cfg = yaml.safe_load(open(%r)); cfg['basic']['sim_device'] = 'cpu'\n"
            "try:\n    T1(cfg)\nexcept Exception as e:\n    print(type(e).__name__)\n") % (ROOT, os.path.join(ROOT, "envs", "T1.yaml"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "B200Error" in out.stdout, out.stdout + out.stderr
