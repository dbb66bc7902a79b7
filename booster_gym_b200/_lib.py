"""ctypes binding of libb200t1.so - the ONLY compute backend of this package.

There is no CPU fallback: if the shared library is missing it is built in-tree with nvcc (see _build.py); if that is
impossible, or the library does not export the C-ABI of include/b200_t1.h, importing this module raises.
"""
import ctypes as C
import os

from . import _abi, _build

_LIB = None

_vp, _i, _i64, _u64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double
_ip = C.POINTER(C.c_int)
_cpp = C.POINTER(C.c_char_p)

# name -> (restype, argtypes); pointers to device memory travel as void* (integers from tensor.data_ptr())
PROTOTYPES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (_i, []),
    "b200_sizeof": (_i, [_i]),
    "b200_t1_num_float_rows": (_i, []),
    "b200_t1_num_int_rows": (_i, []),
    "b200_t1_num_fields": (_i, [_i]),
    "b200_t1_field_info": (_i, [_i, _i, _cpp, _ip, _ip]),
    "b200_t1_create": (_i, [C.POINTER(_abi.ModelF), C.POINTER(_abi.T1Config), _vp, _i, _i, _i, _i, _u64, C.POINTER(_vp)]),
    "b200_t1_destroy": (_i, [_vp]),
    "b200_t1_bind_state": (_i, [_vp, _vp, _vp]),
    "b200_t1_num_envs": (_i, [_vp]),
    "b200_t1_init_params": (_i, [_vp, _i, _i, _vp]),
    "b200_t1_reset": (_i, [_vp, _vp, _vp, _vp]),
    "b200_t1_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b200_t1_physics": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b200_t1_post_physics": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "b200_t1_episode_stats": (_i, [_vp, C.POINTER(_d), C.POINTER(_i64), _vp]),
    "b200_t1_episode_stats_async": (_i, [_vp, _vp, _vp]),
    "b200_terrain_heights": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b200_rng_fill": (_i, [_vp, _u64, _i, _i, _i, _vp, _vp]),
    "b200_t1_bind_curriculum": (_i, [_vp, _vp, _i, _i]),
    "b200_t1_inject_rng": (_i, [_vp, _vp]),
    "b200_t1_rng_slots": (_i, []),
    "b200_t1_counters": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "b200_ppo_num_params": (_i, []),
    "b200_ppo_num_param_tensors": (_i, []),
    "b200_ppo_param_info": (_i, [_i, _cpp, _ip, _ip, _ip]),
    "b200_ppo_workspace_bytes": (_i64, [_i, _i]),
    "b200_ppo_create": (_i, [C.POINTER(_abi.PpoConfig), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_vp)]),
    "b200_ppo_destroy": (_i, [_vp]),
    "b200_policy_act": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _u64, _u64, _i, _vp]),
    "b200_critic_value": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "b200_ppo_old_dist": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_gae": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _i, _i, _vp]),
    "b200_ppo_epoch": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_ppo_epoch_a": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_ppo_epoch_b": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200_ppo_buffer": (_vp, [_vp, _i]),
    "b200_ppo_apply": (_i, [_vp, _vp]),
    "b200_ppo_peer_buffer_bytes": (C.c_longlong, []),
    "b200_ppo_bind_peers": (_i, [_vp, C.POINTER(C.c_ulonglong), _i, _i]),
    "b200_launch_count": (C.c_longlong, []),
    "b200_profile_gemm": (_i, [_i]),
    "b200_profile_gemm_read": (_i, [_i, C.POINTER(_d), C.POINTER(_d), _ip]),
    "b200_profile_gemm_bytes": (_i, [_i, C.POINTER(_d)]),
    "b200_tc_set_pair": (_i, [_i]),
    "b200_fma_peak": (_i, [C.POINTER(_d), _vp]),
    "b200_tc_set_chain": (_i, [_i]),
    "b200_tc_set_h2": (_i, [_i]),
}


class B200Error(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen libb200t1.so, declare every prototype and verify the struct layouts. Raises if anything is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not _build.up_to_date():
        # missing, or built from other sources than the tree holds (a stale library would be loaded silently otherwise): rebuild
        # under a file lock - concurrent ranks wait for the first one.  Without nvcc a stale library is an error, never a fallback.
        if not build_if_missing:
            raise B200Error(f"{path} is missing or stale (no CPU fallback exists; run `python -m booster_gym_b200._build`)")
        try:
            _build.build_locked()
        except RuntimeError as e:
            raise B200Error(f"{path} is missing or stale and cannot be rebuilt here: {e}") from e
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise B200Error(f"{path} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    for what, struct in ((0, _abi.ModelF), (1, _abi.T1Config), (2, _abi.PpoConfig), (3, _abi.ModelD)):
        if lib.b200_sizeof(what) != C.sizeof(struct):
            raise B200Error(f"ABI mismatch for struct #{what}: C {lib.b200_sizeof(what)} vs ctypes {C.sizeof(struct)}")
    _LIB = lib
    return lib


def check(rc, what=""):
    """C return code -> the Python exception type the reference raises for the same condition."""
    if rc == _abi.OK:
        return
    msg = load().b200_last_error().decode("utf-8", "replace")
    if rc == _abi.ERR_ARG:
        raise ValueError(f"{what}: {msg}")
    if rc == _abi.ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise B200Error(f"{what}: {msg} (code {rc})")


def field_table(kind):
    """{name: (row, count)} of the structure-of-arrays state (kind 0 float rows, 1 int rows)."""
    lib = load()
    out = {}
    for i in range(lib.b200_t1_num_fields(kind)):
        name = C.c_char_p()
        row = C.c_int()
        cnt = C.c_int()
        check(lib.b200_t1_field_info(kind, i, C.byref(name), C.byref(row), C.byref(cnt)), "field_info")
        out[name.value.decode()] = (row.value, cnt.value)
    return out


def param_table():
    """[(name, offset, rows, cols)] of the flat learner parameter buffer."""
    lib = load()
    out = []
    for i in range(lib.b200_ppo_num_param_tensors()):
        name = C.c_char_p()
        off, r, c = C.c_int(), C.c_int(), C.c_int()
        check(lib.b200_ppo_param_info(i, C.byref(name), C.byref(off), C.byref(r), C.byref(c)), "param_info")
        out.append((name.value.decode(), off.value, r.value, c.value))
    return out
