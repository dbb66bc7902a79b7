"""the generator names utils/terrain.py imports; bodies are the product's restatement (utils/terrain.py of this repo)"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _impl():
    # import by path: the reference's own `utils` package may be first on sys.path when this stub is in use
    import importlib.util

    name = "_b200_terrain_impl"
    if name in sys.modules:
        return sys.modules[name]
    import types

    pkg = types.ModuleType("_b200_pkg_stub")
    spec = importlib.util.spec_from_file_location(name, os.path.join(_ROOT, "booster_gym_b200", "utils", "terrain_gen.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def SubTerrain(*a, **k):
    return _impl().SubTerrain(*a, **k)


def random_uniform_terrain(*a, **k):
    return _impl().random_uniform_terrain(*a, **k)


def pyramid_sloped_terrain(*a, **k):
    return _impl().pyramid_sloped_terrain(*a, **k)


def discrete_obstacles_terrain(*a, **k):
    return _impl().discrete_obstacles_terrain(*a, **k)


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    import numpy as np

    return np.zeros((0, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.uint32)
