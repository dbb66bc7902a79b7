"""names only: nothing here simulates anything"""
SIM_PHYSX, SIM_FLEX = 0, 1
UP_AXIS_Y, UP_AXIS_Z = 0, 1
LOCAL_SPACE, ENV_SPACE, GLOBAL_SPACE = 0, 1, 2


class _Bag:
    def __init__(self, *a, **k):
        pass

    def __setattr__(self, k, v):
        object.__setattr__(self, k, v)


class Vec3(_Bag):
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


class Transform(_Bag):
    def __init__(self):
        self.p = Vec3()


class PlaneParams(_Bag):
    pass


class TriangleMeshParams(_Bag):
    def __init__(self):
        self.transform = Transform()


class AssetOptions(_Bag):
    pass


class SimParams(_Bag):
    def __init__(self):
        self.physx = _Bag()
        self.flex = _Bag()


def ContactCollection(x):
    return x


def acquire_gym():
    raise RuntimeError("the isaacgym stub cannot simulate; fill tensors by hand (tools/make_golden.py)")
