"""T1 rigid-body model constants -> the ModelF/ModelD structs of the C-ABI.

Numbers come from `assets/t1_model.json`, derived once from the reference's robot description
(resources/T1/T1_locomotion.xml:37-135, T1_locomotion.urdf limits) by tools/extract_model.py; the product never
parses XML and never needs the reference tree.
"""
import json
import os

from . import _abi

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "t1_model.json")

# MuJoCo default soft-constraint reference (solref = (0.02, 1), solimp d_max = 0.95):
#   k = 1 / (d_max^2 tc^2 zeta^2) [1/s^2],  b = 2 / (d_max tc) [1/s]
SOLREF_K = 1.0 / (0.95 ** 2 * 0.02 ** 2 * 1.0 ** 2)
SOLREF_B = 2.0 / (0.95 * 0.02)
# mass each sole corner is assumed to carry when the spring constants are formed (total mass / 4 corners of a stance foot)
CONTACT_MASS_FRACTION = 0.25
STICTION_VEL = 0.01


def load_json(path=_ASSET):
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def fill_model(struct, js=None, foot_corner=None, enable_contact=True, enable_limits=True, dt=0.002, gravity=9.81):
    js = js or load_json()
    for b in range(_abi.NB):
        for r in range(3):
            struct.body_pos[b][r] = js["body_pos"][b][r]
            struct.ipos[b][r] = js["ipos"][b][r]
        for r in range(6):
            struct.inertia[b][r] = js["inertia"][b][r]
        struct.mass[b] = js["mass"][b]
        struct.axis[b] = js["axis"][b]
    # the trunk offset in the MJCF is its initial world position, not a parent offset
    for r in range(3):
        struct.body_pos[0][r] = 0.0
    for j in range(_abi.NU):
        struct.jnt_lower[j] = js["jnt_range"][j][0]
        struct.jnt_upper[j] = js["jnt_range"][j][1]
        struct.dof_inertia[j] = 1.0 / js["dof_invweight0"][6 + j]
    fc = foot_corner or [[0.1215, 0.05, -0.03], [0.1215, -0.05, -0.03], [-0.1015, 0.05, -0.03], [-0.1015, -0.05, -0.03]]
    for c in range(4):
        for r in range(3):
            struct.foot_corner[c][r] = fc[c][r]
    struct.gravity = gravity
    struct.dt = dt
    cm = js["total_mass"] * CONTACT_MASS_FRACTION
    struct.contact_k = SOLREF_K * cm
    struct.contact_c = SOLREF_B * cm
    struct.stiction_vel = STICTION_VEL
    struct.limit_k = SOLREF_K
    struct.limit_c = SOLREF_B
    struct.enable_contact = 1 if enable_contact else 0
    struct.enable_limits = 1 if enable_limits else 0
    struct.pad0 = 0
    return struct


def model_f(**kw):
    return fill_model(_abi.ModelF(), **kw)


def model_d(**kw):
    return fill_model(_abi.ModelD(), **kw)
