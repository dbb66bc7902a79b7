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


# leg-leg contacts act between two links of a few kg, not on the whole robot: springs formed for this mass
SELF_CONTACT_MASS = 2.0
# Isaac Gym's default shape friction (the reference only randomises the foot shapes, envs/t1.py:162-167)
DEFAULT_SHAPE_FRICTION = 1.0


def fill_model(struct, js=None, foot_corner=None, enable_contact=True, enable_limits=True, dt=0.002, gravity=9.81,
               body_contact=True, self_contact=True, terrain_friction=1.0):
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
    # collision primitives other than the soles: trunk box, hip-yaw and shank cylinders (same both sides)
    g = js["geoms"]
    for r in range(3):
        struct.trunk_box_pos[r] = g["Trunk"][0]["pos"][r]
        struct.trunk_box_half[r] = g["Trunk"][0]["size"][r]
    struct.trunk_box_radius = sum(v * v for v in g["Trunk"][0]["size"]) ** 0.5
    for i, name in enumerate(("Hip_Yaw_Left", "Shank_Left")):
        assert g[name][0]["type"] == "cylinder" and g[name] == g[name.replace("Left", "Right")]
        for r in range(3):
            struct.cyl_pos[i][r] = g[name][0]["pos"][r]
        struct.cyl_radius[i] = g[name][0]["size"][0]
        struct.cyl_half[i] = g[name][0]["size"][1]
    struct.body_mu = 0.5 * (DEFAULT_SHAPE_FRICTION + terrain_friction)  # PhysX combines by averaging
    struct.self_k = SOLREF_K * SELF_CONTACT_MASS
    struct.self_c = SOLREF_B * SELF_CONTACT_MASS
    # the foot box (half 0.1115 x 0.05 x 0.015) as a capsule along its long axis for leg-leg contact
    fb = g["left_foot_link"][0]
    rad = fb["size"][1]
    for i, sgn in enumerate((1.0, -1.0)):
        struct.foot_cap[i][0] = fb["pos"][0] + sgn * (fb["size"][0] - rad)
        struct.foot_cap[i][1] = fb["pos"][1]
        struct.foot_cap[i][2] = fb["pos"][2]
    struct.foot_cap_radius = rad
    struct.enable_contact = 1 if enable_contact else 0
    # B200_BODY_CONTACT / B200_SELF_CONTACT = 0: measurement switches (cost of the f3 shapes), not configuration
    struct.enable_body_contact = 1 if (enable_contact and body_contact and os.environ.get("B200_BODY_CONTACT", "1") != "0") else 0
    struct.enable_self_contact = 1 if (enable_contact and self_contact and os.environ.get("B200_SELF_CONTACT", "1") != "0") else 0
    struct.enable_limits = 1 if enable_limits else 0
    struct.pad0 = 0
    return struct


def model_f(**kw):
    return fill_model(_abi.ModelF(), **kw)


def model_d(**kw):
    return fill_model(_abi.ModelD(), **kw)
