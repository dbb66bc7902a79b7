#!/usr/bin/env python3
"""Extract the T1 rigid-body model constants from the reference's robot description files.

Reads (read-only) the reference's MJCF `resources/T1/T1_locomotion.xml:37-135` (kinematic tree, inertials,
collision boxes, actuator ctrlrange) and URDF `resources/T1/T1_locomotion.urdf` (revolute-joint position /
velocity / effort limits, `envs/t1.py:59-67` reads exactly these four numbers per DoF) and writes the numbers
the product needs as one JSON document: `booster_gym_b200/assets/t1_model.json`.

The JSON is *derived data* (numbers only) - the product never parses XML and never needs the reference tree.
Run once in the build container:  python tools/extract_model.py [/root/reference]

It also computes MuJoCo-style `body_invweight0` / `dof_invweight0` (inverse inertia seen at each body / DoF in
the MJCF reference pose qpos0) with a small NumPy fp64 rigid-body implementation that is independent of both the
C oracle and the CUDA kernel; tests cross-check all three.
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np


def quat_to_mat(q):  # wxyz
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def fl(s):
    return [float(v) for v in s.split()]


def parse_mjcf(path):
    root = ET.parse(path).getroot()
    bodies = []

    def visit(elem, parent):
        idx = len(bodies)
        inertial = elem.find("inertial")
        joint = elem.find("joint")
        q = np.array(fl(inertial.get("quat", "1 0 0 0")))
        q = q / np.linalg.norm(q)
        Rq = quat_to_mat(q)
        diag = np.array(fl(inertial.get("diaginertia")))
        I = Rq @ np.diag(diag) @ Rq.T
        b = {
            "name": elem.get("name"),
            "parent": parent,
            "pos": fl(elem.get("pos", "0 0 0")),
            "ipos": fl(inertial.get("pos")),
            "iquat": q.tolist(),
            "diaginertia": diag.tolist(),
            "mass": float(inertial.get("mass")),
            # body-frame inertia about the body's CoM: Ixx Iyy Izz Ixy Ixz Iyz
            "inertia": [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]],
            "joint_type": joint.get("type", "hinge"),
            "joint_name": joint.get("name", "root"),
        }
        if b["joint_type"] == "hinge":
            ax = fl(joint.get("axis"))
            assert sorted(np.abs(ax)) == [0, 0, 1] and sum(ax) == 1, "coordinate-axis hinges only"
            b["axis"] = int(np.argmax(ax))
            b["range"] = fl(joint.get("range"))
            assert fl(joint.get("pos", "0 0 0")) == [0, 0, 0]
        else:
            b["axis"] = -1
            b["range"] = [0.0, 0.0]
        assert elem.get("quat") is None, "body frames carry no rotation in this model"
        geoms = []
        for g in elem.findall("geom"):
            if g.get("contype") == "0":
                continue  # visual mesh
            geoms.append({"type": g.get("type"), "size": fl(g.get("size")), "pos": fl(g.get("pos", "0 0 0"))})
        b["geoms"] = geoms
        bodies.append(b)
        for child in elem.findall("body"):
            visit(child, idx)

    visit(root.find("worldbody").find("body"), -1)
    ctrl = {m.get("joint"): fl(m.get("ctrlrange")) for m in root.find("actuator").findall("motor")}
    act_order = [m.get("joint") for m in root.find("actuator").findall("motor")]
    return bodies, ctrl, act_order


def parse_urdf_limits(path):
    root = ET.parse(path).getroot()
    lim = {}
    for j in root.findall("joint"):
        if j.get("type") != "revolute":
            continue
        l = j.find("limit")
        lim[j.get("name")] = {
            "lower": float(l.get("lower")),
            "upper": float(l.get("upper")),
            "velocity": float(l.get("velocity")),
            "effort": float(l.get("effort")),
        }
    return lim


# ---------------------------------------------------------------------------------------------------------
# tiny fp64 rigid-body code (world frame, Jacobian based: M = sum_b J_b^T diag(m, I_b) J_b) used only to derive
# invweight0 and as a third opinion on the mass matrix in tests.
def rot_axis(k, a):
    c, s = np.cos(a), np.sin(a)
    R = np.eye(3)
    i, j = [(1, 2), (2, 0), (0, 1)][k]
    R[i, i] = c
    R[i, j] = -s
    R[j, i] = s
    R[j, j] = c
    return R


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def mass_matrix_jacobian(bodies, qpos):
    """M(q) with MuJoCo velocity conventions (free joint: world-frame linear, body-frame angular)."""
    nb = len(bodies)
    nv = 6 + (nb - 1)
    xpos = np.zeros((nb, 3))
    xmat = np.zeros((nb, 3, 3))
    q = qpos[3:7] / np.linalg.norm(qpos[3:7])
    xpos[0] = qpos[0:3]
    xmat[0] = quat_to_mat(q)
    for b in range(1, nb):
        p = bodies[b]["parent"]
        xpos[b] = xpos[p] + xmat[p] @ np.array(bodies[b]["pos"])
        xmat[b] = xmat[p] @ rot_axis(bodies[b]["axis"], qpos[7 + b - 1])
    M = np.zeros((nv, nv))
    Jc = []
    for b in range(nb):
        com = xpos[b] + xmat[b] @ np.array(bodies[b]["ipos"])
        Jp = np.zeros((3, nv))
        Jr = np.zeros((3, nv))
        Jp[:, 0:3] = np.eye(3)
        Jr[:, 3:6] = xmat[0]
        Jp[:, 3:6] = -skew(com - xpos[0]) @ xmat[0]
        a = b
        while a > 0:
            axis = xmat[a][:, bodies[a]["axis"]]
            Jr[:, 6 + a - 1] = axis
            Jp[:, 6 + a - 1] = np.cross(axis, com - xpos[a])
            a = bodies[a]["parent"]
        i = bodies[b]["inertia"]
        Ib = np.array([[i[0], i[3], i[4]], [i[3], i[1], i[5]], [i[4], i[5], i[2]]])
        Iw = xmat[b] @ Ib @ xmat[b].T
        M += bodies[b]["mass"] * Jp.T @ Jp + Jr.T @ Iw @ Jr
        Jc.append((Jp, Jr))
    return M, Jc


def invweight0(bodies):
    nb = len(bodies)
    qpos0 = np.zeros(7 + nb - 1)
    qpos0[0:3] = bodies[0]["pos"]
    qpos0[3] = 1.0
    M, Jc = mass_matrix_jacobian(bodies, qpos0)
    Minv = np.linalg.inv(M)
    body_iw = []
    for b in range(nb):
        Jp, Jr = Jc[b]
        At = Jp @ Minv @ Jp.T
        Ar = Jr @ Minv @ Jr.T
        body_iw.append([float(np.trace(At) / 3), float(np.trace(Ar) / 3)])
    d = np.diag(Minv)
    dof_iw = [float(np.mean(d[0:3]))] * 3 + [float(np.mean(d[3:6]))] * 3 + [float(v) for v in d[6:]]
    return body_iw, dof_iw


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("B200_REF_DIR", "/root/reference")
    bodies, ctrl, act_order = parse_mjcf(os.path.join(ref, "resources/T1/T1_locomotion.xml"))
    urdf = parse_urdf_limits(os.path.join(ref, "resources/T1/T1_locomotion.urdf"))
    dof_names = [b["joint_name"] for b in bodies[1:]]
    assert dof_names == act_order, "MJCF actuator order must equal the depth-first joint order (play_mujoco.py:742-743)"
    assert list(urdf.keys()) == dof_names, "URDF revolute order must equal MJCF joint order (SURVEY 5.1)"
    body_iw, dof_iw = invweight0(bodies)
    feet = [i for i, b in enumerate(bodies) if "foot" in b["name"]]
    out = {
        "source": "derived from booster_gym resources/T1/T1_locomotion.xml:37-135 and T1_locomotion.urdf (limits)",
        "nbody": len(bodies),
        "nv": 6 + len(bodies) - 1,
        "body_names": [b["name"] for b in bodies],
        "dof_names": dof_names,
        "parent": [b["parent"] for b in bodies],
        "axis": [b["axis"] for b in bodies],
        "body_pos": [b["pos"] for b in bodies],
        "ipos": [b["ipos"] for b in bodies],
        "iquat": [b["iquat"] for b in bodies],
        "diaginertia": [b["diaginertia"] for b in bodies],
        "inertia": [b["inertia"] for b in bodies],
        "mass": [b["mass"] for b in bodies],
        "total_mass": float(sum(b["mass"] for b in bodies)),
        "jnt_range": [b["range"] for b in bodies[1:]],
        "ctrlrange": [ctrl[n] for n in dof_names],
        "urdf_lower": [urdf[n]["lower"] for n in dof_names],
        "urdf_upper": [urdf[n]["upper"] for n in dof_names],
        "urdf_velocity": [urdf[n]["velocity"] for n in dof_names],
        "urdf_effort": [urdf[n]["effort"] for n in dof_names],
        "feet_bodies": feet,
        "geoms": {b["name"]: b["geoms"] for b in bodies if b["geoms"]},
        "body_invweight0": body_iw,
        "dof_invweight0": dof_iw,
        "mujoco_defaults": {
            "solref": [0.02, 1.0],
            "solimp": [0.9, 0.95, 0.001, 0.5, 2.0],
            "geom_friction": [1.0, 0.005, 0.0001],
            "gravity": [0.0, 0.0, -9.81],
            "cone": "pyramidal",
            "integrator": "Euler",
        },
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "booster_gym_b200", "assets", "t1_model.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.normpath(dst), "total mass", out["total_mass"])


if __name__ == "__main__":
    main()
