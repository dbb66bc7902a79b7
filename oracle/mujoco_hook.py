"""MuJoCo as the physics oracle WHEN IT IS IMPORTABLE ("if `import mujoco` ever succeeds on a box the harness must
prefer it automatically").  TEST INFRASTRUCTURE ONLY.

The reference's sim-to-sim dynamics are `mujoco.mj_step` on `resources/T1/T1_locomotion.xml` (play_mujoco.py:756, model load
:823-824).  Neither that file (reference source, not copied) nor the `mujoco` wheel (no network) is available in the build
container or on the GPU box, so:

* the model is RE-GENERATED as MJCF text from this repo's own constants (booster_gym_b200/assets/t1_model.json, extracted from the
  MJCF / URDF by tools/extract_model.py): same tree, masses, inertias (as `fullinertia`), hinge axes / ranges, collision primitives
  (trunk box, hip-yaw and shank cylinders, foot boxes), motors with the MJCF `ctrlrange` (hip-roll 45 N m and
  knee 65 N m there, 30 / 60 in the URDF the env path uses), timestep 0.002, MuJoCo's default solver / contact parameters, a ground
  plane;
* every entry point raises `Unavailable` when `import mujoco` fails, and the tests that use it skip - they are the tests that PIN
  row a11 the day a wheel is present (tests/test_gpu_mujoco.py).

Nothing in the product package, `bench.py`'s timed region or `smoke()` imports this module.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Unavailable(RuntimeError):
    pass


def available():
    try:
        import mujoco  # noqa: F401

        return True
    except Exception:
        return False


def _require():
    try:
        import mujoco

        return mujoco
    except Exception as e:  # noqa: BLE001
        raise Unavailable(f"mujoco is not importable here ({type(e).__name__}: {e}); parity of row a11 stays unpinned") from e


def model_json():
    return json.load(open(os.path.join(ROOT, "booster_gym_b200", "assets", "t1_model.json")))


def _fmt(v):
    return " ".join(repr(float(x)) for x in v)


def mjcf(model=None, contact=True, limits=True):
    """MJCF text of the T1 (13 bodies, free joint + 12 hinges) on a plane, from the repo's model constants."""
    m = model or model_json()
    names, parent, axis = m["body_names"], m["parent"], m["axis"]
    children = {i: [j for j in range(m["nbody"]) if parent[j] == i] for i in range(m["nbody"])}
    d = m["mujoco_defaults"]
    out = ['<mujoco model="T1_from_b200_constants">',
           f'  <option timestep="0.002" gravity="{_fmt(d["gravity"])}" integrator="{d["integrator"]}" cone="{d["cone"]}"/>',
           '  <compiler angle="radian" autolimits="true"/>',
           f'  <default><geom friction="{_fmt(d["geom_friction"])}" solref="{_fmt(d["solref"])}" solimp="{_fmt(d["solimp"])}" '
           f'contype="{1 if contact else 0}" conaffinity="{1 if contact else 0}"/></default>',
           '  <worldbody>',
           '    <geom name="floor" type="plane" size="0 0 0.05" pos="0 0 0"/>']

    def body(i, ind):
        pad = "  " * ind
        out.append(f'{pad}<body name="{names[i]}" pos="{_fmt(m["body_pos"][i])}">')
        ixx, iyy, izz, ixy, ixz, iyz = m["inertia"][i]
        out.append(f'{pad}  <inertial pos="{_fmt(m["ipos"][i])}" mass="{float(m["mass"][i])!r}" '
                   f'fullinertia="{_fmt([ixx, iyy, izz, ixy, ixz, iyz])}"/>')
        if i == 0:
            out.append(f'{pad}  <freejoint name="root"/>')
        else:
            ax = [0.0, 0.0, 0.0]
            ax[axis[i]] = 1.0
            rng = m["jnt_range"][i - 1]
            lim = f' range="{_fmt(rng)}"' if limits else ' limited="false"'
            out.append(f'{pad}  <joint name="{m["dof_names"][i - 1]}" type="hinge" axis="{_fmt(ax)}"{lim}/>')
        for g in m["geoms"].get(names[i], []):
            out.append(f'{pad}  <geom type="{g["type"]}" size="{_fmt(g["size"])}" pos="{_fmt(g["pos"])}"/>')
        for c in children[i]:
            body(c, ind + 1)
        out.append(f'{pad}</body>')

    body(0, 2)
    out.append('  </worldbody>')
    out.append('  <actuator>')
    for j, nm in enumerate(m["dof_names"]):
        out.append(f'    <motor name="{nm}" joint="{nm}" ctrlrange="{_fmt(m["ctrlrange"][j])}"/>')
    out.append('  </actuator>')
    out.append('</mujoco>')
    return "\n".join(out)


class Sim:
    """one MuJoCo instance of the generated model; state in the conventions of play_mujoco.py:726-730,854-860:
    qpos = [x y z, qw qx qy qz, q(12)], qvel = [v_world(3), omega_body(3), qd(12)]"""

    def __init__(self, contact=True, limits=True):
        mj = _require()
        self.mj = mj
        self.model = mj.MjModel.from_xml_string(mjcf(contact=contact, limits=limits))
        self.data = mj.MjData(self.model)
        assert self.model.nq == 19 and self.model.nv == 18 and self.model.nu == 12
        assert abs(float(self.model.body_mass.sum()) - model_json()["total_mass"]) < 1e-6

    def set_state(self, pos, quat_xyzw, vlin_world, w_body, q, qd):
        d = self.data
        x, y, z, w = quat_xyzw
        d.qpos[:] = np.concatenate([pos, [w, x, y, z], q])
        d.qvel[:] = np.concatenate([vlin_world, w_body, qd])

    def qacc(self, ctrl):
        """mj_forward: generalised acceleration for the current state under joint torques `ctrl` (clamped to ctrlrange by MuJoCo)"""
        self.data.ctrl[:] = ctrl
        self.mj.mj_forward(self.model, self.data)
        return np.array(self.data.qacc)

    def step(self, ctrl, n=1):
        self.data.ctrl[:] = ctrl
        for _ in range(n):
            self.mj.mj_step(self.model, self.data)
        d = self.data
        w, x, y, z = d.qpos[3:7]
        return dict(pos=np.array(d.qpos[0:3]), quat_xyzw=np.array([x, y, z, w]), q=np.array(d.qpos[7:19]), vlin=np.array(d.qvel[0:3]),
                    w_body=np.array(d.qvel[3:6]), qd=np.array(d.qvel[6:18]))

    def pd_rollout(self, targets_fn, kp, kd, policy_steps, decimation=10):
        """play_mujoco.py:751-755: tau = kp (target - q) - kd qd, clamped to ctrlrange, `decimation` mj_steps per policy step.
        targets_fn(state dict) -> 12 joint targets.  Returns the list of states after every policy step."""
        lim = np.array(model_json()["ctrlrange"])
        traj = []
        st = self.step(np.zeros(12), 0)
        for _ in range(policy_steps):
            tgt = targets_fn(st)
            for _ in range(decimation):
                q, qd = np.array(self.data.qpos[7:19]), np.array(self.data.qvel[6:18])
                tau = np.clip(kp * (tgt - q) - kd * qd, lim[:, 0], lim[:, 1])
                st = self.step(tau, 1)
            traj.append(st)
        return traj
