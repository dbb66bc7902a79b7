"""NumPy restatement of T1's post-physics step, reset and helpers.  TEST INFRASTRUCTURE ONLY.

Follows the reference line by line in float32, one NumPy operation per torch operation (so sums, products and
comparisons round identically; only libm transcendentals may differ from torch by an ulp):
  step (post-physics half)   envs/t1.py:460-497         reset / _reset_idx          envs/t1.py:294-341
  _refresh_feet_state        envs/t1.py:529-549         _resample_commands          envs/t1.py:362-389
  _kick / _push              envs/t1.py:499-527         _check_termination          envs/t1.py:551-558
  _compute_reward + terms    envs/t1.py:560-572,606-730 _compute_observations       envs/t1.py:574-603
  _teleport_robot            envs/t1.py:343-360         Terrain.terrain_heights     utils/terrain.py:101-121
  Isaac Gym torch_utils      restated
PINNED: tests/test_oracle_pinning.py checks it against tests/golden/env_*.npz and terrain_lookup.npz, which were produced
by the reference's own classes (tools/make_golden.py).  Random draws come from the injection table (slot layout of
include/b200_t1.h: b200_t1_inject_rng) - the same table the CUDA kernels read in the parity tests.
"""
import numpy as np

f32 = np.float32
PI = f32(np.pi)
TWO_PI = f32(2 * np.pi)


# ---- torch_utils restatement -------------------------------------------------------------------------------------------
def quat_rotate(q, v, sign=1.0):
    qw = q[:, 3]
    qv = q[:, :3]
    a = v * (f32(2.0) * qw ** 2 - f32(1.0))[:, None]
    b = np.cross(qv, v) * qw[:, None] * f32(2.0)
    c = qv * np.sum(qv * v, axis=1, dtype=f32)[:, None] * f32(2.0)
    return (a + b + c) if sign > 0 else (a - b + c)


def quat_rotate_inverse(q, v):
    return quat_rotate(q, v, sign=-1.0)


def get_euler_xyz(q):
    qx, qy, qz, qw = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    sinr_cosp = f32(2.0) * (qw * qx + qy * qz)
    cosr_cosp = qw * qw - qx * qx - qy * qy + qz * qz
    roll = np.arctan2(sinr_cosp, cosr_cosp)
    sinp = f32(2.0) * (qw * qy - qz * qx)
    with np.errstate(invalid="ignore"):
        pitch = np.where(np.abs(sinp) >= 1, np.abs(f32(np.pi / 2.0)) * np.sign(sinp), np.arcsin(sinp)).astype(f32)
    siny_cosp = f32(2.0) * (qw * qz + qx * qy)
    cosy_cosp = qw * qw + qx * qx - qy * qy - qz * qz
    yaw = np.arctan2(siny_cosp, cosy_cosp)
    return np.mod(roll, TWO_PI), np.mod(pitch, TWO_PI), np.mod(yaw, TWO_PI)


def quat_from_yaw(yaw):
    cy, sy = np.cos(yaw * f32(0.5)), np.sin(yaw * f32(0.5))
    z = np.zeros_like(yaw)
    return np.stack([z, z, sy, cy], axis=-1)


def terrain_heights(pos, hf, border_pixels=50, hs=0.1, vs=0.005):
    """utils/terrain.py:101-121; hf None -> plane"""
    pos = np.asarray(pos, dtype=f32)
    if hf is None:
        return np.zeros(len(pos), dtype=f32)
    x = border_pixels + pos[:, 0] / hs   # float32 array / Python float -> float32 (NumPy 2 weak scalars)
    y = border_pixels + pos[:, 1] / hs
    x1 = np.floor(x).astype(int); x2 = x1 + 1
    y1 = np.floor(y).astype(int); y2 = y1 + 1
    return (((x2 - x) * (y2 - y) * hf[x1, y1] + (x - x1) * (y2 - y) * hf[x2, y1] + (x2 - x) * (y - y1) * hf[x1, y2]
             + (x - x1) * (y - y1) * hf[x2, y2]) * vs).astype(f32)


def wrap_pi(a):
    return np.mod(a + PI, TWO_PI) - PI


# ---- injection table accessors -------------------------------------------------------------------------------------------
def _u(t, slot, lane):
    return t[slot, 4 + lane].view(f32)


def _n(t, slot, lane):
    return t[slot, 8 + lane].view(f32)


def _w(t, slot, lane):
    return t[slot, lane]


def _randomize(x, params, u=None, nrm=None):
    """utils/utils.py:5-30 with the raw sample supplied"""
    if params is None:
        return x
    if params["distribution"] == "gaussian":
        val = f32(params["range"][0]) + f32(params["range"][1]) * nrm
    else:
        lo, hi = params["range"]
        val = f32(lo) + f32(hi - lo) * u
    return (x + val if params["operation"] == "additive" else x * val).astype(f32)


class EnvOracle:
    """state: dict of float32 / int64 arrays in the reference's [N, ...] layouts (keys of ref_harness.STATE_KEYS_*)"""

    def __init__(self, cfg, state, hf=None, model_json=None):
        self.cfg = cfg
        self.s = {k: np.array(v, copy=True) for k, v in state.items()}
        self.hf = hf
        self.n = self.s["root_states"].shape[0]
        self.dt = cfg["control"]["decimation"] * cfg["sim"]["dt"]
        js = model_json
        self.dof_lower = np.asarray(js["urdf_lower"], f32)
        self.dof_upper = np.asarray(js["urdf_upper"], f32)
        self.torque_limits = np.asarray(js["urdf_effort"], f32)
        self.dof_vel_limits = np.asarray(js["urdf_velocity"], f32)
        self.default_dof_pos = np.array([-0.2 if "Hip_Pitch" in nm else 0.4 if "Knee_Pitch" in nm else -0.25 if "Ankle_Pitch" in nm
                                         else 0.0 for nm in js["dof_names"]], f32)
        dja = cfg["init_state"]["default_joint_angles"]
        for j, nm in enumerate(js["dof_names"]):
            val = dja["default"]
            for key in dja:
                if key != "default" and key in nm:
                    val = dja[key]
            self.default_dof_pos[j] = val
        tr = cfg["terrain"]
        self.trimesh = tr["type"] != "plane"
        self.border_size = tr.get("border_size", 0.0)
        self.env_width = tr.get("num_terrains", 0) * tr.get("terrain_width", 0.0)
        self.env_length = tr.get("terrain_length", 0.0)
        self.bp = int(self.border_size / tr["horizontal_scale"]) if self.trimesh else 0
        self.hs, self.vs = tr["horizontal_scale"], tr["vertical_scale"]
        self.scales = [(k, v * self.dt) for k, v in cfg["rewards"]["scales"].items() if v != 0]
        self.s.setdefault("base_lin_vel", np.zeros((self.n, 3), f32))
        self.s.setdefault("base_ang_vel", np.zeros((self.n, 3), f32))
        self.s.setdefault("projected_gravity", np.zeros((self.n, 3), f32))
        self.s.setdefault("feet_roll", np.zeros((self.n, 2), f32))
        self.s.setdefault("feet_yaw", np.zeros((self.n, 2), f32))
        self.s.setdefault("feet_contact", np.zeros((self.n, 2), bool))
        self.extras_time_outs = np.zeros(self.n, bool)
        names = js["body_names"]   # envs/t1.py:86-92: substring match, in the order of the YAML lists
        self.penalized = [names.index(b) for key in cfg["rewards"]["penalize_contacts_on"] for b in names if key in b]
        self.terminating = [names.index(b) for key in cfg["rewards"]["terminate_contacts_on"] for b in names if key in b]

    def contact_flags(self):
        """[n, 13] bool: norm(contact_forces[:, b]) > 1.  Input either as the raw force tensor (fixtures made by the reference) or
        as the kernel's per-lane bit masks (state rows `contact_mask`)."""
        s = self.s
        if "contact_forces" in s:
            cf = np.asarray(s["contact_forces"], f32).reshape(self.n, 13, 3)
            return np.sqrt(np.sum(cf * cf, axis=2, dtype=f32)) > f32(1.0)
        if "contact_mask" in s:
            mk = np.bitwise_or.reduce(np.asarray(s["contact_mask"]).reshape(self.n, -1).astype(np.int64), axis=1)
            return ((mk[:, None] >> np.arange(13)[None, :]) & 1).astype(bool)
        return np.zeros((self.n, 13), bool)

    def init_derived(self):
        """envs/t1.py:240-244: base-frame vectors from the state the env is constructed with"""
        s, n = self.s, self.n
        rs = s["root_states"]
        g = np.tile(np.array([0.0, 0.0, -1.0], f32), (n, 1))
        s["base_lin_vel"] = quat_rotate_inverse(rs[:, 3:7], rs[:, 7:10])
        s["base_ang_vel"] = quat_rotate_inverse(rs[:, 3:7], rs[:, 10:13])
        s["projected_gravity"] = quat_rotate_inverse(rs[:, 3:7], g)

    def h(self, pos):
        return terrain_heights(pos, self.hf if self.trimesh else None, self.bp, self.hs, self.vs)

    # ---- envs/t1.py:529-549
    def refresh_feet(self):
        s, n = self.s, self.n
        fq = s["feet_quat"].reshape(n * 2, 4)
        fp = s["feet_pos"].reshape(n * 2, 3)
        roll, _, yaw = get_euler_xyz(fq)
        s["feet_roll"] = wrap_pi(roll.reshape(n, 2))
        s["feet_yaw"] = wrap_pi(yaw.reshape(n, 2))
        edges = np.asarray(self.cfg["asset"]["feet_edge_pos"], f32)
        k = edges.shape[0]
        ep = np.repeat(fp, k, axis=0)
        eq = np.repeat(fq, k, axis=0)
        rel = np.tile(edges, (n * 2, 1))
        world = ep + quat_rotate(eq, rel)
        s["feet_contact"] = np.any((world[:, 2] - self.h(world) < f32(0.01)).reshape(n, 2, k), axis=2)

    # ---- envs/t1.py:391-413 (command curriculum): every successful episode raises its grid cell and the 4 neighbours
    def update_curriculum(self, ids):
        cm = self.cfg["commands"]
        if not cm.get("curriculum"):
            return
        s = self.s
        prob, lev = s["curriculum_prob"], s["env_curriculum_level"]
        success = s["episode_length_buf"][ids].astype(f32) > f32(np.ceil(self.cfg["rewards"]["episode_length_s"] / self.dt) * (1 - cm["episode_length_toler"]))
        success &= np.abs(s["filtered_lin_vel"][ids, 0] - s["commands"][ids, 0]) < f32(cm["lin_vel_x_toler"])
        success &= np.abs(s["filtered_lin_vel"][ids, 1] - s["commands"][ids, 1]) < f32(cm["lin_vel_y_toler"])
        success &= np.abs(s["filtered_ang_vel"][ids, 2] - s["commands"][ids, 2]) < f32(cm["ang_vel_yaw_toler"])
        rate = f32(cm["update_rate"])
        for i in np.nonzero(success)[0]:
            x = int(lev[ids[i], 0]) + cm["lin_vel_levels"]
            y = int(lev[ids[i], 1]) + cm["ang_vel_levels"]
            prob[x, y] = f32(prob[x, y] + rate)
            if x > 0:
                prob[x - 1, y] = f32(prob[x - 1, y] + rate)
            if x < prob.shape[0] - 1:
                prob[x + 1, y] = f32(prob[x + 1, y] + rate)
            if y > 0:
                prob[x, y - 1] = f32(prob[x, y - 1] + rate)
            if y < prob.shape[1] - 1:
                prob[x, y + 1] = f32(prob[x, y + 1] + rate)
        np.minimum(prob, f32(1.0), out=prob)

    # torch.multinomial(prob.flatten(), k, replacement=True) restated as the inverse CDF of the SAME distribution on an injected
    # uniform (torch's sampler cannot be fed samples): sequential fp32 running sum, first cell whose sum exceeds u * total
    @staticmethod
    def multinomial_icdf(prob_flat, u):
        cdf = np.add.accumulate(prob_flat.astype(f32), dtype=f32)
        target = (u.astype(f32) * cdf[-1]).astype(f32)
        return np.minimum(np.searchsorted(cdf, target, side="right"), len(cdf) - 1).astype(np.int64)

    # ---- envs/t1.py:415-435
    def resample_curriculum_commands(self, ids, table):
        s, cm = self.s, self.cfg["commands"]
        prob = s["curriculum_prob"]
        grid_idx = self.multinomial_icdf(prob.reshape(-1), _u(table, 7, 2)[ids])
        # reference quirk (SURVEY 8a note 13): the flat index of the [lin, ang] grid is decoded as (idx % cols, idx // cols),
        # i.e. the LINEAR level comes from the column (ang) coordinate and vice versa - a transposition of _update_curriculum's axes
        lin = grid_idx % prob.shape[1] - cm["lin_vel_levels"]
        ang = grid_idx // prob.shape[1] - cm["ang_vel_levels"]
        s["env_curriculum_level"][ids, 0] = lin
        s["env_curriculum_level"][ids, 1] = ang
        def rnd(lo, hi, lane):
            return (f32(hi - lo) * _u(table, 6, lane)[ids] + f32(lo)).astype(f32)
        s["commands"][ids, 0] = ((lin.astype(f32) + rnd(-0.5, 0.5, 0)).astype(f32) * f32(cm["lin_vel_x_resolution"])).astype(f32)
        s["commands"][ids, 1] = ((np.abs(lin).astype(f32) * rnd(-1.0, 1.0, 1)).astype(f32) * f32(cm["lin_vel_y_resolution"])).astype(f32)
        s["commands"][ids, 2] = ((ang.astype(f32) + rnd(-0.5, 0.5, 2)).astype(f32) * f32(cm["ang_vel_resolution"])).astype(f32)

    # ---- envs/t1.py:301-341 for the envs in ids
    def reset_idx(self, ids, table):
        if len(ids) == 0:
            return
        self.update_curriculum(ids)
        s, rz = self.s, self.cfg["randomization"]
        noise12 = np.stack([_n(table, i // 4, i % 4)[ids[0]] for i in range(12)]).astype(f32)  # one [1,12] draw for the whole call
        u12 = np.stack([_u(table, i // 4, i % 4)[ids[0]] for i in range(12)]).astype(f32)
        s["dof_pos"][ids] = _randomize(self.default_dof_pos[None, :], rz.get("init_dof_pos"), u12[None, :], noise12[None, :])
        s["dof_vel"][ids] = 0.0
        st = self.cfg["init_state"]
        base = np.asarray(st["pos"] + st["rot"] + st["lin_vel"] + st["ang_vel"], f32)
        rs = s["root_states"]
        rs[ids] = base
        rs[ids, :2] += s["env_origins"][ids, :2]
        uu = np.stack([_u(table, 3, 0)[ids], _u(table, 3, 1)[ids]], axis=1)
        nn = np.stack([_n(table, 3, 0)[ids], _n(table, 3, 1)[ids]], axis=1)
        rs[ids, :2] = _randomize(rs[ids, :2], rz.get("init_base_pos_xy"), uu, nn)
        rs[ids, 2] += self.h(rs[ids, :2])
        rs[ids, 3:7] = quat_from_yaw(_u(table, 3, 2)[ids] * f32(2 * np.pi))
        uu = np.stack([_u(table, 4, 0)[ids], _u(table, 4, 1)[ids]], axis=1)
        nn = np.stack([_n(table, 4, 0)[ids], _n(table, 4, 1)[ids]], axis=1)
        rs[ids, 7:9] = _randomize(np.zeros((len(ids), 2), f32), rz.get("init_base_lin_vel_xy"), uu, nn)
        s["last_dof_targets"][ids] = s["dof_pos"][ids]
        s["last_root_vel"][ids] = rs[ids, 7:13]
        s["episode_length_buf"][ids] = 0
        s["filtered_lin_vel"][ids] = 0.0
        s["filtered_ang_vel"][ids] = 0.0
        s["cmd_resample_time"][ids] = 0
        s["delay_steps"][ids] = (_w(table, 5, 0)[ids].astype(np.int64) % self.cfg["control"]["decimation"])
        self.extras_time_outs = self.time_out_buf

    # ---- envs/t1.py:343-360
    def teleport(self):
        if not self.trimesh:
            return
        s = self.s
        rs = s["root_states"]
        bs, ew, el = self.border_size, self.env_width, self.env_length
        xmin = rs[:, 0] < -0.75 * bs
        xmax = rs[:, 0] > ew + 0.75 * bs
        ymin = rs[:, 1] < -0.75 * bs
        ymax = rs[:, 1] > el + 0.75 * bs
        fp = s["feet_pos"].reshape(self.n, 2, 3)
        rs[xmin, 0] += f32(ew + bs); rs[xmax, 0] -= f32(ew + bs)
        rs[ymin, 1] += f32(el + bs); rs[ymax, 1] -= f32(el + bs)
        fp[xmin, :, 0] += f32(ew + bs); fp[xmax, :, 0] -= f32(ew + bs)
        fp[ymin, :, 1] += f32(el + bs); fp[ymax, :, 1] -= f32(el + bs)
        s["feet_pos"] = fp.reshape(self.n, 6)
        if xmin.any() or xmax.any() or ymin.any() or ymax.any():
            self.refresh_feet()

    # ---- envs/t1.py:362-389
    def resample_commands(self, table):
        s, cm = self.s, self.cfg["commands"]
        ids = np.nonzero(s["episode_length_buf"] == s["cmd_resample_time"])[0]
        if len(ids) == 0:
            return
        def rnd(lo, hi, lane):  # torch_rand_float: (upper - lower) * rand + lower
            return (f32(hi - lo) * _u(table, 6, lane)[ids] + f32(lo)).astype(f32)
        if cm.get("curriculum"):
            self.resample_curriculum_commands(ids, table)
        else:
            s["commands"][ids, 0] = rnd(*cm["lin_vel_x"], 0)
            s["commands"][ids, 1] = rnd(*cm["lin_vel_y"], 1)
            s["commands"][ids, 2] = rnd(*cm["ang_vel_yaw"], 2)
        s["gait_frequency"][ids] = rnd(*cm["gait_frequency"], 3)
        still = ids[_u(table, 7, 0)[ids] < f32(cm["still_proportion"])]
        s["commands"][still, :] = 0.0
        s["gait_frequency"][still] = 0.0
        lo, hi = int(cm["resampling_time_s"][0] / self.dt), int(cm["resampling_time_s"][1] / self.dt)
        s["cmd_resample_time"][ids] += lo + (_w(table, 7, 1)[ids].astype(np.int64) % (hi - lo))

    # ---- envs/t1.py:574-603
    def observations(self, table, noise=True):
        s, nz, cfg = self.s, self.cfg["normalization"], self.cfg
        nc = cfg.get("noise", {}) if noise else {}
        def nr(start, width):
            return np.stack([_n(table, 12 + (start + i) // 4, (start + i) % 4) for i in range(width)], axis=1)
        def ur(start, width):
            return np.stack([_u(table, 12 + (start + i) // 4, (start + i) % 4) for i in range(width)], axis=1)
        on = (s["gait_frequency"] > f32(1.0e-8)).astype(f32)
        ph = f32(2 * np.pi) * s["gait_process"]
        cs = np.array([nz["lin_vel"], nz["lin_vel"], nz["ang_vel"]], f32)
        obs = np.concatenate([
            _randomize(s["projected_gravity"], nc.get("gravity"), ur(0, 3), nr(0, 3)) * f32(nz["gravity"]),
            _randomize(s["base_ang_vel"], nc.get("ang_vel"), ur(3, 3), nr(3, 3)) * f32(nz["ang_vel"]),
            s["commands"][:, :3] * cs,
            (np.cos(ph) * on)[:, None],
            (np.sin(ph) * on)[:, None],
            _randomize(s["dof_pos"] - self.default_dof_pos[None, :], nc.get("dof_pos"), ur(6, 12), nr(6, 12)) * f32(nz["dof_pos"]),
            _randomize(s["dof_vel"], nc.get("dof_vel"), ur(18, 12), nr(18, 12)) * f32(nz["dof_vel"]),
            s["actions"],
        ], axis=1).astype(f32)
        height = s["root_states"][:, 2] - self.h(s["root_states"][:, 0:3])
        priv = np.concatenate([
            s["base_mass_scaled"],
            _randomize(s["base_lin_vel"], nc.get("lin_vel"), ur(30, 3), nr(30, 3)) * f32(nz["lin_vel"]),
            _randomize(height, nc.get("height"), ur(33, 1)[:, 0], nr(33, 1)[:, 0])[:, None],
            s["pushing_forces"] * f32(nz["push_force"]),
            s["pushing_torques"] * f32(nz["push_torque"]),
        ], axis=1).astype(f32)
        return obs, priv

    # ---- reward terms, envs/t1.py:606-730 (unscaled)
    def reward(self, name):
        s, rw, n = self.s, self.cfg["rewards"], self.n
        dt = f32(self.dt)
        rs = s["root_states"]
        sq = np.square
        if name == "survival":
            return np.ones(n, f32)
        if name == "tracking_lin_vel_x":
            return np.exp(-sq(s["commands"][:, 0] - s["filtered_lin_vel"][:, 0]) / f32(rw["tracking_sigma"]))
        if name == "tracking_lin_vel_y":
            return np.exp(-sq(s["commands"][:, 1] - s["filtered_lin_vel"][:, 1]) / f32(rw["tracking_sigma"]))
        if name == "tracking_ang_vel":
            return np.exp(-sq(s["commands"][:, 2] - s["filtered_ang_vel"][:, 2]) / f32(rw["tracking_sigma"]))
        if name == "base_height":
            return sq(rs[:, 2] - self.h(rs[:, 0:3]) - f32(rw["base_height_target"]))
        if name == "collision":  # envs/t1.py:627-629 on the per-body flags |contact_forces[b]| > 1
            return np.sum(self.contact_flags()[:, self.penalized], axis=1).astype(f32)
        if name == "lin_vel_z":
            return sq(s["filtered_lin_vel"][:, 2])
        if name == "ang_vel_xy":
            return np.sum(sq(s["base_ang_vel"][:, :2]), axis=1, dtype=f32)
        if name == "orientation":
            return np.sum(sq(s["projected_gravity"][:, :2]), axis=1, dtype=f32)
        if name == "torques":
            return self._rowsum(sq(s["torques"]))
        if name == "dof_vel":
            return self._rowsum(sq(s["dof_vel"]))
        if name == "dof_acc":
            return self._rowsum(sq((s["last_dof_vel"] - s["dof_vel"]) / dt))
        if name == "root_acc":
            return self._rowsum(sq((s["last_root_vel"] - rs[:, 7:13]) / dt))
        if name == "action_rate":
            return self._rowsum(sq(s["last_actions"] - s["actions"]))
        if name == "dof_pos_limits":
            half = f32(0.5 * (1 - rw["soft_dof_pos_limit"]))
            lower = self.dof_lower + half * (self.dof_upper - self.dof_lower)
            upper = self.dof_upper - half * (self.dof_upper - self.dof_lower)
            return self._rowsum(((s["dof_pos"] < lower) | (s["dof_pos"] > upper)).astype(f32))
        if name == "dof_vel_limits":
            return self._rowsum(np.clip(np.abs(s["dof_vel"]) - self.dof_vel_limits * f32(rw["soft_dof_vel_limit"]), 0.0, 1.0))
        if name == "torque_limits":
            return self._rowsum(np.maximum(np.abs(s["torques"]) - self.torque_limits * f32(rw["soft_torque_limit"]), 0.0))
        if name == "torque_tiredness":
            return self._rowsum(np.minimum(sq(s["torques"] / self.torque_limits), f32(1.0)))
        if name == "power":
            return self._rowsum(np.maximum(s["torques"] * s["dof_vel"], f32(0.0)))
        fp = s["feet_pos"].reshape(n, 2, 3)
        lfp = s["last_feet_pos"].reshape(n, 2, 3)
        if name == "feet_slip":
            v2 = np.sum(sq((lfp - fp) / dt), axis=2, dtype=f32)
            return np.sum(v2 * s["feet_contact"].astype(f32), axis=1, dtype=f32) * (s["episode_length_buf"] > 1).astype(f32)
        if name == "feet_vel_z":
            return np.sum(sq((lfp - fp) / dt)[:, :, 2], axis=1, dtype=f32)
        if name == "feet_roll":
            return np.sum(sq(s["feet_roll"]), axis=1, dtype=f32)
        if name == "feet_yaw_diff":
            return sq(wrap_pi(s["feet_yaw"][:, 1] - s["feet_yaw"][:, 0]))
        if name == "feet_yaw_mean":
            mean = (s["feet_yaw"][:, 0] + s["feet_yaw"][:, 1]) / f32(2.0) + PI * (np.abs(s["feet_yaw"][:, 1] - s["feet_yaw"][:, 0]) > PI)
            return sq(wrap_pi(get_euler_xyz(rs[:, 3:7])[2] - mean.astype(f32)))
        if name == "feet_distance":
            yaw = get_euler_xyz(rs[:, 3:7])[2]
            dist = np.abs(np.cos(yaw) * (fp[:, 1, 1] - fp[:, 0, 1]) - np.sin(yaw) * (fp[:, 1, 0] - fp[:, 0, 0]))
            return np.clip(f32(rw["feet_distance_ref"]) - dist, 0.0, f32(0.1))
        if name == "feet_swing":
            moving = s["gait_frequency"] > f32(1.0e-8)
            ls = (np.abs(s["gait_process"] - f32(0.25)) < f32(0.5 * rw["swing_period"])) & moving
            rsw = (np.abs(s["gait_process"] - f32(0.75)) < f32(0.5 * rw["swing_period"])) & moving
            return (ls & ~s["feet_contact"][:, 0]).astype(f32) + (rsw & ~s["feet_contact"][:, 1]).astype(f32)
        raise KeyError(name)

    @staticmethod
    def _rowsum(a):
        """torch.sum(dim=-1) over 12 or 6 contiguous fp32 values: sequential left-to-right in fp32"""
        a = a.astype(f32)
        out = np.zeros(a.shape[0], f32)
        for j in range(a.shape[1]):
            out = out + a[:, j]
        return out

    # ---- envs/t1.py:460-497 with physics = given state; actions already clipped, torques already averaged
    def step_post(self, table, common_step, noise=True):
        s, cfg, n = self.s, self.cfg, self.n
        rs = s["root_states"]
        q = rs[:, 3:7]
        g = np.tile(np.array([0.0, 0.0, -1.0], f32), (n, 1))
        s["base_lin_vel"] = quat_rotate_inverse(q, rs[:, 7:10])
        s["base_ang_vel"] = quat_rotate_inverse(q, rs[:, 10:13])
        s["projected_gravity"] = quat_rotate_inverse(q, g)
        w = cfg["normalization"]["filter_weight"]
        s["filtered_lin_vel"] = s["base_lin_vel"] * f32(w) + s["filtered_lin_vel"] * f32(1.0 - w)
        s["filtered_ang_vel"] = s["base_ang_vel"] * f32(w) + s["filtered_ang_vel"] * f32(1.0 - w)
        self.refresh_feet()
        s["episode_length_buf"] = s["episode_length_buf"] + 1
        s["gait_process"] = np.fmod(s["gait_process"] + f32(self.dt) * s["gait_frequency"], f32(1.0)).astype(f32)
        rz = cfg["randomization"]
        if common_step % int(np.ceil(rz["kick_interval_s"] / self.dt)) == 0:
            nn = np.stack([_n(table, 8, i) for i in range(3)], axis=1); uu = np.stack([_u(table, 8, i) for i in range(3)], axis=1)
            rs[:, 7:10] = _randomize(rs[:, 7:10], rz.get("kick_lin_vel"), uu, nn)
            nn = np.stack([_n(table, 9, i) for i in range(3)], axis=1); uu = np.stack([_u(table, 9, i) for i in range(3)], axis=1)
            rs[:, 10:13] = _randomize(rs[:, 10:13], rz.get("kick_ang_vel"), uu, nn)
        pi_ = int(np.ceil(rz["push_interval_s"] / self.dt))
        if common_step % pi_ == 0:
            z = np.zeros((n, 3), f32)
            nn = np.stack([_n(table, 10, i) for i in range(3)], axis=1); uu = np.stack([_u(table, 10, i) for i in range(3)], axis=1)
            s["pushing_forces"] = _randomize(z, rz.get("push_force"), uu, nn)
            nn = np.stack([_n(table, 11, i) for i in range(3)], axis=1); uu = np.stack([_u(table, 11, i) for i in range(3)], axis=1)
            s["pushing_torques"] = _randomize(z, rz.get("push_torque"), uu, nn)
        elif common_step % pi_ == int(np.ceil(rz["push_duration_s"] / self.dt)):
            s["pushing_forces"] = np.zeros((n, 3), f32)
            s["pushing_torques"] = np.zeros((n, 3), f32)
        # _check_termination
        rw = cfg["rewards"]
        vsq = self._rowsum(np.square(rs[:, 7:13]))
        reset = np.any(self.contact_flags()[:, self.terminating], axis=1)   # :553
        reset |= vsq > f32(rw["terminate_vel"])
        reset |= rs[:, 2] - self.h(rs[:, 0:3]) < f32(rw["terminate_height"])
        time_out = s["episode_length_buf"] > np.ceil(rw["episode_length_s"] / self.dt)
        reset |= time_out
        time_out = time_out | (s["episode_length_buf"] == s["cmd_resample_time"])
        self.reset_buf, self.time_out_buf = reset, time_out
        # _compute_reward
        rew = np.zeros(n, f32)
        terms = {}
        for name, scale in self.scales:
            r = (self.reward(name).astype(f32) * f32(scale)).astype(f32)
            rew = rew + r
            terms[name] = r
        if rw["only_positive_rewards"]:
            rew = np.maximum(rew, f32(0.0))
        self.reset_idx(np.nonzero(reset)[0], table)
        self.teleport()
        self.resample_commands(table)
        obs, priv = self.observations(table, noise)
        s["last_actions"] = s["actions"].copy()
        s["last_dof_vel"] = s["dof_vel"].copy()
        s["last_root_vel"] = rs[:, 7:13].copy()
        s["last_feet_pos"] = s["feet_pos"].copy()
        return dict(obs=obs, priv=priv, rew=rew, reset_buf=reset, time_out_buf=time_out, extras_time_outs=self.extras_time_outs,
                    terms=terms)

    def reset_all(self, table):
        self.time_out_buf = np.zeros(self.n, bool)
        self.reset_idx(np.arange(self.n), table)
        self.resample_commands(table)
        obs, priv = self.observations(table, True)
        return dict(obs=obs, priv=priv)
