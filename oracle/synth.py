"""Seeded synthetic env states for parity tests and fixtures.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def synthetic_state(n, seed, trimesh, ep_case=True):
    """plausible but adversarial env state: some envs below the termination height, some too fast, some at the episode /
    command-resample boundaries, joint angles beyond limits, feet near the ground on both sides of the contact band."""
    g = np.random.default_rng(seed)
    f = np.float32
    st = {}
    rs = np.zeros((n, 13), f)
    if trimesh:
        rs[:, 0] = g.uniform(-3.0, 83.0, n)
        rs[:, 1] = g.uniform(-3.0, 13.0, n)
        rs[: n // 8, 0] = g.uniform(-4.9, -3.8, n // 8)          # outside the x border: teleport
        rs[n // 8: n // 4, 1] = g.uniform(13.8, 14.2, n // 8)    # outside the y border
    else:
        rs[:, 0:2] = g.uniform(-5, 60, (n, 2))
    rs[:, 2] = g.uniform(0.40, 0.80, n)
    ax = g.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = g.uniform(-0.6, 0.6, n)
    rs[:, 3:6] = ax * np.sin(ang / 2)[:, None]
    rs[:, 6] = np.cos(ang / 2)
    rs[:, 7:13] = g.normal(0, 1.0, (n, 6))
    rs[::7, 7:10] *= 6.0                                          # terminate_vel
    st["root_states"] = rs
    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2, f)
    st["dof_pos"] = (q0 + g.normal(0, 0.5, (n, 12))).astype(f)
    st["dof_vel"] = g.normal(0, 3.0, (n, 12)).astype(f)
    st["actions"] = np.zeros((n, 12), f)                          # overwritten by step()
    st["last_actions"] = g.uniform(-1, 1, (n, 12)).astype(f)
    st["last_dof_vel"] = g.normal(0, 3.0, (n, 12)).astype(f)
    st["last_root_vel"] = g.normal(0, 1.0, (n, 6)).astype(f)
    st["last_dof_targets"] = (q0 + g.normal(0, 0.3, (n, 12))).astype(f)
    st["torques"] = np.zeros((n, 12), f)
    st["commands"] = g.uniform(-1, 1, (n, 3)).astype(f)
    gf = g.uniform(1, 2, n).astype(f); gf[::5] = 0.0
    st["gait_frequency"] = gf
    st["gait_process"] = g.uniform(0, 1, n).astype(f)
    st["filtered_lin_vel"] = g.normal(0, 0.5, (n, 3)).astype(f)
    st["filtered_ang_vel"] = g.normal(0, 0.5, (n, 3)).astype(f)
    st["pushing_forces"] = g.normal(0, 10, (n, 3)).astype(f)
    st["pushing_torques"] = g.normal(0, 2, (n, 3)).astype(f)
    fp = np.zeros((n, 2, 3), f)
    fp[:, :, 0:2] = rs[:, None, 0:2] + g.normal(0, 0.15, (n, 2, 2))
    fp[:, :, 2] = g.uniform(0.0, 0.12, (n, 2))
    st["feet_pos"] = fp.reshape(n, 6)
    fq = g.normal(size=(n, 2, 4)) * np.array([0.1, 0.1, 0.5, 1.0])
    fq /= np.linalg.norm(fq, axis=2, keepdims=True)
    st["feet_quat"] = fq.reshape(n, 8).astype(f)
    st["last_feet_pos"] = (fp + g.normal(0, 0.01, (n, 2, 3))).reshape(n, 6).astype(f)
    st["dof_stiffness"] = (np.array([200, 200, 200, 200, 50, 50] * 2, f) * g.uniform(0.95, 1.05, (n, 12))).astype(f)
    st["dof_damping"] = (np.array([5, 5, 5, 5, 1, 1] * 2, f) * g.uniform(0.95, 1.05, (n, 12))).astype(f)
    st["dof_friction"] = g.uniform(0, 2, (n, 12)).astype(f)
    st["base_mass_scaled"] = g.uniform(0, 1, (n, 4)).astype(f)
    eo = np.zeros((n, 3), f)
    if trimesh:
        eo[:, 0] = g.uniform(5, 75, n)
        eo[:, 1] = g.uniform(1.5, 8.5, n)
    else:
        eo[:, 0:2] = g.uniform(0, 60, (n, 2))
    st["env_origins"] = eo
    ep = g.integers(0, 1400, n)
    crt = ep + g.integers(1, 500, n)
    if ep_case:
        ep[1::9] = 1500                 # -> 1501 after the increment: episode time-out
        k = np.arange(2, n, 3)
        crt[k] = ep[k] + 1              # command resample boundary (time_out without reset)
        ep[3::11] = 0                   # feet_slip mask (episode_length_buf > 1)
        crt[3::11] = 400
    st["episode_length_buf"] = ep.astype(np.int64)
    st["cmd_resample_time"] = crt.astype(np.int64)
    st["delay_steps"] = g.integers(0, 10, n).astype(np.int64)
    return st
