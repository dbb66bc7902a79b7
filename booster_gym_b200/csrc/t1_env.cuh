// t1_env.cuh - per-environment bodies of T1.step()/reset() (envs/t1.py:294-603) as __host__ __device__ functions over
// the structure-of-arrays state (t1_state.h).  One call handles ONE environment `e`; the CUDA kernels map one thread
// per environment so every state access `f[row * n + e]` is coalesced across the warp.
//
// Order of operations, quirks and fp32 operation order follow the reference line by line (citations inline;
// SURVEY 8a notes 1-11).  The host build in tests/hostcheck is test infrastructure only.
#pragma once
#include <string.h>

#include "rng.cuh"
#include "t1_dynamics.cuh"
#include "t1_state.h"
#include "terrain.cuh"

namespace b200 {

#define FS(row) f[(size_t)(row) * (size_t)n + (size_t)e]
#define IS(row) is[(size_t)(row) * (size_t)n + (size_t)e]

struct EnvView {
    float* f;
    int32_t* is;
    int n;             // envs in this shard
    int env_base;      // global index of env 0 of this shard (RNG counter)
    uint64_t seed;
    // Parity-test hook: when non-null, step-time random draws are READ from this table instead of generated, so the
    // kernel, the CPU oracle and the reference (with torch.randn_like & co. patched) run on identical samples.
    // Layout: uint32 [B200_RNG_SLOTS][12][n]: per slot 4 raw words, 4 uniforms (float bits), 4 normals (float bits).
    const uint32_t* inject;
    // command curriculum (null when off): per-cell success counts of this step (k_curriculum_apply turns them into the grid
    // update) and the running sums of the grid the resampling envs draw from
    int32_t* cur_count;
    const float* cur_cdf;
};

// slot of a step-time draw in the injection table; -1 for the one-off construction draws (never injected)
B200_HD int rng_slot(int purpose, int sub) {
    switch (purpose) {
        case RP_RESET_DOF: return 0 + sub;   // 0..2
        case RP_RESET_ROOT: return 3 + sub;  // 3..4
        case RP_RESET_DELAY: return 5;
        case RP_COMMAND: return 6 + sub;     // 6..7
        case RP_KICK: return 8 + sub;        // 8..9
        case RP_PUSH: return 10 + sub;       // 10..11
        case RP_OBS_NOISE: return 12 + sub;  // 12..20
    }
    return -1;
}
#define B200_RNG_SLOTS 21

struct Draw {
    Philox4 p;
    Rand4 r;
};
B200_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
B200_HD Draw env_draw(const EnvView& v, int e, uint32_t env_key, uint64_t step, int purpose, int sub) {
    Draw d;
    const int slot = v.inject ? rng_slot(purpose, sub) : -1;
    if (slot >= 0) {
        const uint32_t* base = v.inject + ((size_t)slot * 12) * (size_t)v.n + (size_t)e;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            d.p.w[l] = base[(size_t)l * v.n];
            d.r.u[l] = bits_to_float(base[(size_t)(4 + l) * v.n]);
            d.r.n[l] = bits_to_float(base[(size_t)(8 + l) * v.n]);
        }
    } else {
        d.p = rng_words(v.seed, env_key, step, purpose, sub);
        d.r = rand4(d.p);
    }
    return d;
}

// ---- Isaac Gym torch_utils semantics (third-party, restated) ---------------------------------------
B200_HD_CALL void quat_rotate_inverse(const float* q, const float* v, float* o) {
    const float qw = q[3];
    const float s = 2.0f * qw * qw - 1.0f;
    float c[3];
    cross3(q, v, c);
    const float d = q[0] * v[0] + q[1] * v[1] + q[2] * v[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = v[i] * s - c[i] * qw * 2.0f + q[i] * d * 2.0f;
}
B200_HD_CALL void quat_rotate(const float* q, const float* v, float* o) {
    const float qw = q[3];
    const float s = 2.0f * qw * qw - 1.0f;
    float c[3];
    cross3(q, v, c);
    const float d = q[0] * v[0] + q[1] * v[1] + q[2] * v[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = v[i] * s + c[i] * qw * 2.0f + q[i] * d * 2.0f;
}
B200_HD float py_mod(float a, float b) {  // torch.remainder / Python % for b > 0
    // fmodf(a, b) == a exactly while |a| < b - every call site's common case (angles from atan2f, the gait phase); the library
    // fmodf is a division loop that showed up with 5 % of k_post's stall samples
    float r = (fabsf(a) < b) ? a : fmodf(a, b);
    if (r != 0.0f && r < 0.0f) r += b;
    return r;
}
#define B200_PI_F 3.14159265358979323846f
#define B200_2PI_F 6.28318530717958647692f
B200_HD_CALL void get_euler_xyz(const float* q, float& roll, float& pitch, float& yaw) {
    const float qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    const float sinr_cosp = 2.0f * (qw * qx + qy * qz);
    const float cosr_cosp = qw * qw - qx * qx - qy * qy + qz * qz;
    roll = atan2f(sinr_cosp, cosr_cosp);
    const float sinp = 2.0f * (qw * qy - qz * qx);
    pitch = (fabsf(sinp) >= 1.0f) ? copysignf(B200_PI_F / 2.0f, sinp) : asinf(sinp);
    const float siny_cosp = 2.0f * (qw * qz + qx * qy);
    const float cosy_cosp = qw * qw + qx * qx - qy * qy - qz * qz;
    yaw = atan2f(siny_cosp, cosy_cosp);
    roll = py_mod(roll, B200_2PI_F);
    pitch = py_mod(pitch, B200_2PI_F);
    yaw = py_mod(yaw, B200_2PI_F);
}
B200_HD float wrap_pi(float a) { return py_mod(a + B200_PI_F, B200_2PI_F) - B200_PI_F; }

// utils/utils.py:5-30.  For the uniform law `r.b` holds (hi - lo) evaluated in fp64 on the host, like Python does.
B200_HD float apply_rand(float x, const B200Rand& r, float u, float nrm) {
    if (!r.enabled) return x;
    const float nv = (r.dist == 0) ? (r.a + r.b * nrm) : (r.a + r.b * u);
    return (r.op == 0) ? (x + nv) : (x * nv);
}
B200_HD float raw_rand(const B200Rand& r, float u, float nrm) { return (r.dist == 0) ? nrm : u; }

// ---- envs/t1.py:439-456 + the physics engine: the decimated PD-torque loop ----------------------------------------
// Device only: one environment runs on a PAIR of adjacent lanes (side = lane & 1 = left / right leg, see the leg-parallel
// formulation in t1_dynamics.cuh); both lanes carry the base redundantly.  act6: this leg's 6 actions (policy output,
// unclipped) or raw torques if !apply_pd.  Lanes whose env index is out of range compute on a clamped index and store
// nothing (they must still take part in the shuffles).
#if defined(__CUDACC__)
template <typename T> __device__ __forceinline__ T pair_sum(T x) { return x + __shfl_xor_sync(0xffffffffu, x, 1); }

template <typename Model>
__device__ __forceinline__ void env_physics_pair(const EnvView& v, int e, bool valid, int side, const Model& m, const B200T1Config& c,
                                                 const TerrainView& terr, const float* act6, int n_substeps, int apply_pd,
                                                 float* qacc_out /*[18][n] or null*/) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    LegState<float> s;
    LegParams<float> par;
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.pos[i] = FS(F_root_states + i); s.vlin[i] = FS(F_root_states + 7 + i); }
#pragma unroll
    for (int i = 0; i < 4; ++i) s.quat[i] = FS(F_root_states + 3 + i);
    {
        const float ww[3] = {FS(F_root_states + 10), FS(F_root_states + 11), FS(F_root_states + 12)};
        float R0[3][3];
        quat_to_mat(s.quat, R0);
#pragma unroll
        for (int k = 0; k < 3; ++k) s.wb[k] = R0[0][k] * ww[0] + R0[1][k] * ww[1] + R0[2][k] * ww[2];
    }
    const int j0 = 6 * side, b0 = 1 + 6 * side;
    par.mass0 = FS(F_body_mass);
#pragma unroll
    for (int r = 0; r < 3; ++r) par.com0[r] = FS(F_body_com + r);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        s.q[k] = FS(F_dof_pos + j0 + k);
        s.qd[k] = FS(F_dof_vel + j0 + k);
        par.mass[k] = FS(F_body_mass + b0 + k);
#pragma unroll
        for (int r = 0; r < 3; ++r) par.com[k][r] = FS(F_body_com + 3 * (b0 + k) + r);
    }
    par.mu = FS(F_foot_friction + side);
    par.kscale = FS(F_foot_kscale + side);
    par.cscale = FS(F_foot_cscale + side);
    float push_f[3], push_t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { push_f[r] = FS(F_pushing_forces + r); push_t[r] = FS(F_pushing_torques + r); }

    float target[6], last_target[6], tsum[6], kp[6], kd[6], fr[6], tlim[6];
    const int delay = IS(I_delay_steps);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (apply_pd) {
            const float a = fminf(fmaxf(act6[k], -c.clip_actions), c.clip_actions);  // envs/t1.py:439
            if (valid) FS(F_actions + j0 + k) = a;
            target[k] = c.default_dof_pos[j0 + k] + c.action_scale * a;             // :440
        } else {
            target[k] = act6[k];
        }
        last_target[k] = FS(F_last_dof_targets + j0 + k);
        kp[k] = FS(F_dof_stiffness + j0 + k);
        kd[k] = FS(F_dof_damping + j0 + k);
        fr[k] = FS(F_dof_friction + j0 + k);
        tlim[k] = c.torque_limits[j0 + k];
        tsum[k] = 0.0f;
    }
    LegWork<float> W;
    W.foot_fn = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { W.trunk_f[r] = 0.0f; W.body_f2[r] = 0.0f; }
    float qb[6], ql[6];
    // per-thread slices of shared memory: geometry and contact matrices of the rarely touching shapes (t1_dynamics.cuh)
    __shared__ float kx_scratch[B200_KX_SIZE * PHYS_BLOCK];
    float* Kx = kx_scratch + threadIdx.x;
    const float* Kp = kx_scratch + (threadIdx.x ^ 1);   // the partner leg lane's slice (leg-leg contacts)
    for (int i = 0; i < n_substeps; ++i) {
        float tau[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (apply_pd) {
                if (delay == i) last_target[k] = target[k];                                 // :445
                float t = kp[k] * (last_target[k] - s.q[k]) - kd[k] * s.qd[k];             // :446
                const float fric = fminf(fr[k], fabsf(t)) * ((t > 0.0f) ? 1.0f : ((t < 0.0f) ? -1.0f : 0.0f));  // :447
                t = fminf(fmaxf(t - fric, -tlim[k]), tlim[k]);                             // :448
                tau[k] = t;
                tsum[k] += t;                                                              // :449
            } else {
                tau[k] = target[k];
            }
        }
        t1_leg_phase1<float, PHYS_BLOCK>(m, par, s, side, tau, push_f, push_t, terr, W, Kx, Kp);
#pragma unroll
        for (int q = 0; q < 21; ++q) W.Mbb[q] = pair_sum(W.Mbb[q]);
#pragma unroll
        for (int q = 0; q < 6; ++q) W.rb[q] = pair_sum(W.rb[q]);
        t1_leg_phase2<float>(m, s, W, qb, ql, true);
        // Isaac Gym applies the tensors of apply_rigid_body_force_tensors (registered once per step(), envs/t1.py:522-527) to the
        // next simulate() only: the push acts on the first substep
        if (i == 0 && !c.push_all_substeps) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { push_f[r] = 0.0f; push_t[r] = 0.0f; }
        }
    }
    // net contact force per body after the last substep (gym.refresh_net_contact_force_tensor, envs/t1.py:462) reduced to what
    // the env reads: |F| > 1 N (envs/t1.py:553,628).  The trunk's force is the sum of the two lanes' shares.
    int cmask = 0;
    {
        const float tf[3] = {pair_sum(W.trunk_f[0]), pair_sum(W.trunk_f[1]), pair_sum(W.trunk_f[2])};
        if (side == 0 && tf[0] * tf[0] + tf[1] * tf[1] + tf[2] * tf[2] > 1.0f) cmask |= 1;
        if (W.body_f2[0] > 1.0f) cmask |= 1 << (1 + 6 * side + 2);
        if (W.body_f2[1] > 1.0f) cmask |= 1 << (1 + 6 * side + 3);
        if (W.body_f2[2] > 1.0f) cmask |= 1 << (1 + 6 * side + 5);
    }
    if (!valid) return;
    const float inv_substeps = 1.0f / (float)n_substeps;   // one division instead of six (:456 divides; <= 1 ulp apart)
    IS(I_contact_mask + side) = cmask;
    // write back (refresh_* tensors of envs/t1.py:454,460-462): the left-leg lane stores the base
    if (side == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { FS(F_root_states + i) = s.pos[i]; FS(F_root_states + 7 + i) = s.vlin[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) FS(F_root_states + 3 + i) = s.quat[i];
        float R0[3][3];
        quat_to_mat(s.quat, R0);
#pragma unroll
        for (int r = 0; r < 3; ++r) FS(F_root_states + 10 + r) = R0[r][0] * s.wb[0] + R0[r][1] * s.wb[1] + R0[r][2] * s.wb[2];
        if (qacc_out) {
#pragma unroll
            for (int i = 0; i < 6; ++i) qacc_out[(size_t)i * n + e] = qb[i];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        FS(F_dof_pos + j0 + k) = s.q[k];
        FS(F_dof_vel + j0 + k) = s.qd[k];
        if (apply_pd) {
            FS(F_last_dof_targets + j0 + k) = last_target[k];
            FS(F_torques + j0 + k) = tsum[k] * inv_substeps;                          // :456
        }
        if (qacc_out) qacc_out[(size_t)(6 + j0 + k) * n + e] = ql[k];
    }
    // pose of this foot link after the last tick (rigid_body_state rows read at envs/t1.py:223-224,530-531)
    float fp[3], fq[4];
    t1_foot_fk<float>(m, s, side, fp, fq);
#pragma unroll
    for (int r = 0; r < 3; ++r) FS(F_feet_pos + 3 * side + r) = fp[r];
#pragma unroll
    for (int r = 0; r < 4; ++r) FS(F_feet_quat + 4 * side + r) = fq[r];
    FS(F_feet_force + side) = W.foot_fn;
}
#endif  // __CUDACC__

// ---- envs/t1.py:529-549 ------------------------------------------------------------------------------------------
// FEET: bit k set = refresh foot k (the two feet are independent: k_post_pair refreshes one per warp)
template <typename Model, int FEET = 3>
B200_HD void env_refresh_feet(const EnvView& v, int e, const Model& m, const TerrainView& terr) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!((FEET >> k) & 1)) continue;
        float q[4], p[3];
#pragma unroll
        for (int r = 0; r < 4; ++r) q[r] = FS(F_feet_quat + 4 * k + r);
#pragma unroll
        for (int r = 0; r < 3; ++r) p[r] = FS(F_feet_pos + 3 * k + r);
        float roll, pitch, yaw;
        get_euler_xyz(q, roll, pitch, yaw);
        FS(F_feet_roll + k) = wrap_pi(roll);  // :533
        FS(F_feet_yaw + k) = wrap_pi(yaw);    // :534
        bool contact = false;
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
            const float rel[3] = {m.foot_corner[cidx][0], m.foot_corner[cidx][1], m.foot_corner[cidx][2]};
            float w[3];
            quat_rotate(q, rel, w);
            const float ex = p[0] + w[0], ey = p[1] + w[1], ez = p[2] + w[2];  // :543
            contact = contact || (ez - terr(ex, ey) < 0.01f);               // :544-549
        }
        IS(I_feet_contact + k) = contact ? 1 : 0;
    }
}

// Everything the 26 reward functions read, loaded ONCE into registers after the kick / push / termination bookkeeping:
// the loads are independent and issue back to back (the pass is latency-bound, not bandwidth-bound), and the stores the
// reward loop performs cannot force reloads through pointer aliasing.
struct RewardSnap {
    float rs[13];   // root_states
    float dof_pos[12], dof_vel[12], torques[12], actions[12], last_actions[12], last_dof_vel[12];
    float last_root_vel[6], feet_pos[6], last_feet_pos[6];
    float commands[3], filt_lin[3], filt_ang[3], base_ang[3], pg[3];
    float feet_roll[2], feet_yaw[2];
    int contact[2];
    float gait_process, gait_frequency;
    int ep_len;
    int contact_mask;
};
B200_HD void reward_snapshot(const EnvView& v, int e, RewardSnap& s) {
    const float* f = v.f;
    const int32_t* is = v.is;
    const int n = v.n;
#pragma unroll
    for (int i = 0; i < 13; ++i) s.rs[i] = FS(F_root_states + i);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        s.dof_pos[j] = FS(F_dof_pos + j); s.dof_vel[j] = FS(F_dof_vel + j); s.torques[j] = FS(F_torques + j);
        s.actions[j] = FS(F_actions + j); s.last_actions[j] = FS(F_last_actions + j); s.last_dof_vel[j] = FS(F_last_dof_vel + j);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        s.last_root_vel[j] = FS(F_last_root_vel + j); s.feet_pos[j] = FS(F_feet_pos + j); s.last_feet_pos[j] = FS(F_last_feet_pos + j);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        s.commands[r] = FS(F_commands + r); s.filt_lin[r] = FS(F_filtered_lin_vel + r); s.filt_ang[r] = FS(F_filtered_ang_vel + r);
        s.base_ang[r] = FS(F_base_ang_vel + r); s.pg[r] = FS(F_projected_gravity + r);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) { s.feet_roll[k] = FS(F_feet_roll + k); s.feet_yaw[k] = FS(F_feet_yaw + k); s.contact[k] = IS(I_feet_contact + k); }
    s.gait_process = FS(F_gait_process);
    s.gait_frequency = FS(F_gait_frequency);
    s.ep_len = IS(I_episode_length_buf);
    s.contact_mask = IS(I_contact_mask) | IS(I_contact_mask + 1);
}

// reward term k of envs/t1.py:606-730 (unscaled); `h_base` = terrain height under the base
B200_HD float reward_term(int id, const RewardSnap& s, const B200T1Config& c, float h_base) {
    switch (id) {
        case B200_REW_SURVIVAL: return 1.0f;
        case B200_REW_TRACK_LIN_X: { const float d = s.commands[0] - s.filt_lin[0]; return expf(-(d * d) / c.tracking_sigma); }
        case B200_REW_TRACK_LIN_Y: { const float d = s.commands[1] - s.filt_lin[1]; return expf(-(d * d) / c.tracking_sigma); }
        case B200_REW_TRACK_ANG: { const float d = s.commands[2] - s.filt_ang[2]; return expf(-(d * d) / c.tracking_sigma); }
        case B200_REW_BASE_HEIGHT: { const float bh = s.rs[2] - h_base; const float d = bh - c.base_height_target; return d * d; }
        case B200_REW_ORIENTATION: return s.pg[0] * s.pg[0] + s.pg[1] * s.pg[1];
        case B200_REW_TORQUES: { float a = 0; for (int j = 0; j < 12; ++j) a += s.torques[j] * s.torques[j]; return a; }
        case B200_REW_TORQUE_TIREDNESS: { float a = 0; for (int j = 0; j < 12; ++j) { const float t = s.torques[j] / c.torque_limits[j]; a += fminf(t * t, 1.0f); } return a; }
        case B200_REW_POWER: { float a = 0; for (int j = 0; j < 12; ++j) a += fmaxf(s.torques[j] * s.dof_vel[j], 0.0f); return a; }
        case B200_REW_LIN_VEL_Z: return s.filt_lin[2] * s.filt_lin[2];
        case B200_REW_ANG_VEL_XY: return s.base_ang[0] * s.base_ang[0] + s.base_ang[1] * s.base_ang[1];
        case B200_REW_DOF_VEL: { float a = 0; for (int j = 0; j < 12; ++j) a += s.dof_vel[j] * s.dof_vel[j]; return a; }
        case B200_REW_DOF_ACC: { float a = 0; for (int j = 0; j < 12; ++j) { const float t = (s.last_dof_vel[j] - s.dof_vel[j]) / c.env_dt; a += t * t; } return a; }
        case B200_REW_ROOT_ACC: { float a = 0; for (int j = 0; j < 6; ++j) { const float t = (s.last_root_vel[j] - s.rs[7 + j]) / c.env_dt; a += t * t; } return a; }
        case B200_REW_ACTION_RATE: { float a = 0; for (int j = 0; j < 12; ++j) { const float t = s.last_actions[j] - s.actions[j]; a += t * t; } return a; }
        case B200_REW_DOF_POS_LIMITS: { float a = 0; for (int j = 0; j < 12; ++j) a += ((s.dof_pos[j] < c.dof_pos_soft_lower[j]) || (s.dof_pos[j] > c.dof_pos_soft_upper[j])) ? 1.0f : 0.0f; return a; }
        case B200_REW_DOF_VEL_LIMITS: { float a = 0; for (int j = 0; j < 12; ++j) a += fminf(fmaxf(fabsf(s.dof_vel[j]) - c.dof_vel_limits[j] * c.soft_dof_vel_limit, 0.0f), 1.0f); return a; }
        case B200_REW_TORQUE_LIMITS: { float a = 0; for (int j = 0; j < 12; ++j) a += fmaxf(fabsf(s.torques[j]) - c.torque_limits[j] * c.soft_torque_limit, 0.0f); return a; }
        case B200_REW_COLLISION: {  // :627-629: number of penalised bodies whose net contact force exceeds 1 N
            int cnt = 0;
            for (int mk = s.contact_mask & c.penalized_body_mask; mk; mk &= mk - 1) ++cnt;
            return (float)cnt;
        }
        case B200_REW_FEET_SLIP: {
            float a = 0;
            for (int k = 0; k < 2; ++k) {
                float q = 0;
                for (int r = 0; r < 3; ++r) { const float t = (s.last_feet_pos[3 * k + r] - s.feet_pos[3 * k + r]) / c.env_dt; q += t * t; }
                a += q * (s.contact[k] ? 1.0f : 0.0f);
            }
            return a * ((s.ep_len > 1) ? 1.0f : 0.0f);
        }
        case B200_REW_FEET_VEL_Z: { float a = 0; for (int k = 0; k < 2; ++k) { const float t = (s.last_feet_pos[3 * k + 2] - s.feet_pos[3 * k + 2]) / c.env_dt; a += t * t; } return a; }
        case B200_REW_FEET_YAW_DIFF: { const float d = wrap_pi(s.feet_yaw[1] - s.feet_yaw[0]); return d * d; }
        case B200_REW_FEET_YAW_MEAN: {
            const float y0 = s.feet_yaw[0], y1 = s.feet_yaw[1];
            const float mean = (y0 + y1) / 2.0f + B200_PI_F * ((fabsf(y1 - y0) > B200_PI_F) ? 1.0f : 0.0f);
            float r, p, y;
            get_euler_xyz(s.rs + 3, r, p, y);
            const float d = wrap_pi(y - mean);
            return d * d;
        }
        case B200_REW_FEET_ROLL: return s.feet_roll[0] * s.feet_roll[0] + s.feet_roll[1] * s.feet_roll[1];
        case B200_REW_FEET_DISTANCE: {
            float r, p, y;
            get_euler_xyz(s.rs + 3, r, p, y);
            const float dist = fabsf(cosf(y) * (s.feet_pos[3 + 1] - s.feet_pos[1]) - sinf(y) * (s.feet_pos[3 + 0] - s.feet_pos[0]));
            return fminf(fmaxf(c.feet_distance_ref - dist, 0.0f), 0.1f);
        }
        case B200_REW_FEET_SWING: {
            const bool moving = s.gait_frequency > 1.0e-8f;
            const bool ls = (fabsf(s.gait_process - 0.25f) < 0.5f * c.swing_period) && moving;
            const bool rsw = (fabsf(s.gait_process - 0.75f) < 0.5f * c.swing_period) && moving;
            return ((ls && !s.contact[0]) ? 1.0f : 0.0f) + ((rsw && !s.contact[1]) ? 1.0f : 0.0f);
        }
    }
    return 0.0f;
}

// envs/t1.py:391-413 for ONE resetting env: a successful episode (long enough, command tracked) raises its grid cell and the
// 4 neighbours by update_rate.  The adds are recorded as integer COUNTS per cell; k_curriculum_apply performs them as the
// reference does (count sequential fp32 adds of the rate, then clamp to 1) - order-free and bit-exact.
B200_HD void env_update_curriculum(const EnvView& v, int e, const B200T1Config& c) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    bool success = (float)IS(I_episode_length_buf) > c.cur_success_len;
    success = success && (fabsf(FS(F_filtered_lin_vel + 0) - FS(F_commands + 0)) < c.cur_tol_x);
    success = success && (fabsf(FS(F_filtered_lin_vel + 1) - FS(F_commands + 1)) < c.cur_tol_y);
    success = success && (fabsf(FS(F_filtered_ang_vel + 2) - FS(F_commands + 2)) < c.cur_tol_yaw);
    if (!success || v.cur_count == nullptr) return;
    const int rows = 2 * c.cur_lin_levels + 1, cols = 2 * c.cur_ang_levels + 1;
    const int x = IS(I_env_curriculum_level + 0) + c.cur_lin_levels, y = IS(I_env_curriculum_level + 1) + c.cur_ang_levels;
    if (x < 0 || x >= rows || y < 0 || y >= cols) return;
#if defined(__CUDA_ARCH__)
#define B200_CUR_INC(ix) atomicAdd(v.cur_count + (ix), 1)
#else
#define B200_CUR_INC(ix) (v.cur_count[(ix)] += 1)
#endif
    B200_CUR_INC(x * cols + y);
    if (x > 0) B200_CUR_INC((x - 1) * cols + y);
    if (x < rows - 1) B200_CUR_INC((x + 1) * cols + y);
    if (y > 0) B200_CUR_INC(x * cols + y - 1);
    if (y < cols - 1) B200_CUR_INC(x * cols + y + 1);
#undef B200_CUR_INC
}

// the grid update of one step (:404-413) and the running sums for the draws of :416 - ONE thread per cell, thread 0 scans
template <typename IntT>
B200_HD void curriculum_apply_cell(float* prob, IntT* count, int i, float rate) {
    float p = prob[i];
    for (int k = (int)count[i]; k > 0; --k) p = p + rate;
    prob[i] = fminf(p, 1.0f);
    count[i] = 0;
}
B200_HD void curriculum_scan(const float* prob, float* cdf, int cells) {
    float run = 0.0f;
    for (int i = 0; i < cells; ++i) { run = run + prob[i]; cdf[i] = run; }
}

// envs/t1.py:301-341 for ONE env (always resets; the caller decides).  `step` = RNG counter of this call.
B200_HD void env_reset_one(const EnvView& v, int e, const B200T1Config& c, const TerrainView& terr, uint64_t step) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    const uint32_t ge = (uint32_t)(v.env_base + e);
    if (c.curriculum) env_update_curriculum(v, e, c);   // :305, before the episode's state is overwritten
    // _reset_dofs :319-325.  Reference quirk (SURVEY 8a note 12): the noise has the shape of default_dof_pos, [1,12],
    // so every env that resets in the same call receives the SAME 12 joint offsets -> the draw is keyed by the step
    // only (env id B200_RNG_SHARED_ENV), not by the env.
#pragma unroll
    for (int sub = 0; sub < 3; ++sub) {
        const Rand4 r = env_draw(v, e, B200_RNG_SHARED_ENV, step, RP_RESET_DOF, sub).r;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int j = 4 * sub + l;
            FS(F_dof_pos + j) = apply_rand(c.default_dof_pos[j], c.init_dof_pos, r.u[l], r.n[l]);
            FS(F_dof_vel + j) = 0.0f;
        }
    }
    // _reset_root_states :327-341
    const Rand4 r0 = env_draw(v, e, ge, step, RP_RESET_ROOT, 0).r;
    const Rand4 r1 = env_draw(v, e, ge, step, RP_RESET_ROOT, 1).r;
    float x = c.init_root[0] + FS(F_env_origins + 0);
    float y = c.init_root[1] + FS(F_env_origins + 1);
    x = apply_rand(x, c.init_base_pos_xy, r0.u[0], r0.n[0]);
    y = apply_rand(y, c.init_base_pos_xy, r0.u[1], r0.n[1]);
    FS(F_root_states + 0) = x;
    FS(F_root_states + 1) = y;
    FS(F_root_states + 2) = c.init_root[2] + terr(x, y);
    {
        // quat_from_euler_xyz(0, 0, yaw), yaw = rand * 2 pi  (:332-336)
        const float yaw = r0.u[2] * B200_2PI_F;
        const float cy = cosf(yaw * 0.5f), sy = sinf(yaw * 0.5f);
        // cr = cp = 1, sr = sp = 0  (cos(0) = 1 exactly)
        FS(F_root_states + 3) = 0.0f;  // qx = cy*sr*cp - sy*cr*sp
        FS(F_root_states + 4) = 0.0f;  // qy = cy*cr*sp + sy*sr*cp
        FS(F_root_states + 5) = sy;    // qz = sy*cr*cp - cy*sr*sp
        FS(F_root_states + 6) = cy;    // qw = cy*cr*cp + sy*sr*sp
    }
    FS(F_root_states + 7) = apply_rand(0.0f, c.init_base_lin_vel_xy, r1.u[0], r1.n[0]);
    FS(F_root_states + 8) = apply_rand(0.0f, c.init_base_lin_vel_xy, r1.u[1], r1.n[1]);
    FS(F_root_states + 9) = c.init_root[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) FS(F_root_states + 10 + r) = c.init_root[10 + r];
    // :309-316
#pragma unroll
    for (int j = 0; j < 12; ++j) FS(F_last_dof_targets + j) = FS(F_dof_pos + j);
#pragma unroll
    for (int j = 0; j < 6; ++j) FS(F_last_root_vel + j) = FS(F_root_states + 7 + j);
    IS(I_episode_length_buf) = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) { FS(F_filtered_lin_vel + r) = 0.0f; FS(F_filtered_ang_vel + r) = 0.0f; }
    IS(I_cmd_resample_time) = 0;
    const Philox4 d = env_draw(v, e, ge, step, RP_RESET_DELAY, 0).p;
    IS(I_delay_steps) = (int32_t)(d.w[0] % (uint32_t)c.decimation);
}

// envs/t1.py:362-389 (+ :415-435 when the curriculum is on) for one env whose episode_length == cmd_resample_time
B200_HD void env_resample_command(const EnvView& v, int e, const B200T1Config& c, uint64_t step) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    const uint32_t ge = (uint32_t)(v.env_base + e);
    const Draw d0 = env_draw(v, e, ge, step, RP_COMMAND, 0);
    const Draw d1 = env_draw(v, e, ge, step, RP_COMMAND, 1);
    // torch_rand_float(lo, hi) = (hi - lo) * rand + lo
    float cx, cy, cz;
    if (c.curriculum && v.cur_cdf != nullptr) {
        // _resample_curriculum_commands :415-435.  torch.multinomial(prob.flatten(), ., replacement=True) is restated as the inverse
        // CDF of the same distribution on this env's uniform (no torch sampler on the device): first cell whose running fp32 sum
        // exceeds u * total.  The grid was updated and scanned by k_curriculum_apply after this step's resets (:305 before :488).
        const int cols = 2 * c.cur_ang_levels + 1, cells = (2 * c.cur_lin_levels + 1) * cols;
        const float target = d1.r.u[2] * v.cur_cdf[cells - 1];
        int lo = 0, hi = cells - 1;             // smallest i with cdf[i] > target, clamped to the last cell
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (v.cur_cdf[mid] > target) hi = mid; else lo = mid + 1;
        }
        // reference quirk (SURVEY 8a note 13): the flat index of the [lin, ang] grid is decoded as (idx % cols, idx // cols)
        const int lin = lo % cols - c.cur_lin_levels, ang = lo / cols - c.cur_ang_levels;
        IS(I_env_curriculum_level + 0) = lin;
        IS(I_env_curriculum_level + 1) = ang;
        const float r0 = (0.5f - -0.5f) * d0.r.u[0] + -0.5f, r1 = (1.0f - -1.0f) * d0.r.u[1] + -1.0f, r2 = (0.5f - -0.5f) * d0.r.u[2] + -0.5f;
        cx = ((float)lin + r0) * c.cur_res_x;
        cy = ((float)(lin < 0 ? -lin : lin) * r1) * c.cur_res_y;
        cz = ((float)ang + r2) * c.cur_res_ang;
    } else {
        cx = (c.lin_vel_x[1] - c.lin_vel_x[0]) * d0.r.u[0] + c.lin_vel_x[0];
        cy = (c.lin_vel_y[1] - c.lin_vel_y[0]) * d0.r.u[1] + c.lin_vel_y[0];
        cz = (c.ang_vel_yaw[1] - c.ang_vel_yaw[0]) * d0.r.u[2] + c.ang_vel_yaw[0];
    }
    float gf = (c.gait_frequency[1] - c.gait_frequency[0]) * d0.r.u[3] + c.gait_frequency[0];
    // reference: an exact still_proportion of the resampled envs via randperm (:381); here an independent Bernoulli
    // draw per env (no cross-env dependency on the device) - DESIGN.md "deviations"
    if (d1.r.u[0] < c.still_proportion) { cx = cy = cz = 0.0f; gf = 0.0f; }
    FS(F_commands + 0) = cx; FS(F_commands + 1) = cy; FS(F_commands + 2) = cz;
    FS(F_gait_frequency) = gf;
    IS(I_cmd_resample_time) += c.resample_lo + (int32_t)(d1.p.w[1] % (uint32_t)(c.resample_hi - c.resample_lo));
}

// envs/t1.py:574-603.  obs[47], priv[14] are this env's rows.
B200_HD void env_observations(const EnvView& v, int e, const B200T1Config& c, const TerrainView& terr, uint64_t step,
                              int noise_on, float* obs, float* priv) {
    float* f = v.f;
    int32_t* is = v.is;
    (void)is;
    const int n = v.n;
    const uint32_t ge = (uint32_t)(v.env_base + e);
    float un[36], nn[36];
    if (noise_on) {
#pragma unroll
        for (int sub = 0; sub < 9; ++sub) {
            const Rand4 r = env_draw(v, e, ge, step, RP_OBS_NOISE, sub).r;
#pragma unroll
            for (int l = 0; l < 4; ++l) { un[4 * sub + l] = r.u[l]; nn[4 * sub + l] = r.n[l]; }
        }
    }
    B200Rand off;
    off.enabled = 0; off.dist = 0; off.op = 0; off.a = 0; off.b = 0;
    const B200Rand& ng = noise_on ? c.noise_gravity : off;
    const B200Rand& na = noise_on ? c.noise_ang_vel : off;
    const B200Rand& np_ = noise_on ? c.noise_dof_pos : off;
    const B200Rand& nv = noise_on ? c.noise_dof_vel : off;
    const B200Rand& nl = noise_on ? c.noise_lin_vel : off;
    const B200Rand& nh = noise_on ? c.noise_height : off;
    if (!noise_on) {
#pragma unroll
        for (int i = 0; i < 36; ++i) { un[i] = 0.0f; nn[i] = 0.0f; }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        obs[r] = apply_rand(FS(F_projected_gravity + r), ng, un[r], nn[r]) * c.norm_gravity;
        obs[3 + r] = apply_rand(FS(F_base_ang_vel + r), na, un[3 + r], nn[3 + r]) * c.norm_ang_vel;
    }
    obs[6] = FS(F_commands + 0) * c.norm_lin_vel;
    obs[7] = FS(F_commands + 1) * c.norm_lin_vel;
    obs[8] = FS(F_commands + 2) * c.norm_ang_vel;
    {
        const float ph = B200_2PI_F * FS(F_gait_process);
        const float on = (FS(F_gait_frequency) > 1.0e-8f) ? 1.0f : 0.0f;
        obs[9] = cosf(ph) * on;
        obs[10] = sinf(ph) * on;
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        obs[11 + j] = apply_rand(FS(F_dof_pos + j) - c.default_dof_pos[j], np_, un[6 + j], nn[6 + j]) * c.norm_dof_pos;
        obs[23 + j] = apply_rand(FS(F_dof_vel + j), nv, un[18 + j], nn[18 + j]) * c.norm_dof_vel;
        obs[35 + j] = FS(F_actions + j);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) priv[r] = FS(F_base_mass_scaled + r);
#pragma unroll
    for (int r = 0; r < 3; ++r) priv[4 + r] = apply_rand(FS(F_base_lin_vel + r), nl, un[30 + r], nn[30 + r]) * c.norm_lin_vel;
    priv[7] = apply_rand(FS(F_root_states + 2) - terr(FS(F_root_states + 0), FS(F_root_states + 1)), nh, un[33], nn[33]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        priv[8 + r] = FS(F_pushing_forces + r) * c.norm_push_force;
        priv[11 + r] = FS(F_pushing_torques + r) * c.norm_push_torque;
    }
}

// envs/t1.py:343-360 (trimesh only).  The reference re-runs _refresh_feet_state for everyone when any env moved
// (:358-360); for an env that did not move that recomputation is the identity, so it is done per moved env here.
template <typename Model>
B200_HD void env_teleport(const EnvView& v, int e, const Model& m, const B200T1Config& c, const TerrainView& terr) {
    if (c.terrain_type == 0) return;
    float* f = v.f;
    int32_t* is = v.is;
    (void)is;
    const int n = v.n;
    const float x = FS(F_root_states + 0), y = FS(F_root_states + 1);
    float dx = 0.0f, dy = 0.0f;
    bool xmin = x < -0.75f * c.border_size, xmax = x > c.env_width + 0.75f * c.border_size;
    bool ymin = y < -0.75f * c.border_size, ymax = y > c.env_length + 0.75f * c.border_size;
    if (xmin) dx += c.env_width + c.border_size;
    if (xmax) dx -= c.env_width + c.border_size;
    if (ymin) dy += c.env_length + c.border_size;
    if (ymax) dy -= c.env_length + c.border_size;
    if (xmin || xmax) { FS(F_root_states + 0) = x + dx; FS(F_feet_pos + 0) += dx; FS(F_feet_pos + 3) += dx; }
    if (ymin || ymax) { FS(F_root_states + 1) = y + dy; FS(F_feet_pos + 1) += dy; FS(F_feet_pos + 4) += dy; }
    if (xmin || xmax || ymin || ymax) env_refresh_feet(v, e, m, terr);
}

struct StepOut {
    float rew;
    int done, time_out;
};

// envs/t1.py:460-497 for one env.  common_step = common_step_counter after the increment (:477); step = RNG counter.
// stats (nullable): double[1 + n_rew + 1] sums + int64 count at stats_count, accumulated with atomics on the device.
// PHASE 0 = the whole step.  With the command curriculum the resampling envs draw from a grid that ALL of this step's resets
// have updated first (:305 before :488), a dependency across envs: PHASE 1 stops after the resets (:460-486), the grid is
// updated (k_curriculum_apply), PHASE 2 finishes the step (:487-495).  Nothing but the env's state crosses the cut.
// The step in four parts, so that the device kernel can run independent parts of ONE env on different warps (k_post_pair,
// env_kernels.cu): A = derived state, feet, counters, kicks / pushes, termination (:463-478,499-527,551-558); the reward snapshot;
// R = the reward terms and episode sums (:560-572); [reset :485-486]; T = teleport, command resampling, observations, last_* (:487-495).
// R reads only the snapshot (registers) and the episode sums, which neither the reset nor T touches.
struct PostA {
    float h_base;
    bool finite, reset, time_out;
};
template <typename Model, int FEET = 3>
B200_HD PostA env_post_a(const EnvView& v, int e, const Model& m, const B200T1Config& c, const TerrainView& terr, int64_t common_step,
                         uint64_t step) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    const uint32_t ge = (uint32_t)(v.env_base + e);
    // :463-473
    float q[4] = {FS(F_root_states + 3), FS(F_root_states + 4), FS(F_root_states + 5), FS(F_root_states + 6)};
    {
        const float vl[3] = {FS(F_root_states + 7), FS(F_root_states + 8), FS(F_root_states + 9)};
        const float va[3] = {FS(F_root_states + 10), FS(F_root_states + 11), FS(F_root_states + 12)};
        const float g[3] = {0.0f, 0.0f, -1.0f};
        float bl[3], ba[3], pg[3];
        quat_rotate_inverse(q, vl, bl);
        quat_rotate_inverse(q, va, ba);
        quat_rotate_inverse(q, g, pg);
        const float w = c.filter_weight, w1 = (float)(1.0 - (double)c.filter_weight);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            FS(F_base_lin_vel + r) = bl[r];
            FS(F_base_ang_vel + r) = ba[r];
            FS(F_projected_gravity + r) = pg[r];
            FS(F_filtered_lin_vel + r) = bl[r] * w + FS(F_filtered_lin_vel + r) * w1;
            FS(F_filtered_ang_vel + r) = ba[r] * w + FS(F_filtered_ang_vel + r) * w1;
        }
    }
    env_refresh_feet<Model, FEET>(v, e, m, terr);  // :474
    // :476-478
    const int ep_len = IS(I_episode_length_buf) + 1;
    IS(I_episode_length_buf) = ep_len;
    FS(F_gait_process) = fmodf(FS(F_gait_process) + c.env_dt * FS(F_gait_frequency), 1.0f);
    // _kick_robots :499-504
    if (c.kick_interval > 0 && (common_step % c.kick_interval) == 0) {
        const Rand4 a = env_draw(v, e, ge, step, RP_KICK, 0).r;
        const Rand4 b = env_draw(v, e, ge, step, RP_KICK, 1).r;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            FS(F_root_states + 7 + r) = apply_rand(FS(F_root_states + 7 + r), c.kick_lin_vel, a.u[r], a.n[r]);
            FS(F_root_states + 10 + r) = apply_rand(FS(F_root_states + 10 + r), c.kick_ang_vel, b.u[r], b.n[r]);
        }
    }
    // _push_robots :506-527
    if (c.push_interval > 0) {
        const int64_t ph = common_step % c.push_interval;
        if (ph == 0) {
            const Rand4 a = env_draw(v, e, ge, step, RP_PUSH, 0).r;
            const Rand4 b = env_draw(v, e, ge, step, RP_PUSH, 1).r;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                FS(F_pushing_forces + r) = apply_rand(0.0f, c.push_force, a.u[r], a.n[r]);
                FS(F_pushing_torques + r) = apply_rand(0.0f, c.push_torque, b.u[r], b.n[r]);
            }
        } else if (ph == c.push_duration) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { FS(F_pushing_forces + r) = 0.0f; FS(F_pushing_torques + r) = 0.0f; }
        }
    }
    // _check_termination :551-558
    PostA a;
    a.h_base = terr(FS(F_root_states + 0), FS(F_root_states + 1));
    float vsq = 0.0f;
#pragma unroll
    for (int r = 0; r < 6; ++r) { const float t = FS(F_root_states + 7 + r); vsq += t * t; }
    a.finite = (vsq == vsq) && (fabsf(vsq) <= 3.0e38f) && (FS(F_root_states + 2) == FS(F_root_states + 2));
    const int contact_mask = IS(I_contact_mask) | IS(I_contact_mask + 1);  // bit b: |contact_forces[b]| > 1 (:553)
    bool reset = ((contact_mask & c.termination_body_mask) != 0) || (vsq > c.terminate_vel) || (FS(F_root_states + 2) - a.h_base < c.terminate_height);
    if (!a.finite) { reset = true; IS(I_nan_resets) += 1; }  // a diverged env is reset, not propagated
    bool time_out = ep_len > c.max_episode_length;
    reset = reset || time_out;
    time_out = time_out || (ep_len == IS(I_cmd_resample_time));
    IS(I_reset_buf) = reset ? 1 : 0;
    IS(I_time_out_buf) = time_out ? 1 : 0;
    a.reset = reset;
    a.time_out = time_out;
    return a;
}

// _compute_reward :560-572 from the snapshot; returns the step's reward
B200_HD float env_post_rewards(const EnvView& v, int e, const B200T1Config& c, const RewardSnap& snap, float h_base, bool finite,
                               float* rew_terms /* [n_rew][n] or null */) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    float rew = 0.0f;
    // the per-term episode sums are read 4 terms ahead: a load inside the rolled loop stalled every iteration for a full memory
    // round trip (17 % of k_post's stall samples sat on the one FADD that consumed it)
    const int nr = c.n_rew;
    float es0 = (0 < nr) ? FS(F_episode_sums + 1) : 0.0f, es1 = (1 < nr) ? FS(F_episode_sums + 2) : 0.0f;
    float es2 = (2 < nr) ? FS(F_episode_sums + 3) : 0.0f, es3 = (3 < nr) ? FS(F_episode_sums + 4) : 0.0f;
    for (int k = 0; k < nr; ++k) {
        const float es4 = (k + 4 < nr) ? FS(F_episode_sums + 5 + k) : 0.0f;
        float r = reward_term(c.rew_id[k], snap, c, h_base) * c.rew_scale[k];
        if (!finite) r = 0.0f;
        rew += r;
        if (rew_terms) rew_terms[(size_t)k * (size_t)n + (size_t)e] = r;
        FS(F_episode_sums + 1 + k) = es0 + r;
        es0 = es1; es1 = es2; es2 = es3; es3 = es4;
    }
    if (c.only_positive_rewards) rew = fmaxf(rew, 0.0f);
    FS(F_episode_sums + 0) += rew;
    IS(I_episode_steps) += 1;
    return rew;
}

// :487-495: teleport, command resampling, observations, last_* bookkeeping
template <typename Model>
B200_HD void env_post_tail(const EnvView& v, int e, const Model& m, const B200T1Config& c, const TerrainView& terr, uint64_t step,
                           int noise_on, float* obs, float* priv) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    env_teleport(v, e, m, c, terr);
    if (IS(I_episode_length_buf) == IS(I_cmd_resample_time)) env_resample_command(v, e, c, step);
    // :490
    env_observations(v, e, c, terr, step, noise_on, obs, priv);
    // :492-495
#pragma unroll
    for (int j = 0; j < 12; ++j) { FS(F_last_actions + j) = FS(F_actions + j); FS(F_last_dof_vel + j) = FS(F_dof_vel + j); }
#pragma unroll
    for (int j = 0; j < 6; ++j) { FS(F_last_root_vel + j) = FS(F_root_states + 7 + j); FS(F_last_feet_pos + j) = FS(F_feet_pos + j); }
}

template <typename Model, int PHASE = 0>
B200_HD StepOut env_post_physics(const EnvView& v, int e, const Model& m, const B200T1Config& c, const TerrainView& terr,
                                 int64_t common_step, uint64_t step, int noise_on, float* obs, float* priv,
                                 float* rew_terms /* [n_rew][n] or null */) {
    StepOut o;
    o.rew = 0.0f; o.done = 0; o.time_out = 0;
    if (PHASE != 2) {
        const PostA a = env_post_a(v, e, m, c, terr, common_step, step);
        RewardSnap snap;
        reward_snapshot(v, e, snap);
        const float rew = env_post_rewards(v, e, c, snap, a.h_base, a.finite, rew_terms);
        // :485-488
        if (a.reset) env_reset_one(v, e, c, terr, step);
        o.rew = rew;
        o.done = a.reset ? 1 : 0;
        o.time_out = a.time_out ? 1 : 0;
        if (PHASE == 1) return o;
    }
    env_post_tail(v, e, m, c, terr, step, noise_on, obs, priv);
    return o;
}

// One-off per-env model randomisation + buffer initialisation: envs/t1.py:69-83,139-167,187-272.
template <typename Model>
B200_HD void env_init_params(const EnvView& v, int e, const Model& m, const B200T1Config& c) {
    float* f = v.f;
    int32_t* is = v.is;
    const int n = v.n;
    const uint32_t ge = (uint32_t)(v.env_base + e);
#pragma unroll
    for (int sub = 0; sub < 3; ++sub) {
        const Rand4 a = rand4(rng_words(v.seed, ge, 0, RP_INIT_GAINS, sub));
        const Rand4 b = rand4(rng_words(v.seed, ge, 0, RP_INIT_GAINS, 3 + sub));
        const Rand4 d = rand4(rng_words(v.seed, ge, 0, RP_INIT_GAINS, 6 + sub));
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int j = 4 * sub + l;
            FS(F_dof_stiffness + j) = apply_rand(c.kp_nominal[j], c.dof_stiffness, a.u[l], a.n[l]);
            FS(F_dof_damping + j) = apply_rand(c.kd_nominal[j], c.dof_damping, b.u[l], b.n[l]);
            FS(F_dof_friction + j) = apply_rand(0.0f, c.dof_friction, d.u[l], d.n[l]);
        }
    }
    for (int b = 0; b < B200_NB; ++b) {
        // CoM and mass may be configured with DIFFERENT distribution kinds: a uniform and a normal of one Philox block share words
        // (rng.cuh: u[i] and n[i] both read word i), so the mass draws from its own block
        const Rand4 r = rand4(rng_words(v.seed, ge, 0, RP_INIT_BODY, b));
        const Rand4 rmass = rand4(rng_words(v.seed, ge, 0, RP_INIT_BODY, 16 + b));
        const B200Rand& rc = (b == 0) ? c.base_com : c.other_com;
        const B200Rand& rm = (b == 0) ? c.base_mass : c.other_mass;
#pragma unroll
        for (int k = 0; k < 3; ++k) FS(F_body_com + 3 * b + k) = apply_rand((float)m.ipos[b][k], rc, r.u[k], r.n[k]);
        FS(F_body_mass + b) = apply_rand((float)m.mass[b], rm, rmass.u[0], rmass.n[0]);
        if (b == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) FS(F_base_mass_scaled + k) = rc.enabled ? raw_rand(rc, r.u[k], r.n[k]) : 0.0f;
            FS(F_base_mass_scaled + 3) = rm.enabled ? raw_rand(rm, rmass.u[0], rmass.n[0]) : 0.0f;
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const Rand4 r = rand4(rng_words(v.seed, ge, 0, RP_INIT_FOOT, k));        // friction
        const Rand4 r2 = rand4(rng_words(v.seed, ge, 0, RP_INIT_FOOT, 2 + k));   // compliance   (own blocks: the three keys may be
        const Rand4 r3 = rand4(rng_words(v.seed, ge, 0, RP_INIT_FOOT, 4 + k));   // restitution   configured with different kinds)
        // PhysX material of the foot shapes (envs/t1.py:162-167): friction is combined with the ground by averaging
        // (PhysX default combine mode); compliance / restitution scale this build's contact spring / damper.
        const float mu_foot = c.friction.enabled ? apply_rand(0.0f, c.friction, r.u[0], r.n[0]) : c.terrain_friction;
        FS(F_foot_friction + k) = 0.5f * (mu_foot + c.terrain_friction);
        const float comp = c.compliance.enabled ? apply_rand(0.0f, c.compliance, r2.u[0], r2.n[0]) : 1.0f;
        FS(F_foot_kscale + k) = 1.0f / fmaxf(comp, 0.1f);
        const float rest = c.restitution.enabled ? apply_rand(0.0f, c.restitution, r3.u[0], r3.n[0]) : 0.0f;
        FS(F_foot_cscale + k) = 1.0f - 0.5f * fminf(fmaxf(rest, 0.0f), 1.0f);
    }
    // _init_buffers :193-272 (the values a freshly created sim reports: identity pose at the env origin, at rest)
#pragma unroll
    for (int r = 0; r < 13; ++r) FS(F_root_states + r) = c.init_root[r];
    FS(F_root_states + 0) += FS(F_env_origins + 0);
    FS(F_root_states + 1) += FS(F_env_origins + 1);
    FS(F_root_states + 2) += FS(F_env_origins + 2);
    const int zero_rows[] = {F_dof_pos, F_dof_vel, F_actions, F_last_actions, F_last_dof_vel, F_last_dof_targets, F_torques};
    for (int zr = 0; zr < 7; ++zr)
        for (int j = 0; j < 12; ++j) FS(zero_rows[zr] + j) = 0.0f;
    for (int j = 0; j < 6; ++j) { FS(F_last_root_vel + j) = 0.0f; FS(F_feet_pos + j) = 0.0f; FS(F_last_feet_pos + j) = 0.0f; }
    for (int j = 0; j < 8; ++j) FS(F_feet_quat + j) = (j % 4 == 3) ? 1.0f : 0.0f;
    for (int r = 0; r < 3; ++r) {
        FS(F_commands + r) = 0.0f; FS(F_base_lin_vel + r) = 0.0f; FS(F_base_ang_vel + r) = 0.0f;
        FS(F_projected_gravity + r) = (r == 2) ? -1.0f : 0.0f;
        FS(F_filtered_lin_vel + r) = 0.0f; FS(F_filtered_ang_vel + r) = 0.0f;
        FS(F_pushing_forces + r) = 0.0f; FS(F_pushing_torques + r) = 0.0f;
    }
    FS(F_gait_frequency) = 0.0f; FS(F_gait_process) = 0.0f;
    for (int k = 0; k < 2; ++k) { FS(F_feet_roll + k) = 0.0f; FS(F_feet_yaw + k) = 0.0f; FS(F_feet_force + k) = 0.0f; IS(I_feet_contact + k) = 0; }
    for (int k = 0; k < 27; ++k) FS(F_episode_sums + k) = 0.0f;
    IS(I_episode_length_buf) = 0; IS(I_cmd_resample_time) = 0; IS(I_delay_steps) = 0;
    IS(I_reset_buf) = 1; IS(I_time_out_buf) = 0; IS(I_nan_resets) = 0;
    IS(I_episode_steps) = -1;   // the reference Recorder's step counter is 0 (not 1) after the first step() (utils/recorder.py:37-40)
}

}  // namespace b200
