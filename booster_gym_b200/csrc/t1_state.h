// t1_state.h - structure-of-arrays layout of the per-environment state: float fstate[F_ROWS][N], int32 istate[I_ROWS][N].
// One row per scalar so that every warp access is a coalesced 128-byte line.  Mirrors the tensors created by
// envs/t1.py:_init_buffers (:187-272) and _create_envs (:69-83,122) - SURVEY Appendix C - plus the per-env model
// parameters that the reference hands to PhysX (:139-167).
#pragma once

// X(name, count)
#define B200_FLOAT_FIELDS(X)                                                                                     \
    X(root_states, 13)      /* pos 3, quat xyzw 4, lin vel world 3, ang vel world 3 (envs/t1.py:215) */           \
    X(dof_pos, 12)                                                                                               \
    X(dof_vel, 12)                                                                                               \
    X(actions, 12)                                                                                               \
    X(last_actions, 12)                                                                                          \
    X(last_dof_vel, 12)                                                                                          \
    X(last_root_vel, 6)                                                                                          \
    X(last_dof_targets, 12)                                                                                      \
    X(torques, 12)          /* mean over the decimation loop (envs/t1.py:456) */                                  \
    X(commands, 3)                                                                                               \
    X(gait_frequency, 1)                                                                                         \
    X(gait_process, 1)                                                                                           \
    X(base_lin_vel, 3)                                                                                           \
    X(base_ang_vel, 3)                                                                                           \
    X(projected_gravity, 3)                                                                                      \
    X(filtered_lin_vel, 3)                                                                                       \
    X(filtered_ang_vel, 3)                                                                                       \
    X(pushing_forces, 3)    /* on the Trunk, local frame (envs/t1.py:509) */                                      \
    X(pushing_torques, 3)                                                                                        \
    X(feet_pos, 6)                                                                                               \
    X(feet_quat, 8)                                                                                              \
    X(last_feet_pos, 6)                                                                                          \
    X(feet_roll, 2)                                                                                              \
    X(feet_yaw, 2)                                                                                               \
    X(feet_force, 2)        /* normal force estimate per foot (net_contact_force z of the foot links) */          \
    X(episode_sums, 27)     /* device episode statistics: [0] reward, [1+k] term k (utils/recorder.py:36-62) */    \
    X(dof_stiffness, 12)                                                                                         \
    X(dof_damping, 12)                                                                                           \
    X(dof_friction, 12)                                                                                          \
    X(base_mass_scaled, 4)  /* RAW rand samples (utils/utils.py:27-28 quirk) */                                   \
    X(body_mass, 13)                                                                                             \
    X(body_com, 39)                                                                                              \
    X(foot_friction, 2)     /* combined Coulomb coefficient foot/ground */                                        \
    X(foot_kscale, 2)                                                                                            \
    X(foot_cscale, 2)                                                                                            \
    X(env_origins, 3)

#define B200_INT_FIELDS(X)                                                                                       \
    X(episode_length_buf, 1)                                                                                     \
    X(cmd_resample_time, 1)                                                                                      \
    X(delay_steps, 1)                                                                                            \
    X(feet_contact, 2)                                                                                           \
    X(reset_buf, 1)                                                                                              \
    X(time_out_buf, 1)                                                                                           \
    X(episode_steps, 1)     /* utils/recorder.py:37-43 */                                                         \
    X(nan_resets, 1)                                                                                             \
    X(env_curriculum_level, 2) /* (lin, ang) level of the current command (envs/t1.py:262) */                    \
    X(contact_mask, 2)      /* per leg lane: bit b = |net contact force on body b| > 1 N after the last substep (envs/t1.py:553,628) */

namespace b200 {

enum FloatRow {
#define X(name, count) F_##name, F_##name##_last = F_##name + (count) - 1,
    B200_FLOAT_FIELDS(X)
#undef X
    F_ROWS
};
enum IntRow {
#define X(name, count) I_##name, I_##name##_last = I_##name + (count) - 1,
    B200_INT_FIELDS(X)
#undef X
    I_ROWS
};

struct FieldInfo {
    const char* name;
    int row, count;
};
static const FieldInfo kFloatFields[] = {
#define X(name, count) {#name, F_##name, count},
    B200_FLOAT_FIELDS(X)
#undef X
};
static const FieldInfo kIntFields[] = {
#define X(name, count) {#name, I_##name, count},
    B200_INT_FIELDS(X)
#undef X
};

}  // namespace b200
