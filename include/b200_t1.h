/*
 * b200_t1.h - C-ABI of the B200-native T1 training hot path (libb200t1.so).
 *
 * Every entry point replaces one reference (booster_gym) call site; the reference is pure Python, so the
 * "FFI" a maintainer binds is ctypes.  Plain pointers and sizes only, no torch types.
 * All device pointers are caller-owned (torch allocates); the library keeps only the constants uploaded in
 * b200_t1_create() / b200_ppo_create().  All launches are stream-ordered on the cudaStream_t passed as
 * `void* stream` (0 = legacy default stream); no call synchronises the host, allocates, or frees except
 * create/destroy (and b200_t1_episode_stats, which is an explicit read-back).
 *
 * Return value: 0 on success, negative B200_ERR_* otherwise (never throws).
 */
#ifndef B200_T1_H
#define B200_T1_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_NB 13   /* bodies: Trunk, 6 left-leg, 6 right-leg   (resources/T1/T1_locomotion.xml:37-119) */
#define B200_NV 18   /* 6 free-joint + 12 hinge DoF */
#define B200_NQ 19
#define B200_NU 12
#define B200_NCON 8  /* sole-corner contact points, 4 per foot (envs/T1.yaml:79-82); the other collision shapes: model fields below */
#define B200_NOBS 47
#define B200_NPRIV 14
#define B200_MAX_REW 26

#define B200_OK 0
#define B200_ERR_ARG (-1)      /* bad argument / config (reference raises ValueError) */
#define B200_ERR_CUDA (-2)     /* a CUDA runtime call failed; see b200_last_error() */
#define B200_ERR_STATE (-3)    /* state not bound */
#define B200_ERR_UNSUPPORTED (-4)

/* ---- rigid-body model constants (from booster_gym_b200/assets/t1_model.json) -------------------------------
 * Tree: body 0 = Trunk (free joint), leg s in {0 left, 1 right}, link k in 0..5 -> body 1+6s+k, DoF 6+6s+k.
 * Hinge axes are coordinate axes and body frames carry no rotation (resources/T1/T1_locomotion.xml:54-119).
 * Generalised velocity follows MuJoCo: [v_world(3), omega_body(3), qd(12)]. */
#define B200_MODEL_FIELDS(REAL)                                                                              \
    REAL body_pos[B200_NB][3];   /* child offset in parent frame */                                        \
    REAL ipos[B200_NB][3];       /* CoM in body frame */                                                    \
    REAL inertia[B200_NB][6];    /* body-frame inertia about CoM: xx yy zz xy xz yz */                      \
    REAL mass[B200_NB];                                                                                      \
    REAL jnt_lower[B200_NU];     /* hinge limits enforced by the physics (MJCF range == URDF limits) */      \
    REAL jnt_upper[B200_NU];                                                                                 \
    REAL dof_inertia[B200_NU];   /* 1 / dof_invweight0: effective inertia used to scale the limit spring */  \
    REAL foot_corner[4][3];      /* sole corners in the foot frame (envs/T1.yaml:79-82) */                  \
    REAL gravity;                /* 9.81, along -z */                                                       \
    REAL dt;                     /* physics tick, 0.002 */                                                   \
    REAL contact_k;              /* normal spring per corner [N/m]   = solref K * contact_mass */           \
    REAL contact_c;              /* normal damper per corner [N s/m] = solref B * contact_mass */           \
    REAL stiction_vel;           /* regularisation velocity of the Coulomb law [m/s] */                      \
    REAL limit_k, limit_c;       /* joint-limit spring/damper per unit effective inertia [1/s^2, 1/s] */     \
    /* collision primitives other than the soles (resources/T1/T1_locomotion.xml:42,66,71,99,104) */ \
    REAL trunk_box_pos[3];       /* box centre in the trunk frame */                                         \
    REAL trunk_box_half[3];                                                                                  \
    REAL trunk_box_radius;       /* |half|: bounding sphere used to cull the box against the highest terrain sample */ \
    REAL cyl_pos[2][3];          /* cylinder centre in its body frame: [0] Hip_Yaw (leg link 2), [1] Shank (3) */ \
    REAL cyl_radius[2];          /* the cylinder axis is the body z axis */                                  \
    REAL cyl_half[2];                                                                                        \
    REAL body_mu;                /* combined Coulomb coefficient of those shapes against the ground */       \
    REAL self_k, self_c;         /* leg-leg penalty spring / damper [N/m, N s/m] */                          \
    REAL foot_cap[2][3];         /* capsule of the foot box for leg-leg contact: segment end points, foot frame */ \
    REAL foot_cap_radius;                                                                                    \
    int32_t axis[B200_NB];       /* hinge axis 0/1/2, -1 for the free joint */                              \
    int32_t enable_contact;      /* 0 = contact-free dynamics (BASELINE config 5 i) */                      \
    int32_t enable_limits;                                                                                   \
    int32_t enable_body_contact; /* 1: trunk box / hip-yaw / shank cylinders against the ground */           \
    int32_t enable_self_contact; /* 1: leg-leg capsule contacts (asset.self_collisions: 0 = enabled) */      \
    int32_t pad0;

typedef struct B200T1ModelF { B200_MODEL_FIELDS(float) } B200T1ModelF;
typedef struct B200T1ModelD { B200_MODEL_FIELDS(double) } B200T1ModelD;

/* ---- apply_randomization() descriptor (utils/utils.py:5-30) ---------------------------------------------- */
typedef struct B200Rand {
    int32_t enabled;  /* 0: params == None -> identity */
    int32_t dist;     /* 0 gaussian (mu + "var"*randn; "var" is used as a std), 1 uniform (lo + (hi-lo)*rand) */
    int32_t op;       /* 0 additive, 1 scaling */
    float a, b;       /* range[0], range[1] */
} B200Rand;

/* reward term ids (envs/t1.py:606-730), independent of YAML order */
enum {
    B200_REW_SURVIVAL = 0, B200_REW_TRACK_LIN_X, B200_REW_TRACK_LIN_Y, B200_REW_TRACK_ANG, B200_REW_BASE_HEIGHT,
    B200_REW_ORIENTATION, B200_REW_TORQUES, B200_REW_TORQUE_TIREDNESS, B200_REW_POWER, B200_REW_LIN_VEL_Z,
    B200_REW_ANG_VEL_XY, B200_REW_DOF_VEL, B200_REW_DOF_ACC, B200_REW_ROOT_ACC, B200_REW_ACTION_RATE,
    B200_REW_DOF_POS_LIMITS, B200_REW_DOF_VEL_LIMITS, B200_REW_TORQUE_LIMITS, B200_REW_COLLISION,
    B200_REW_FEET_SLIP, B200_REW_FEET_VEL_Z, B200_REW_FEET_YAW_DIFF, B200_REW_FEET_YAW_MEAN, B200_REW_FEET_ROLL,
    B200_REW_FEET_DISTANCE, B200_REW_FEET_SWING, B200_REW_COUNT
};

/* ---- everything T1.step()/reset() reads from envs/T1.yaml -------------------------------------------------- */
typedef struct B200T1Config {
    /* control / sim */
    float env_dt;            /* decimation * sim dt (envs/t1.py:191) */
    int32_t decimation;
    float action_scale, clip_actions;
    float default_dof_pos[B200_NU];
    float kp_nominal[B200_NU], kd_nominal[B200_NU];
    float init_root[13];     /* pos, quat xyzw, lin vel, ang vel (envs/t1.py:110-113) */
    float env_spacing;
    /* normalization (envs/T1.yaml:136-145) */
    float norm_gravity, norm_lin_vel, norm_ang_vel, norm_dof_pos, norm_dof_vel, filter_weight, norm_push_force,
        norm_push_torque;
    /* observation noise (envs/T1.yaml:147-171) */
    B200Rand noise_gravity, noise_lin_vel, noise_ang_vel, noise_dof_pos, noise_dof_vel, noise_height;
    /* randomization (envs/T1.yaml:173-249) */
    B200Rand init_dof_pos, init_base_pos_xy, init_base_lin_vel_xy, kick_lin_vel, kick_ang_vel, push_force,
        push_torque, dof_stiffness, dof_damping, dof_friction, friction, compliance, restitution, base_com,
        base_mass, other_com, other_mass;
    int32_t kick_interval, push_interval, push_duration; /* in env steps: ceil(s / dt) */
    int32_t push_all_substeps; /* 0 (default, the reference): the push registered by gym.apply_rigid_body_force_tensors once per
                                * step() (envs/t1.py:522-527) acts on the NEXT simulate() only, i.e. on the first of the
                                * `decimation` substeps; 1: on every substep (extension key randomization.push_all_substeps) */
    /* commands (envs/T1.yaml:112-134) */
    float lin_vel_x[2], lin_vel_y[2], ang_vel_yaw[2], gait_frequency[2];
    float still_proportion;
    int32_t resample_lo, resample_hi; /* int(s/dt) */
    int32_t curriculum;               /* envs/T1.yaml:124: command curriculum (envs/t1.py:391-435) */
    int32_t cur_lin_levels, cur_ang_levels;      /* grid = [2 lin + 1][2 ang + 1] */
    float cur_update_rate;
    float cur_res_x, cur_res_y, cur_res_ang;     /* lin_vel_x_resolution, lin_vel_y_resolution, ang_vel_resolution */
    float cur_success_len;                       /* ceil(episode_length_s / dt) * (1 - episode_length_toler) */
    float cur_tol_x, cur_tol_y, cur_tol_yaw;
    /* rewards (envs/T1.yaml:251-291) */
    int32_t n_rew;                    /* number of non-zero scales, YAML order */
    int32_t rew_id[B200_MAX_REW];     /* B200_REW_* per active term */
    float rew_scale[B200_MAX_REW];    /* scale * dt (envs/t1.py:285) */
    int32_t max_episode_length;       /* ceil(episode_length_s / dt) */
    float terminate_height, terminate_vel;
    int32_t only_positive_rewards;
    float tracking_sigma, base_height_target, soft_dof_vel_limit, soft_torque_limit, swing_period,
        feet_distance_ref;
    float dof_pos_soft_lower[B200_NU], dof_pos_soft_upper[B200_NU]; /* envs/t1.py:665-670 precomputed */
    float dof_vel_limits[B200_NU], torque_limits[B200_NU];
    int32_t penalized_body_mask, termination_body_mask; /* bit b = body b */
    /* terrain (utils/terrain.py:30-45) */
    int32_t terrain_type;             /* 0 plane, 1 trimesh (heightfield) */
    int32_t border_pixels;
    float horizontal_scale, env_width, env_length, border_size;
    float terrain_friction;           /* static friction of the ground (envs/T1.yaml:99) */
    int32_t pad1, pad2;
    double vertical_scale;            /* fp64: the reference multiplies the fp64 interpolant by the Python float */
} B200T1Config;

typedef struct B200T1Handle B200T1Handle;

const char* b200_last_error(void);
int b200_version(void);
int b200_sizeof(int what); /* 0 B200T1ModelF, 1 B200T1Config, 2 B200PpoConfig, 3 B200T1ModelD: ABI self-check */

/* State is two caller-owned device arrays in structure-of-arrays form: float fstate[NF][N], int32 istate[NI][N].
 * Field tables: kind 0 = float fields, 1 = int fields; b200_t1_field_info(kind, idx, &name, &row, &count). */
int b200_t1_num_float_rows(void);
int b200_t1_num_int_rows(void);
int b200_t1_num_fields(int kind);
int b200_t1_field_info(int kind, int idx, const char** name, int* row, int* count);

/* replaces T1.__init__ (envs/t1.py:26-31): uploads model, config and the int16 heightfield (may be NULL for plane) */
int b200_t1_create(const B200T1ModelF* model, const B200T1Config* cfg, const int16_t* hf_host, int hf_rows,
                   int hf_cols, int num_envs, int device, uint64_t seed, B200T1Handle** out);
int b200_t1_destroy(B200T1Handle* h);
int b200_t1_bind_state(B200T1Handle* h, float* fstate, int32_t* istate);
int b200_t1_num_envs(const B200T1Handle* h);
/* replaces _create_envs per-env DR + _get_env_origins + _init_buffers (envs/t1.py:69-83,123-272); env_index_base =
 * first global env index of this rank's shard, total_envs = global env count (origins use the global layout) */
int b200_t1_init_params(B200T1Handle* h, int env_index_base, int total_envs, void* stream);
/* replaces T1.reset() (envs/t1.py:294-299): resets every env, resamples commands, writes obs/priv */
int b200_t1_reset(B200T1Handle* h, float* obs, float* priv, void* stream);
/* replaces T1.step() (envs/t1.py:437-497).  actions [N,12]; outputs obs [N,47], priv [N,14], rew [N],
 * done [N] (u8), time_out [N] (u8), rew_terms [n_rew][N] (may be NULL). common_step is the value of
 * common_step_counter AFTER the increment at envs/t1.py:477. */
int b200_t1_step(B200T1Handle* h, const float* actions, float* obs, float* priv, float* rew, uint8_t* done,
                 uint8_t* time_out, float* rew_terms, int64_t common_step, void* stream);
/* the two halves of step(), exposed for parity tests and the physics-only sweep (BASELINE config 5):
 * physics = envs/t1.py:439-456 (+ what gym.simulate does), post = envs/t1.py:460-497.
 * apply_pd = 0: `actions` holds raw joint torques [N,12] applied every substep (play_mujoco-style / qacc tests).
 * qacc_out (nullable, device [18][N]) receives the generalised acceleration of the LAST substep. */
int b200_t1_physics(B200T1Handle* h, const float* actions, int n_substeps, int apply_pd, float* qacc_out,
                    void* stream);
int b200_t1_post_physics(B200T1Handle* h, float* obs, float* priv, float* rew, uint8_t* done, uint8_t* time_out,
                         float* rew_terms, int64_t common_step, int noise_on, void* stream);
/* device-side episode statistics (replaces utils/recorder.py:36-62): sums[n_rew+2] =
 * {sum of finished-episode reward, per-term sums..., sum of steps}, count = finished episodes; read-and-clear.
 * This call synchronises the stream. */
int b200_t1_episode_stats(B200T1Handle* h, double* sums_host, int64_t* count_host, void* stream);
/* the same read-and-clear WITHOUT the synchronisation: out_pinned (page-locked host memory, n_rew + 3 doubles) receives
 * {reward sum, per-term sums..., step sum, episode count} in stream order; the caller waits on an event of its own before reading
 * (Runner.train logs iteration i while the graphs of iteration i + 1 run). */
int b200_t1_episode_stats_async(B200T1Handle* h, double* out_pinned, void* stream);

/* replaces Terrain.terrain_heights (utils/terrain.py:101-121) on the device heightfield; xy rows of `stride` floats */
int b200_terrain_heights(const B200T1Handle* h, const float* xy, int stride, int count, float* out, void* stream);

/* counter-based RNG probe: the exact samples the kernels draw for (step, purpose, sub) of every env, so the CPU
 * oracle can be fed identical noise. kind 0 = uniform [0,1), 1 = standard normal, 2 = raw 32-bit words.
 * out is DEVICE float/uint32 [4][N]. */
int b200_rng_fill(const B200T1Handle* h, uint64_t step, int purpose, int sub, int kind, float* out, void* stream);

/* Parity-test hook: table = DEVICE uint32 [b200_t1_rng_slots()][12][N] (per slot: 4 raw words, 4 uniforms and 4 normals
 * as float bits); while set, reset()/step() READ their random draws from it instead of generating them, so kernel,
 * CPU oracle and the reference (torch.randn_like & co. patched) see identical samples. NULL restores Philox.
 * Slots: 0-2 reset dof noise, 3-4 reset root, 5 delay, 6-7 command, 8-9 kick, 10-11 push, 12-20 observation noise. */
/* Command curriculum (`commands.curriculum: true`): `prob` is the caller-owned device grid float[(2 lin + 1) * (2 ang + 1)]
 * (env.curriculum_prob, row-major [lin, ang]); resets of successful episodes raise it, command resampling draws from it.
 * Must be bound before the first reset/step when the curriculum is enabled. */
int b200_t1_bind_curriculum(B200T1Handle* h, float* prob, int rows, int cols);
int b200_t1_inject_rng(B200T1Handle* h, const uint32_t* table);
int b200_t1_rng_slots(void);

/* current (rng step, common_step_counter) of the handle; synchronises the stream. Used by parity tests to ask
 * b200_rng_fill() for the samples a given reset()/step() call drew. */
int b200_t1_counters(B200T1Handle* h, int64_t* rng_step, int64_t* common_step, void* stream);

/* ---- learner: policy inference, GAE, PPO epoch (utils/model.py, utils/utils.py, utils/runner.py:123-185) ---- */
#define B200_ACTOR_PARAMS 63244
#define B200_CRITIC_PARAMS 114689
#define B200_NPARAMS 177945        /* trainable scalars of ActorCritic(12, 47, 14) (utils/model.py:9-27) */
#define B200_NPARAMS_PADDED 177948 /* length of the flat param/grad/Adam buffers: every tensor starts 16-byte aligned */

typedef struct B200PpoConfig {
    double gamma, lam;  /* kept in fp64 like the Python floats they replace (gamma*lam is formed in fp64, utils/utils.py:43) */
    float e_clip, bound_coef, entropy_coef, desired_kl, max_grad_norm;
    float lr_min, lr_max, lr_factor;
    float adam_beta1, adam_beta2, adam_eps;
    int32_t horizon, num_envs, world_size;
    int32_t env_base; /* global index of env 0 of this rank's shard (keys the action-sampling RNG) */
    int32_t pad0;
} B200PpoConfig;

typedef struct B200Ppo B200Ppo;

/* flat parameter layout: tensor idx -> state_dict name, offset (floats) into the flat buffers, shape */
int b200_ppo_num_params(void);        /* B200_NPARAMS_PADDED */
int b200_ppo_num_param_tensors(void); /* 17 */
int b200_ppo_param_info(int idx, const char** name, int* offset, int* rows, int* cols);

/* workspace_bytes: caller allocates one device scratch buffer of that size (256-byte aligned) and passes it to create */
int64_t b200_ppo_workspace_bytes(int horizon, int num_envs);
/* params/grads/adam_m/adam_v: device float[B200_NPARAMS_PADDED]; scalars: device float[32] (B200_SC_*; the caller
 * writes B200_SC_LR before the first epoch); dstats: device double[32] (B200_DS_*: reduction targets, the buffers a
 * multi-GPU caller all-reduces). */
int b200_ppo_create(const B200PpoConfig* cfg, float* params, float* grads, float* adam_m, float* adam_v, float* scalars,
                    double* dstats, void* workspace, int device, B200Ppo** out);
int b200_ppo_destroy(B200Ppo* p);

enum { /* rows of the device `scalars` array */
    B200_SC_LR = 0, B200_SC_ADAM_STEP, B200_SC_VALUE_LOSS, B200_SC_ACTOR_LOSS, B200_SC_BOUND_LOSS, B200_SC_ENTROPY,
    B200_SC_KL, B200_SC_ADV_MEAN, B200_SC_ADV_STD, B200_SC_GRAD_NORM,
    /* running sums over the epochs since the caller last zeroed them (utils/runner.py:182-189) */
    B200_SC_SUM_VALUE_LOSS = 10, B200_SC_SUM_ACTOR_LOSS, B200_SC_SUM_BOUND_LOSS, B200_SC_SUM_ENTROPY, B200_SC_EPOCHS,
    B200_SC_OLD_LOGSTD = 16, /* 12 entries: logstd at old_dist time (utils/runner.py:123-125) */
    B200_SC_COUNT = 32
};
enum { /* rows of the device `dstats` array (sums; cleared by b200_ppo_epoch_a) */
    B200_DS_ADV_SUM = 0, B200_DS_ADV_SUMSQ, B200_DS_ADV_COUNT, B200_DS_PAD, B200_DS_VALUE_LOSS, B200_DS_ACTOR_LOSS,
    B200_DS_BOUND_LOSS, B200_DS_ENTROPY, B200_DS_KL, B200_DS_SAMPLES, B200_DS_GRAD_SQ, B200_DS_DLOGSTD = 16,
    B200_DS_COUNT = 32
};

/* replaces model.act(obs).sample() (utils/runner.py:109-111): mu = actor(obs); act = mu + exp(logstd)*eps.
 * eps drawn in-kernel (Philox, keyed by seed/env/step) unless eps_in != NULL. deterministic != 0 -> act = mu (play).
 * step = B200_STEP_AUTO: the RNG step is a device-side counter incremented by every such call (CUDA-graph replayable). */
#define B200_STEP_AUTO 0xFFFFFFFFFFFFFFFFull
/* `deterministic` is a flag word: bit 0 = act = mu; bit 1 (B200_ACT_REUSE_WEIGHTS) = the parameters have not changed since this
 * caller's previous b200_policy_act call, so the tensor-core path (>= 2048 rows: hidden layers on the fused tcgen05 forward chain)
 * may reuse the weight operands it prepared then.  Without the bit every call rebuilds them (always correct). */
#define B200_ACT_REUSE_WEIGHTS 2
int b200_policy_act(B200Ppo* p, const float* obs, int n, float* actions, float* mu_out /*nullable*/,
                    const float* eps_in /*nullable*/, uint64_t seed, uint64_t step, int deterministic, void* stream);
/* replaces est_value (utils/model.py:34-36) */
int b200_critic_value(B200Ppo* p, const float* obs, const float* priv, int n, float* values, void* stream);
/* replaces utils/runner.py:123-125: old mu [M,12] and old log-prob [M] of the stored actions. Also stages the rollout
 * observations (obses [T,N,47], privs [T,N,14]) as the padded GEMM operands every following epoch reads. */
int b200_ppo_old_dist(B200Ppo* p, const float* obses, const float* privs, const float* actions, float* old_mu,
                      float* old_logp, void* stream);
/* replaces utils/utils.py:33-44 + utils/runner.py:135,144 (time-out bootstrap in place, GAE, returns) and the
 * advantage moments of :145. rewards is modified in place like the reference. stats (nullable; device double[>=3])
 * accumulates sum, sum of squares and count of the raw advantages. */
int b200_gae(float* rewards, const uint8_t* dones, const uint8_t* time_outs, const float* values,
             const float* last_values, double gamma, double lam, float* advantages, float* returns, double* stats,
             int horizon, int num_envs, void* stream);
/* one full-batch epoch body, utils/runner.py:132-161 + backward: fills grads (flat, mean over local samples) and
 * the loss sums; the optional allreduce between this and b200_ppo_apply is done by the caller (NCCL) */
int b200_ppo_epoch(B200Ppo* p, const float* actions, float* rewards, const uint8_t* dones, const uint8_t* time_outs,
                   const float* last_obs, const float* last_priv, const float* old_mu, const float* old_logp,
                   void* stream);
/* the two halves of b200_ppo_epoch for multi-GPU runs: stage A = critic forward, time-out bootstrap, GAE and the
 * advantage moments (dstats[0..2], all-reduced by the caller); stage B = actor forward, losses, backward */
/* last_obs / last_priv (the post-rollout observations, the same tensors in every mini-epoch of the reference, utils/runner.py:132) are
 * staged by the first b200_ppo_epoch_a after b200_ppo_old_dist - or whenever the pointers change - and reused by the following epochs:
 * a caller that rewrites them in place between two epochs of one update must call b200_ppo_old_dist again. */
int b200_ppo_epoch_a(B200Ppo* p, float* rewards, const uint8_t* dones, const uint8_t* time_outs, const float* last_obs,
                     const float* last_priv, void* stream);
int b200_ppo_epoch_b(B200Ppo* p, const float* actions, const float* old_mu, const float* old_logp, void* stream);
/* views into the workspace for parity tests (device pointers, valid after an epoch): 0 values [M], 1 advantages
 * (raw) [M], 2 returns [M], 3 mu [M,12], 4 last_values [N], 5 dL/dV [M], 6 dL/dmu [M,12] */
float* b200_ppo_buffer(B200Ppo* p, int which);
/* replaces utils/runner.py:162-180: clip_grad_norm_(1.0), Adam step, KL-adaptive learning rate (all on device) */
int b200_ppo_apply(B200Ppo* p, void* stream);
/* Multi-GPU exchange over NVLink peer memory instead of three NCCL all-reduces per epoch.  Every rank allocates
 * one SYMMETRIC device buffer of b200_ppo_peer_buffer_bytes() bytes, zero-filled, peer-mapped into all ranks of the node
 * (torch.distributed._symmetric_memory.empty + rendezvous), and passes the device addresses of all ranks' buffers, indexed by
 * rank.  After binding, b200_ppo_epoch_a posts this rank's advantage moments, b200_ppo_epoch_b sums all ranks' moments
 * before the losses, and b200_ppo_apply sums gradients + loss sums over the peers (fused with the gradient norm): the caller
 * must NOT all-reduce dstats / grads itself any more.  All ranks must call the epoch functions in lockstep. */
long long b200_ppo_peer_buffer_bytes(void);
int b200_ppo_bind_peers(B200Ppo* p, const unsigned long long* buffer_ptrs, int rank, int world);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------------------
 * b200_launch_count: kernels (and memset nodes) this library has launched in the calling process so far.
 * b200_profile_gemm(1): bracket every following tensor-core GEMM launch with CUDA events on its stream (up to 8192
 * launches); b200_profile_gemm_read() synchronises them and returns summed duration, algorithmic FLOPs
 * (2 * rows * out * reduction, unpadded, counted once - not 3x for the split) and the number of launches. */
long long b200_launch_count(void);
int b200_profile_gemm(int enable);
/* kind: 0 = k_gemm3x (mma.sync), 1 = k_tc_rowmajor (tcgen05 forward / dgrad), 2 = k_tc_wgrad (tcgen05), -1 = all */
int b200_profile_gemm_read(int kind, double* total_ms, double* total_flops, int* launches);
/* algorithmic HBM bytes of the same launches: every fp32 operand read once, every result written once (weights excluded:
 * they stay in L2) - the numerator of the HBM roofline of the GEMM families */
int b200_profile_gemm_bytes(int kind, double* total_bytes);
/* k_tc_rowmajor variant for the critic forward and the dgrad GEMMs: 0 (default) = one CTA per 128-row tile, 1 = CTA pairs
 * (clusters of 2, tcgen05 cta_group::2, 256-row tiles, each CTA stages half of the weight tile).  Same results; the pair
 * variant halves the per-CTA L2 -> shared-memory weight traffic but measured ~4 % slower on this shape. */
int b200_tc_set_pair(int enable);
/* measured FP32 FMA throughput of the current device in TFLOP/s (FMA = 2 FLOP): 8 independent chains per thread, 8 CTAs of 256
 * threads per SM, best of 5 launches timed with CUDA events on `stream`.  Roofline denominator of k_physics. */
int b200_fma_peak(double* tflops, void* stream);
/* hidden layers of the PPO epoch: 1 (default) = fused layer chains (mlp_chain.cuh: one persistent tcgen05 kernel per direction,
 * activations handed from layer to layer in TMEM), 2 = the same chains on CTA pairs (clusters of 2, tcgen05 cta_group::2, 256-row
 * tiles, each CTA stages half of every weight k-block), 0 = one GEMM launch per layer (k_tc_rowmajor).  Same results within the
 * stated tolerances; replaces the autograd graph of utils/runner.py:132-133,148,163. */
int b200_tc_set_chain(int enable);
/* operand format of the hidden-layer GEMMs of the PPO epoch: bit 0 = 1 (default) "h2 words" (h2.cuh: two fp16 halves per 32-bit word,
 * power-of-two scales, tcgen05.mma kind::f16 - two MMAs per k-step at twice the kind::tf32 rate, fp32-class products; forward and
 * input-gradient chains in mlp_chain_h2.cuh, all six weight gradients in one k_wgrad_h2 launch), 0 = the 3xTF32 kernels.  Bits 4-5 /
 * 6-7: epilogue warp groups (1 or 2) of the forward / backward chain (tuning).  Same results within the stated tolerances. */
int b200_tc_set_h2(int mode);

#ifdef __cplusplus
}
#endif
#endif /* B200_T1_H */
