"""ctypes mirror of include/b200_t1.h (struct layouts and constants).

This is the binding a maintainer of the reference would add: the reference is pure Python, so its
"FFI" is ctypes over the C-ABI.  Layouts are checked at load time against `b200_sizeof()`.
"""
import ctypes as C

NB, NV, NQ, NU, NCON, NOBS, NPRIV, MAX_REW = 13, 18, 19, 12, 8, 47, 14, 26
NPARAMS, NPARAMS_PADDED, ACTOR_PARAMS, CRITIC_PARAMS = 177945, 177948, 63244, 114689

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4

REW_NAMES = [  # enum order of B200_REW_* (envs/t1.py:606-730)
    "survival", "tracking_lin_vel_x", "tracking_lin_vel_y", "tracking_ang_vel", "base_height", "orientation",
    "torques", "torque_tiredness", "power", "lin_vel_z", "ang_vel_xy", "dof_vel", "dof_acc", "root_acc",
    "action_rate", "dof_pos_limits", "dof_vel_limits", "torque_limits", "collision", "feet_slip", "feet_vel_z",
    "feet_yaw_diff", "feet_yaw_mean", "feet_roll", "feet_distance", "feet_swing",
]

SC = dict(LR=0, ADAM_STEP=1, VALUE_LOSS=2, ACTOR_LOSS=3, BOUND_LOSS=4, ENTROPY=5, KL=6, ADV_MEAN=7, ADV_STD=8,
          GRAD_NORM=9, SUM_VALUE_LOSS=10, SUM_ACTOR_LOSS=11, SUM_BOUND_LOSS=12, SUM_ENTROPY=13, EPOCHS=14,
          OLD_LOGSTD=16, COUNT=32)
DS = dict(ADV_SUM=0, ADV_SUMSQ=1, ADV_COUNT=2, VALUE_LOSS=4, ACTOR_LOSS=5, BOUND_LOSS=6, ENTROPY=7, KL=8, SAMPLES=9,
          GRAD_SQ=10, DLOGSTD=16, COUNT=32)


def _model_fields(real):
    return [
        ("body_pos", real * 3 * NB), ("ipos", real * 3 * NB), ("inertia", real * 6 * NB), ("mass", real * NB),
        ("jnt_lower", real * NU), ("jnt_upper", real * NU), ("dof_inertia", real * NU),
        ("foot_corner", real * 3 * 4), ("gravity", real), ("dt", real), ("contact_k", real), ("contact_c", real),
        ("stiction_vel", real), ("limit_k", real), ("limit_c", real),
        ("trunk_box_pos", real * 3), ("trunk_box_half", real * 3), ("trunk_box_radius", real), ("cyl_pos", real * 3 * 2), ("cyl_radius", real * 2),
        ("cyl_half", real * 2), ("body_mu", real), ("self_k", real), ("self_c", real), ("foot_cap", real * 3 * 2),
        ("foot_cap_radius", real),
        ("axis", C.c_int32 * NB), ("enable_contact", C.c_int32), ("enable_limits", C.c_int32),
        ("enable_body_contact", C.c_int32), ("enable_self_contact", C.c_int32), ("pad0", C.c_int32),
    ]


class ModelF(C.Structure):
    _fields_ = _model_fields(C.c_float)


class ModelD(C.Structure):
    _fields_ = _model_fields(C.c_double)


class Rand(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("dist", C.c_int32), ("op", C.c_int32), ("a", C.c_float), ("b", C.c_float)]


class T1Config(C.Structure):
    _fields_ = [
        ("env_dt", C.c_float), ("decimation", C.c_int32), ("action_scale", C.c_float), ("clip_actions", C.c_float),
        ("default_dof_pos", C.c_float * NU), ("kp_nominal", C.c_float * NU), ("kd_nominal", C.c_float * NU),
        ("init_root", C.c_float * 13), ("env_spacing", C.c_float),
        ("norm_gravity", C.c_float), ("norm_lin_vel", C.c_float), ("norm_ang_vel", C.c_float),
        ("norm_dof_pos", C.c_float), ("norm_dof_vel", C.c_float), ("filter_weight", C.c_float),
        ("norm_push_force", C.c_float), ("norm_push_torque", C.c_float),
        ("noise_gravity", Rand), ("noise_lin_vel", Rand), ("noise_ang_vel", Rand), ("noise_dof_pos", Rand),
        ("noise_dof_vel", Rand), ("noise_height", Rand),
        ("init_dof_pos", Rand), ("init_base_pos_xy", Rand), ("init_base_lin_vel_xy", Rand), ("kick_lin_vel", Rand),
        ("kick_ang_vel", Rand), ("push_force", Rand), ("push_torque", Rand), ("dof_stiffness", Rand),
        ("dof_damping", Rand), ("dof_friction", Rand), ("friction", Rand), ("compliance", Rand),
        ("restitution", Rand), ("base_com", Rand), ("base_mass", Rand), ("other_com", Rand), ("other_mass", Rand),
        ("kick_interval", C.c_int32), ("push_interval", C.c_int32), ("push_duration", C.c_int32),
        ("push_all_substeps", C.c_int32),
        ("lin_vel_x", C.c_float * 2), ("lin_vel_y", C.c_float * 2), ("ang_vel_yaw", C.c_float * 2),
        ("gait_frequency", C.c_float * 2), ("still_proportion", C.c_float),
        ("resample_lo", C.c_int32), ("resample_hi", C.c_int32), ("curriculum", C.c_int32),
        ("cur_lin_levels", C.c_int32), ("cur_ang_levels", C.c_int32), ("cur_update_rate", C.c_float),
        ("cur_res_x", C.c_float), ("cur_res_y", C.c_float), ("cur_res_ang", C.c_float), ("cur_success_len", C.c_float),
        ("cur_tol_x", C.c_float), ("cur_tol_y", C.c_float), ("cur_tol_yaw", C.c_float),
        ("n_rew", C.c_int32), ("rew_id", C.c_int32 * MAX_REW), ("rew_scale", C.c_float * MAX_REW),
        ("max_episode_length", C.c_int32), ("terminate_height", C.c_float), ("terminate_vel", C.c_float),
        ("only_positive_rewards", C.c_int32),
        ("tracking_sigma", C.c_float), ("base_height_target", C.c_float), ("soft_dof_vel_limit", C.c_float),
        ("soft_torque_limit", C.c_float), ("swing_period", C.c_float), ("feet_distance_ref", C.c_float),
        ("dof_pos_soft_lower", C.c_float * NU), ("dof_pos_soft_upper", C.c_float * NU),
        ("dof_vel_limits", C.c_float * NU), ("torque_limits", C.c_float * NU),
        ("penalized_body_mask", C.c_int32), ("termination_body_mask", C.c_int32),
        ("terrain_type", C.c_int32), ("border_pixels", C.c_int32),
        ("horizontal_scale", C.c_float), ("env_width", C.c_float),
        ("env_length", C.c_float), ("border_size", C.c_float), ("terrain_friction", C.c_float), ("pad1", C.c_int32),
        ("pad2", C.c_int32), ("vertical_scale", C.c_double),
    ]


class PpoConfig(C.Structure):
    _fields_ = [
        ("gamma", C.c_double), ("lam", C.c_double),
        ("e_clip", C.c_float), ("bound_coef", C.c_float),
        ("entropy_coef", C.c_float), ("desired_kl", C.c_float), ("max_grad_norm", C.c_float),
        ("lr_min", C.c_float), ("lr_max", C.c_float), ("lr_factor", C.c_float),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("horizon", C.c_int32), ("num_envs", C.c_int32), ("world_size", C.c_int32), ("env_base", C.c_int32),
        ("pad0", C.c_int32),
    ]
