'''ctypes mirror of include/scg.h and loader of the in-tree CUDA library.

This is the binding a maintainer of the reference would add: plain pointers and
sizes, no torch types in the signatures.  The library is REQUIRED: there is no CPU fallback, `lib()`
raises if `libscg_b200.so` is missing or cannot be loaded.
'''

import ctypes as C
import os

SCG_ABI_VERSION = 4
SCG_MAX_PEERS = 8
SCG_MAX_SAFETY = 16
SCG_MAX_NX = 12
SCG_MAX_NU = 4
SCG_MAX_OBS = 48
SCG_MAX_INERTIAL = 4
SCG_MAX_DIST = 4
SCG_MAX_DIST_DIM = 12
SCG_MAX_CON_ROWS = 96
SCG_MAX_CON_NNZ = 512
SCG_MAX_ADV = 4

SCG_OK, SCG_EINVAL, SCG_ECUDA, SCG_ESTATE, SCG_ENOMEM = 0, -1, -2, -3, -4

SCG_SYS_CARTPOLE, SCG_SYS_QUAD1D, SCG_SYS_QUAD2D, SCG_SYS_QUAD3D = 0, 1, 2, 3
SCG_TASK_STABILIZATION, SCG_TASK_TRAJ_TRACKING = 0, 1
SCG_COST_RL_REWARD, SCG_COST_QUADRATIC = 0, 1
SCG_PREC_F64, SCG_PREC_F32 = 0, 1
SCG_RNG_TABLE, SCG_RNG_PHILOX = 0, 1
SCG_DIST_IMPULSE, SCG_DIST_STEP, SCG_DIST_UNIFORM, SCG_DIST_WHITE, SCG_DIST_PERIODIC = 0, 1, 2, 3, 4
SCG_MODE_ACTION, SCG_MODE_DYNAMICS, SCG_MODE_OBSERVATION = 0, 1, 2
SCG_RAND_NONE, SCG_RAND_UNIFORM, SCG_RAND_NORMAL = 0, 1, 2
SCG_ACT_TANH, SCG_ACT_RELU, SCG_ACT_LEAKY_RELU = 0, 1, 2
SCG_CON_LINEAR, SCG_CON_ABS, SCG_CON_QUAD = 0, 1, 2
SCG_PHYS_PYB, SCG_PHYS_PYB_GND, SCG_PHYS_PYB_DRAG, SCG_PHYS_PYB_DW, SCG_PHYS_PYB_GND_DRAG_DW = 0, 1, 2, 3, 4
PHYSICS = {'pyb': SCG_PHYS_PYB, 'pyb_gnd': SCG_PHYS_PYB_GND, 'pyb_drag': SCG_PHYS_PYB_DRAG, 'pyb_dw': SCG_PHYS_PYB_DW, 'pyb_gnd_drag_dw': SCG_PHYS_PYB_GND_DRAG_DW}
SCG_ADV_NONE, SCG_ADV_ACTION, SCG_ADV_DYNAMICS = 0, 1, 2
SCG_MLP_AUTO, SCG_MLP_FFMA2, SCG_MLP_TCGEN05, SCG_MLP_TCGEN05_TF32 = 0, 1, 2, 3
MLP_ENGINES = {'auto': SCG_MLP_AUTO, 'ffma2': SCG_MLP_FFMA2, 'tcgen05': SCG_MLP_TCGEN05, 'tcgen05_tf32': SCG_MLP_TCGEN05_TF32}

ACTIVATIONS = {'tanh': SCG_ACT_TANH, 'relu': SCG_ACT_RELU, 'leaky_relu': SCG_ACT_LEAKY_RELU}

_dbl_dd = C.c_double * SCG_MAX_DIST_DIM
_dbl_nx = C.c_double * SCG_MAX_NX
_dbl_nu = C.c_double * SCG_MAX_NU
_dbl_in = C.c_double * SCG_MAX_INERTIAL
_i32_nx = C.c_int32 * SCG_MAX_NX
_i32_in = C.c_int32 * SCG_MAX_INERTIAL


class scg_disturbance(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('dim', C.c_int32), ('has_mask', C.c_int32), ('step_offset', C.c_int32),
        ('magnitude', C.c_double), ('duration', C.c_double), ('decay_rate', C.c_double),
        ('scale', C.c_double), ('frequency', C.c_double),
        ('low', _dbl_dd), ('high', _dbl_dd), ('std', _dbl_dd), ('mask', _dbl_dd),
    ]


class scg_config(C.Structure):
    _fields_ = [
        ('struct_size', C.c_int32), ('abi_version', C.c_int32), ('system', C.c_int32), ('num_envs', C.c_int32),
        ('env_index_offset', C.c_int32), ('device', C.c_int32), ('precision', C.c_int32), ('rng_mode', C.c_int32),
        ('seed', C.c_uint64),
        ('ctrl_freq', C.c_int32), ('pyb_freq', C.c_int32), ('ctrl_steps', C.c_int32),
        ('task', C.c_int32), ('cost', C.c_int32), ('obs_goal_horizon', C.c_int32), ('obs_wrap_angle', C.c_int32),
        ('normalized_action', C.c_int32), ('rew_exponential', C.c_int32), ('done_on_out_of_bound', C.c_int32),
        ('done_on_violation', C.c_int32), ('use_constraint_penalty', C.c_int32), ('auto_reset', C.c_int32),
        ('_pad0', C.c_int32),
        ('constraint_penalty', C.c_double), ('norm_act_scale', C.c_double), ('goal_tolerance', C.c_double),
        ('rew_state_weight', _dbl_nx), ('rew_act_weight', _dbl_nu), ('mse_weight', _dbl_nx),
        ('Q', _dbl_nx), ('R', _dbl_nu), ('u_goal', _dbl_nu), ('act_low', _dbl_nu), ('act_high', _dbl_nu),
        ('hover_thrust', C.c_double), ('state_low', _dbl_nx), ('state_high', _dbl_nx),
        ('x_threshold', C.c_double), ('theta_threshold', C.c_double),
        ('gravity', C.c_double), ('kf', C.c_double), ('km', C.c_double), ('arm', C.c_double),
        ('pwm2rpm_scale', C.c_double), ('pwm2rpm_const', C.c_double), ('pwm_min', C.c_double), ('pwm_max', C.c_double),
        ('inertial_nominal', _dbl_in), ('pole_width', C.c_double),
        ('init_nominal', _dbl_nx), ('randomized_init', C.c_int32), ('randomized_inertial', C.c_int32),
        ('init_rand_kind', _i32_nx), ('inertial_rand_kind', _i32_in),
        ('init_rand_a', _dbl_nx), ('init_rand_b', _dbl_nx), ('inertial_rand_a', _dbl_in), ('inertial_rand_b', _dbl_in),
        ('n_dist', C.c_int32 * 3), ('_pad1', C.c_int32),
        ('dist', (scg_disturbance * SCG_MAX_DIST) * 3),
        ('physics', C.c_int32), ('adversary_mode', C.c_int32), ('adversary_scale', C.c_double), ('adversary_offset', C.c_double),
        ('gnd_eff_coeff', C.c_double), ('prop_radius', C.c_double), ('gnd_eff_h_clip', C.c_double),
        ('drag_coeff_xy', C.c_double), ('drag_coeff_z', C.c_double),
    ]


class scg_step_out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'd_obs', 'd_obs_f64', 'd_state_f64', 'd_rew', 'd_rew_f64', 'd_done', 'd_truncated', 'd_cviol', 'd_cvals', 'd_mse',
        'd_out_of_bounds', 'd_goal_reached', 'd_current_step', 'd_terminal_obs', 'd_terminal_obs_f64', 'd_terminal_state_f64')]


class scg_rollout_out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'd_obs', 'd_act', 'd_rew', 'd_mask', 'd_v', 'd_logp', 'd_terminal_v', 'd_ret', 'd_adv', 'd_cviol', 'd_cvals', 'd_mse',
        'd_last_obs', 'd_last_val', 'd_ep_return_sum', 'd_ep_length_sum', 'd_ep_cviol_sum', 'd_ep_count', 'd_c', 'd_obs_raw', 'd_rew_raw')]


class scg_rollout_args(C.Structure):
    _fields_ = [
        ('T', C.c_int32), ('obs_dim', C.c_int32), ('act_dim', C.c_int32), ('hidden_dim', C.c_int32),
        ('activation', C.c_int32), ('use_gae', C.c_int32), ('deterministic', C.c_int32), ('envs_per_block', C.c_int32),
        ('gamma', C.c_double), ('gae_lambda', C.c_double),
        ('d_policy', C.c_void_p), ('d_eps', C.c_void_p),
        ('n_peers', C.c_int32), ('mlp_engine', C.c_int32), ('peer_byte_offset', C.c_int64 * SCG_MAX_PEERS),
        ('multicast_byte_offset', C.c_int64), ('d_norm', C.c_void_p), ('d_obs0', C.c_void_p),
        ('safety_constraints', C.c_int32), ('safety_hidden', C.c_int32), ('d_safety', C.c_void_p), ('safety_slack', C.c_double * SCG_MAX_SAFETY),
        ('d_adv_policy', C.c_void_p), ('d_eps_adv', C.c_void_p), ('adv_act_dim', C.c_int32), ('adv_record', C.c_int32),
    ]


class scg_ppo_args(C.Structure):
    _fields_ = [
        ('obs_dim', C.c_int32), ('act_dim', C.c_int32), ('hidden_dim', C.c_int32), ('activation', C.c_int32),
        ('mini_batch_size', C.c_int32), ('num_mini_batch', C.c_int32), ('opt_epochs', C.c_int32), ('use_clipped_value', C.c_int32),
        ('block_rows', C.c_int64), ('block_stride', C.c_int64 * 6), ('d_key', C.c_void_p * 6),
        ('d_perm', C.c_void_p), ('perm_stride', C.c_int64),
        ('clip_param', C.c_double), ('target_kl', C.c_double), ('entropy_coef', C.c_double),
        ('actor_lr', C.c_double), ('critic_lr', C.c_double), ('beta1', C.c_double), ('beta2', C.c_double), ('adam_eps', C.c_double),
        ('d_params', C.c_void_p), ('d_exp_avg', C.c_void_p), ('d_exp_avg_sq', C.c_void_p), ('d_adam_steps', C.c_void_p),
        ('d_stats', C.c_void_p), ('d_workspace', C.c_void_p), ('workspace_bytes', C.c_size_t), ('d_grad_out', C.c_void_p),
    ]


PPO_KEYS = ('obs', 'act', 'logp', 'adv', 'ret', 'v')      # SCG_PPO_KEY_*

# every symbol include/scg.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_SYMBOLS = {
    'scg_abi_version': (C.c_int, []),
    'scg_create': (C.c_int, [C.POINTER(scg_config), C.POINTER(_VP)]),
    'scg_destroy': (None, [_VP]),
    'scg_last_error': (C.c_char_p, [_VP]),
    'scg_dims': (C.c_int, [_VP] + [C.POINTER(C.c_int32)] * 6),
    'scg_set_reference': (C.c_int, [_VP, _VP, C.c_int32]),
    'scg_set_constraints': (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP]),
    'scg_set_episode_table': (C.c_int, [_VP, _VP, C.c_int32]),
    'scg_set_noise_table': (C.c_int, [_VP, _VP, C.c_int32]),
    'scg_reset': (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    'scg_step': (C.c_int, [_VP, _VP, _VP, C.POINTER(scg_step_out), _VP]),
    'scg_rollout': (C.c_int, [_VP, C.POINTER(scg_rollout_args), C.POINTER(scg_rollout_out), _VP]),
    'scg_set_adversary_control': (C.c_int, [_VP, _VP, _VP]),
    'scg_adversary_dim': (C.c_int, [_VP]),
    'scg_policy_blob_floats': (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    'scg_policy_pack': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP]),
    'scg_state_bytes': (C.c_int, [_VP, C.POINTER(C.c_size_t)]),
    'scg_get_state': (C.c_int, [_VP, _VP, C.c_size_t, _VP]),
    'scg_set_state': (C.c_int, [_VP, _VP, C.c_size_t, _VP]),
    'scg_launch_count': (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    'scg_philox_selftest': (C.c_int, [C.c_int32, _VP, _VP, _VP, C.c_int32]),
    'scg_ppo_param_layout': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32 * 15)]),
    'scg_ppo_workspace_bytes': (C.c_size_t, [C.POINTER(scg_ppo_args)]),
    'scg_ppo_update': (C.c_int, [C.POINTER(scg_ppo_args), _VP]),
    'scg_policy_pack_device': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP]),
}

LIB_NAME = os.environ.get('SCG_LIB_NAME', 'libscg_b200.so')   # SCG_LIB_NAME: developer builds (e.g. the -DSCG_PROF variant)
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

_lib = None


class ScgError(RuntimeError):
    '''Raised when a C-ABI call fails; `code` is the SCG_E* value.'''

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def lib():
    '''Loads (once) and returns the CUDA library.  Fails loudly when it was not built.'''
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f'{LIB_PATH} is missing: build it with `python -m safe_control_gym_b200.build` '
                              '(nvcc, sm_100a).  There is no CPU fallback for the rollout engine.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SYMBOLS.items():
            fn = getattr(handle, name)   # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.scg_abi_version() != SCG_ABI_VERSION:
            raise ImportError(f'{LIB_NAME}: ABI version mismatch')
        _lib = handle
    return _lib


def check(code, handle=None):
    '''Maps a C-ABI return code onto the exception type the reference raises on this path.'''
    if code == SCG_OK:
        return
    msg = lib().scg_last_error(handle)
    msg = msg.decode() if msg else f'scg error {code}'
    if code == SCG_EINVAL:
        raise ValueError(msg)           # e.g. benchmark_env.py:141-142
    if code == SCG_ESTATE:
        raise RuntimeError(msg)         # benchmark_env.py:230-235
    raise ScgError(code, msg)


def new_config():
    cfg = scg_config()
    cfg.struct_size = C.sizeof(scg_config)
    cfg.abi_version = SCG_ABI_VERSION
    return cfg
