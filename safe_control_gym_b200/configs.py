'''Task configurations of the benchmark (BASELINE.json configs) as plain dicts.

Values restate the reference's example overrides (facts, not code):
  QUADROTOR_2D_TRACK  examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml  (task_config)
  QUADROTOR_3D_TRACK  examples/rl/config_overrides/quadrotor_3D/quadrotor_3D_track.yaml  (task_config)
  CARTPOLE_STAB       examples/rl/config_overrides/cartpole/cartpole_stab.yaml           (task_config)
Use as `make('quadrotor', **QUADROTOR_2D_TRACK)` exactly like the reference's `env_func = partial(make, task, **task_config)`.
'''

from copy import deepcopy


def _uniform(low, high):
    return {'distrib': 'uniform', 'low': low, 'high': high}


QUADROTOR_2D_TRACK = dict(
    seed=1337, ctrl_freq=50, pyb_freq=1000, physics='pyb', quad_type=2, normalized_rl_action_space=True,
    init_state=dict(init_x=0, init_x_dot=0, init_z=1, init_z_dot=0, init_theta=0, init_theta_dot=0),
    randomized_init=True, randomized_inertial_prop=False,
    init_state_randomization_info=dict(
        init_x=_uniform(-2, 2), init_x_dot=_uniform(-1, 1), init_z=_uniform(0.3, 2), init_z_dot=_uniform(-1, 1),
        init_theta=_uniform(-0.2, 0.2), init_theta_dot=_uniform(-1.5, 1.5)),
    task='traj_tracking',
    task_info=dict(trajectory_type='figure8', num_cycles=1, trajectory_plane='xz', trajectory_position_offset=[0, 1],
                   trajectory_scale=1),
    inertial_prop=dict(M=0.027, Iyy=1.4e-05),
    episode_len_sec=5, cost='rl_reward', obs_goal_horizon=1,
    rew_state_weight=[1, 0.01, 1, 0.01, 0.01, 0.01], rew_act_weight=0.01, rew_exponential=True,
    constraints=[
        dict(constraint_form='default_constraint', constrained_variable='state',
             upper_bounds=[2, 1, 2, 1, 0.2, 1.5], lower_bounds=[-2, -1, 0, -1, -0.2, -1.5]),
        dict(constraint_form='default_constraint', constrained_variable='input',
             upper_bounds=[0.29, 0.29], lower_bounds=[0.06, 0.06]),
    ],
    done_on_out_of_bound=True, done_on_violation=False,
)

QUADROTOR_3D_TRACK = dict(
    seed=1337, ctrl_freq=50, pyb_freq=1000, physics='pyb', quad_type=3, normalized_rl_action_space=True,
    init_state=dict(init_x=0.4, init_x_dot=0, init_y=0.4, init_y_dot=0, init_z=1.4, init_z_dot=0,
                    init_phi=0, init_theta=0, init_psi=0, init_p=0, init_q=0, init_r=0),
    randomized_init=True, randomized_inertial_prop=False,
    init_state_randomization_info=dict(
        init_x=_uniform(-2, 2), init_x_dot=_uniform(-1, 1), init_y=_uniform(-2, 2), init_y_dot=_uniform(-1, 1),
        init_z=_uniform(0.3, 2), init_z_dot=_uniform(-1, 1), init_phi=_uniform(-0.2, 0.2), init_theta=_uniform(-0.2, 0.2),
        init_psi=_uniform(-0.2, 0.2), init_p=_uniform(-1, 1), init_q=_uniform(-1, 1), init_r=_uniform(-1, 1)),
    task='traj_tracking',
    task_info=dict(trajectory_type='figure8', num_cycles=1, trajectory_plane='xz', trajectory_position_offset=[0, 1],
                   trajectory_scale=1, proj_point=[0, 0, 0.5], proj_normal=[0, 1, 1]),
    inertial_prop=dict(M=0.027, Ixx=1.4e-5, Iyy=1.4e-5, Izz=2.17e-5),
    episode_len_sec=5, cost='rl_reward', obs_goal_horizon=1,
    rew_state_weight=[1, 0.01, 1, 0.01, 1, 0.01, 0.5, 0.5, 0.5, 0.01, 0.01, 0.01], rew_act_weight=0.0001, rew_exponential=True,
    info_mse_metric_state_weight=[1, 0.01, 1, 0.01, 1, 0.01, 0.5, 0.5, 0.5, 0.01, 0.01, 0.01],
    constraints=[
        dict(constraint_form='default_constraint', constrained_variable='state',
             upper_bounds=[2, 1, 2, 1, 2, 1, 0.2, 0.2, 0.2, 1, 1, 1],
             lower_bounds=[-2, -1, -2, -1, 0, -1, -0.2, -0.2, -0.2, -1, -1, -1]),
        dict(constraint_form='default_constraint', constrained_variable='input',
             upper_bounds=[0.148] * 4, lower_bounds=[0.029] * 4),
    ],
    done_on_out_of_bound=True, done_on_violation=False,
)

CARTPOLE_STAB = dict(
    seed=42, ctrl_freq=15, pyb_freq=750, normalized_rl_action_space=True,
    init_state=dict(init_x=0.1, init_x_dot=-1.5, init_theta=-0.155, init_theta_dot=0.75),
    randomized_init=True, randomized_inertial_prop=False,
    init_state_randomization_info=dict(init_x=_uniform(-2, 2), init_x_dot=_uniform(-2, 2), init_theta=_uniform(-0.16, 0.16),
                                       init_theta_dot=_uniform(-1, 1)),
    task='stabilization', task_info=dict(stabilization_goal=[0.7, 0], stabilization_goal_tolerance=0.0),
    inertial_prop=dict(pole_length=0.5, cart_mass=1, pole_mass=0.1),
    episode_len_sec=10, cost='rl_reward', obs_goal_horizon=0,
    rew_state_weight=[1, 1, 1, 1], rew_act_weight=0.1, rew_exponential=True,
    constraints=[
        dict(constraint_form='default_constraint', constrained_variable='state',
             upper_bounds=[2, 2, 0.16, 1], lower_bounds=[-2, -2, -0.16, -1]),
        dict(constraint_form='default_constraint', constrained_variable='input'),
    ],
    done_on_out_of_bound=True, done_on_violation=False,
)


# BASELINE.json config 2: "CartPole stabilization, 4096 batched envs, random-MLP policy rollout" = the packaged cartpole.yaml defaults
# (ctrl 50 / pyb 50, init state uniform +-0.05 per component cartpole.py:92-113, obs 4, act 1; SURVEY.md 8d "C2").
CARTPOLE_C2 = dict(ctrl_freq=50, pyb_freq=50, randomized_init=True, task='stabilization', cost='rl_reward', obs_goal_horizon=0,
                   normalized_rl_action_space=False, done_on_out_of_bound=True)

# BASELINE.json config 4: "Quadrotor-3D lemniscate tracking with parametric mass/inertia disturbances + constraint eval": the 3-D
# tracking YAML verbatim + additive zero-mean inertial randomisation (masses stay physical) + white-noise force on the dynamics
# (SURVEY.md 8d "C4").  `honor_randomization_info` keeps these ranges instead of the class defaults (quadrotor.py:208,233 quirk).
QUADROTOR_3D_TRACK_C4 = dict(
    QUADROTOR_3D_TRACK, randomized_inertial_prop=True, honor_randomization_info=True,
    inertial_prop_randomization_info=dict(M=_uniform(-0.005, 0.005), Ixx=_uniform(-1e-6, 1e-6), Iyy=_uniform(-1e-6, 1e-6),
                                          Izz=_uniform(-1e-6, 1e-6)),
    disturbances=dict(dynamics=[dict(disturbance_func='white_noise', std=0.01)]),
)

# name -> (env id, task config).  c2/c3/c4 are the BASELINE.json configs bench.py --config selects.
BENCH_CONFIGS = {
    'c2': ('cartpole', 'cartpole_c2', 4096),
    'c3': ('quadrotor', 'quadrotor_2D_track', 4096),
    'c4': ('quadrotor', 'quadrotor_3D_track_c4', 8192),
}


def get(name):
    cfg = deepcopy({'quadrotor_2D_track': QUADROTOR_2D_TRACK, 'quadrotor_3D_track': QUADROTOR_3D_TRACK,
                    'cartpole_stab': CARTPOLE_STAB, 'cartpole_c2': CARTPOLE_C2, 'quadrotor_3D_track_c4': QUADROTOR_3D_TRACK_C4}[name])
    if name == 'quadrotor_3D_track_c4':      # initial-state ranges: keep the reference quirk (class defaults win, quadrotor.py:208)
        cfg['init_state_randomization_info'] = None
    return cfg
