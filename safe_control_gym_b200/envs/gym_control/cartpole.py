'''CartPole task on the CUDA engine.

API of safe_control_gym/envs/gym_control/cartpole.py:29-775 (kwargs, spaces, X_GOAL/U_GOAL, action
normalisation, randomisation info, info dict); the PyBullet 2-DoF chain is integrated by the engine with
Bullet's semantics (semi-implicit Euler per sub-step, pole inertia recomputed from the collision box --
pinned against real PyBullet transitions in tests/golden/cartpole_pybullet_transitions.npz).
'''

import math
from copy import deepcopy

import numpy as np

from safe_control_gym_b200 import _abi
from safe_control_gym_b200.envs.benchmark_env import BenchmarkEnv, Cost, Task, get_cost_weight_matrix
from safe_control_gym_b200.envs.constraints import GENERAL_CONSTRAINTS, SymmetricStateConstraint
from safe_control_gym_b200.spaces import Box

# defaults of the packaged cartpole.yaml (gym_control/cartpole.yaml:1-37); note constraint_penalty: -1
DEFAULT_CONFIG = dict(
    ctrl_freq=50, pyb_freq=50, gui=False, normalized_rl_action_space=False, episode_len_sec=5,
    init_state=None, randomized_init=True, init_state_randomization_info=None,
    inertial_prop=None, randomized_inertial_prop=False, inertial_prop_randomization_info=None,
    task='stabilization', task_info=None, cost='rl_reward',
    disturbances=None, adversary_disturbance=None, adversary_disturbance_offset=0.0, adversary_disturbance_scale=0.01,
    constraints=None, done_on_violation=False, use_constraint_penalty=False, constraint_penalty=-1,
    verbose=False,
    obs_wrap_angle=False, obs_goal_horizon=0, rew_state_weight=1.0, rew_act_weight=0.0001, rew_exponential=True,
    done_on_out_of_bound=True,
)

# geometry / masses of assets/cartpole_template.urdf:37,52,58-66
URDF_POLE_BOX = (0.05, 0.05, 1.0)
URDF_POLE_MASS = 0.1
URDF_CART_MASS = 1.0


class CartPole(BenchmarkEnv):
    NAME = 'cartpole'
    SYSTEM = _abi.SCG_SYS_CARTPOLE

    AVAILABLE_CONSTRAINTS = {'abs_bound': SymmetricStateConstraint}
    AVAILABLE_CONSTRAINTS.update(deepcopy(GENERAL_CONSTRAINTS))

    DISTURBANCE_MODES = {'observation': {'dim': 4}, 'action': {'dim': 1}, 'dynamics': {'dim': 2}}

    INERTIAL_PROP_RAND_INFO = {     # cartpole.py:75-90
        'pole_length': {'distrib': 'choice', 'args': [[1, 5, 10]]},
        'cart_mass': {'distrib': 'uniform', 'low': 0.5, 'high': 1.5},
        'pole_mass': {'distrib': 'uniform', 'low': 0.05, 'high': 0.15},
    }
    INIT_STATE_RAND_INFO = {        # cartpole.py:92-113
        'init_x': {'distrib': 'uniform', 'low': -0.05, 'high': 0.05},
        'init_x_dot': {'distrib': 'uniform', 'low': -0.05, 'high': 0.05},
        'init_theta': {'distrib': 'uniform', 'low': -0.05, 'high': 0.05},
        'init_theta_dot': {'distrib': 'uniform', 'low': -0.05, 'high': 0.05},
    }
    TASK_INFO = {                   # cartpole.py:115-123
        'stabilization_goal': [0], 'stabilization_goal_tolerance': 0.05,
        'trajectory_type': 'circle', 'num_cycles': 1, 'trajectory_plane': 'zx',
        'trajectory_position_offset': [0, 0], 'trajectory_scale': 0.2,
    }

    def __init__(self, init_state=None, inertial_prop=None, obs_goal_horizon=0, obs_wrap_angle=False,
                 rew_state_weight=1.0, rew_act_weight=0.0001, rew_exponential=True, done_on_out_of_bound=True,
                 info_mse_metric_state_weight=None, **kwargs):
        self.obs_goal_horizon = obs_goal_horizon
        self.obs_wrap_angle = obs_wrap_angle
        self.rew_state_weight = np.array(rew_state_weight, ndmin=1, dtype=float)
        self.rew_act_weight = np.array(rew_act_weight, ndmin=1, dtype=float)
        self.Q = get_cost_weight_matrix(self.rew_state_weight, 4)
        self.R = get_cost_weight_matrix(self.rew_act_weight, 1)
        self.rew_exponential = rew_exponential
        self.done_on_out_of_bound = done_on_out_of_bound
        if info_mse_metric_state_weight is None:
            self.info_mse_metric_state_weight = np.array([1, 0, 1, 0], ndmin=1, dtype=float)
        elif len(info_mse_metric_state_weight) == 4:
            self.info_mse_metric_state_weight = np.array(info_mse_metric_state_weight, ndmin=1, dtype=float)
        else:
            raise ValueError('[ERROR] in CartPole.__init__(), wrong info_mse_metric_state_weight argument size.')
        super().__init__(init_state=init_state, inertial_prop=inertial_prop, **kwargs)
        # initial state (cartpole.py:186-197)
        if init_state is None:
            self.INIT_X, self.INIT_X_DOT, self.INIT_THETA, self.INIT_THETA_DOT = np.zeros(4)
        elif isinstance(init_state, np.ndarray):
            self.INIT_X, self.INIT_X_DOT, self.INIT_THETA, self.INIT_THETA_DOT = init_state
        elif isinstance(init_state, dict):
            self.INIT_X = init_state.get('init_x', 0)
            self.INIT_X_DOT = init_state.get('init_x_dot', 0)
            self.INIT_THETA = init_state.get('init_theta', 0)
            self.INIT_THETA_DOT = init_state.get('init_theta_dot', 0)
        else:
            raise ValueError('[ERROR] in CartPole.__init__(), init_state incorrect format.')
        # physical properties (cartpole.py:199-213); EFFECTIVE length = half the pole box
        self.GRAVITY_ACC = 9.8
        eff_len, pole_mass, cart_mass = 0.5 * URDF_POLE_BOX[2], URDF_POLE_MASS, URDF_CART_MASS
        if inertial_prop is None:
            self.EFFECTIVE_POLE_LENGTH, self.POLE_MASS, self.CART_MASS = eff_len, pole_mass, cart_mass
        elif isinstance(inertial_prop, dict):
            self.EFFECTIVE_POLE_LENGTH = inertial_prop.get('pole_length', eff_len)
            self.POLE_MASS = inertial_prop.get('pole_mass', pole_mass)
            self.CART_MASS = inertial_prop.get('cart_mass', cart_mass)
        else:
            raise ValueError('[ERROR] in CartPole.__init__(), inertial_prop incorrect format.')
        # references (cartpole.py:215-233)
        self.U_GOAL = np.zeros(1)
        if self.TASK == Task.STABILIZATION:
            self.X_GOAL = np.hstack([self.TASK_INFO['stabilization_goal'][0], 0., 0., 0.])
        else:
            pos, vel, _ = self._generate_trajectory(traj_type=self.TASK_INFO['trajectory_type'],
                                                    traj_length=self.EPISODE_LEN_SEC,
                                                    num_cycles=self.TASK_INFO['num_cycles'],
                                                    traj_plane=self.TASK_INFO['trajectory_plane'],
                                                    position_offset=np.array(self.TASK_INFO['trajectory_position_offset']),
                                                    scaling=self.TASK_INFO['trajectory_scale'],
                                                    sample_time=self.CTRL_TIMESTEP)
            self.X_GOAL = np.vstack([pos[:, 0], vel[:, 0], np.zeros(pos.shape[0]), np.zeros(pos.shape[0])]).transpose()

    # -- spaces (cartpole.py:439-477) -----------------------------------------------------------
    def _set_action_space(self):
        self.action_scale = 10
        self.physical_action_bounds = (-1 * np.atleast_1d(self.action_scale), np.atleast_1d(self.action_scale))
        self.action_threshold = 1 if self.NORMALIZED_RL_ACTION_SPACE else self.action_scale
        self.action_space = Box(low=-self.action_threshold, high=self.action_threshold, shape=(1,))
        self.ACTION_LABELS = ['U']
        self.ACTION_UNITS = ['N'] if not self.NORMALIZED_RL_ACTION_SPACE else ['-']

    def _set_observation_space(self):
        self.x_threshold = 2.4
        self.x_dot_threshold = 20
        self.theta_threshold_radians = 90 * math.pi / 180
        self.theta_dot_threshold = 20
        obs_bound = np.array([self.x_threshold * 2, self.x_dot_threshold, self.theta_threshold_radians * 2, self.theta_dot_threshold])
        self.state_space = Box(low=-obs_bound, high=obs_bound, dtype=np.float32)
        if self.COST == Cost.RL_REWARD and self.TASK == Task.TRAJ_TRACKING and self.obs_goal_horizon > 0:
            obs_bound = np.concatenate([obs_bound] * (1 + self.obs_goal_horizon))
        elif self.COST == Cost.RL_REWARD and self.TASK == Task.STABILIZATION and self.obs_goal_horizon > 0:
            obs_bound = np.concatenate([obs_bound] * 2)
        self.observation_space = Box(low=-obs_bound, high=obs_bound, dtype=np.float32)
        self.STATE_LABELS = ['x', 'x_dot', 'theta', 'theta_dot']
        self.STATE_UNITS = ['m', 'm/s', 'rad', 'rad/s']

    def normalize_action(self, action):
        return action / self.action_scale if self.NORMALIZED_RL_ACTION_SPACE else action

    def denormalize_action(self, action):
        return self.action_scale * action if self.NORMALIZED_RL_ACTION_SPACE else action

    # -- engine -----------------------------------------------------------------------------------
    def _inertial_keys(self):
        return ['pole_length', 'cart_mass', 'pole_mass']     # draw order of cartpole.py:287

    def _init_keys(self):
        return ['init_x', 'init_x_dot', 'init_theta', 'init_theta_dot']

    def _fill_system_config(self, cfg):
        cfg.hover_thrust = float(self.action_scale)           # normalised action = action_scale * a (cartpole.py:527-528)
        cfg.norm_act_scale = 0.0
        cfg.x_threshold = float(self.x_threshold)
        cfg.theta_threshold = float(self.theta_threshold_radians)
        cfg.gravity = float(self.GRAVITY_ACC)
        cfg.pole_width = float(URDF_POLE_BOX[0])
        cfg.inertial_nominal[0] = float(self.EFFECTIVE_POLE_LENGTH)
        cfg.inertial_nominal[1] = float(self.CART_MASS)
        cfg.inertial_nominal[2] = float(self.POLE_MASS)
        for j, v in enumerate([self.INIT_X, self.INIT_X_DOT, self.INIT_THETA, self.INIT_THETA_DOT]):
            cfg.init_nominal[j] = float(v)

    def _check_inertial(self, additive):
        vals = [self.EFFECTIVE_POLE_LENGTH + additive[0], self.CART_MASS + additive[1], self.POLE_MASS + additive[2]]
        if any(v < 0 for v in vals):
            raise ValueError('[ERROR] in CartPole.reset(), negative randomized inertial properties.')

    def _get_reset_info(self, inertial_additive):
        info = {
            'symbolic_model': self.symbolic,
            'physical_parameters': {
                'pole_effective_length': self.EFFECTIVE_POLE_LENGTH + inertial_additive[0],
                'pole_mass': self.POLE_MASS + inertial_additive[2],
                'cart_mass': self.CART_MASS + inertial_additive[1],
            },
            'x_reference': self.X_GOAL,
            'u_reference': self.U_GOAL,
        }
        if self.constraints is not None:
            info['symbolic_constraints'] = self.constraints.get_all_symbolic_models()
        return info
