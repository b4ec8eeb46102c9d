'''1D / 2D / 3D Crazyflie quadrotor task on the CUDA engine.

API of safe_control_gym/envs/gym_pybullet_drones/quadrotor.py:25-941 + base_aviary.py (kwargs, spaces,
X_GOAL / U_GOAL generation incl. the 3D plane projection, action normalisation, randomisation info, info
dict).  PyBullet's rigid-body step is integrated by the engine (see csrc/scg_device.cuh).  Only the default
`physics: pyb` model is provided (ground effect / drag / downwash variants are SURVEY.md 8f "next").
'''

import math
from copy import deepcopy

import numpy as np

from safe_control_gym_b200 import _abi
from safe_control_gym_b200.envs.benchmark_env import BenchmarkEnv, Cost, Task, get_cost_weight_matrix
from safe_control_gym_b200.envs.constraints import GENERAL_CONSTRAINTS
from safe_control_gym_b200.envs.gym_pybullet_drones.quadrotor_utils import CF2X, QuadType
from safe_control_gym_b200.spaces import Box

# defaults of the packaged quadrotor.yaml (gym_pybullet_drones/quadrotor.yaml:1-39); constraint_penalty: -1
DEFAULT_CONFIG = dict(
    ctrl_freq=60, pyb_freq=240, physics='pyb', gui=False, quad_type=2, normalized_rl_action_space=False, episode_len_sec=5,
    init_state=None, randomized_init=False, init_state_randomization_info=None,
    inertial_prop=None, randomized_inertial_prop=False, inertial_prop_randomization_info=None,
    task='stabilization', task_info=None, cost='rl_reward',
    disturbances=None, adversary_disturbance=None, adversary_disturbance_offset=0.0, adversary_disturbance_scale=0.01,
    constraints=None, done_on_violation=False, use_constraint_penalty=False, constraint_penalty=-1,
    verbose=False,
    norm_act_scale=0.1, obs_goal_horizon=0, rew_state_weight=1.0, rew_act_weight=0.0001, rew_exponential=True,
    done_on_out_of_bound=True,
)


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / math.sqrt(float(np.dot(v, v)))


def project_onto_plane(pos, vel, point, normal):
    '''Orthogonal projection used for the 3D reference (math_and_models/transformations.py:54-125).

    The reference multiplies homogeneous rows [p, 1] by M^T for the positions AND for the velocities, so the
    velocities also pick up the plane offset; reproduced on purpose.'''
    n = _unit(normal[:3])
    p0 = np.asarray(point[:3], dtype=np.float64)
    lin = np.identity(3) - np.outer(n, n)
    shift = np.dot(p0, n) * n
    return pos @ lin.T + shift, vel @ lin.T + shift


class Quadrotor(BenchmarkEnv):
    NAME = 'quadrotor'
    AVAILABLE_CONSTRAINTS = deepcopy(GENERAL_CONSTRAINTS)
    DISTURBANCE_MODES = {'observation': {'dim': -1}, 'action': {'dim': -1}, 'dynamics': {'dim': -1}}

    BASE_INERTIAL_PROP_RAND_INFO = {    # quadrotor.py:47-68
        'M': {'distrib': 'uniform', 'low': 0.022, 'high': 0.032},
        'Ixx': {'distrib': 'uniform', 'low': 1.3e-5, 'high': 1.5e-5},
        'Iyy': {'distrib': 'uniform', 'low': 1.3e-5, 'high': 1.5e-5},
        'Izz': {'distrib': 'uniform', 'low': 2.07e-5, 'high': 2.27e-5},
    }
    BASE_INIT_STATE_RAND_INFO = {       # quadrotor.py:70-136
        'init_x': {'distrib': 'uniform', 'low': -0.5, 'high': 0.5},
        'init_x_dot': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
        'init_y': {'distrib': 'uniform', 'low': -0.5, 'high': 0.5},
        'init_y_dot': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
        'init_z': {'distrib': 'uniform', 'low': 0.1, 'high': 1.5},
        'init_z_dot': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
        'init_phi': {'distrib': 'uniform', 'low': -0.3, 'high': 0.3},
        'init_theta': {'distrib': 'uniform', 'low': -0.3, 'high': 0.3},
        'init_psi': {'distrib': 'uniform', 'low': -0.3, 'high': 0.3},
        'init_p': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
        'init_theta_dot': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
        'init_q': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
        'init_r': {'distrib': 'uniform', 'low': -0.01, 'high': 0.01},
    }
    TASK_INFO = {                       # quadrotor.py:138-148
        'stabilization_goal': [0, 1], 'stabilization_goal_tolerance': 0.05,
        'trajectory_type': 'circle', 'num_cycles': 1, 'trajectory_plane': 'zx',
        'trajectory_position_offset': [0.5, 0], 'trajectory_scale': -0.5,
        'proj_point': [0, 0, 0.5], 'proj_normal': [0, 1, 1],
    }
    INIT_STATE_LABELS = {               # quadrotor.py:209-214
        QuadType.ONE_D: ['init_x', 'init_x_dot'],
        QuadType.TWO_D: ['init_x', 'init_x_dot', 'init_z', 'init_z_dot', 'init_theta', 'init_theta_dot'],
        QuadType.THREE_D: ['init_x', 'init_x_dot', 'init_y', 'init_y_dot', 'init_z', 'init_z_dot',
                           'init_phi', 'init_theta', 'init_psi', 'init_p', 'init_q', 'init_r'],
    }

    def __init__(self, init_state=None, inertial_prop=None, quad_type=QuadType.TWO_D, norm_act_scale=0.1, obs_goal_horizon=0,
                 rew_state_weight=1.0, rew_act_weight=0.0001, rew_exponential=True, done_on_out_of_bound=True,
                 info_mse_metric_state_weight=None, physics='pyb', drone_model='cf2x', record=False,
                 honor_randomization_info=False, **kwargs):
        self.QUAD_TYPE = QuadType(quad_type)
        self.SYSTEM = int(self.QUAD_TYPE)          # SCG_SYS_QUAD{1,2,3}D share the enum values
        # Physics enum (base_aviary.py:32-40).  `dyn` is dead code in the reference (it reads an undefined attribute, base_aviary.py:495-498).
        self.PHYSICS = str(getattr(physics, 'value', physics))
        if self.PHYSICS not in _abi.PHYSICS:
            raise NotImplementedError(f'physics \'{self.PHYSICS}\' is not available (pyb, pyb_gnd, pyb_drag, pyb_dw, pyb_gnd_drag_dw)')
        if str(getattr(drone_model, 'value', drone_model)) != 'cf2x':
            raise NotImplementedError('only the cf2x drone model exists')
        self.norm_act_scale = norm_act_scale
        self.obs_goal_horizon = obs_goal_horizon
        self.rew_state_weight = np.array(rew_state_weight, ndmin=1, dtype=float)
        self.rew_act_weight = np.array(rew_act_weight, ndmin=1, dtype=float)
        self.rew_exponential = rew_exponential
        self.done_on_out_of_bound = done_on_out_of_bound
        nx = {QuadType.ONE_D: 2, QuadType.TWO_D: 6, QuadType.THREE_D: 12}[self.QUAD_TYPE]
        if info_mse_metric_state_weight is None:
            default = {QuadType.ONE_D: [1, 0], QuadType.TWO_D: [1, 0, 1, 0, 0, 0],
                       QuadType.THREE_D: [1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0]}[self.QUAD_TYPE]
            self.info_mse_metric_state_weight = np.array(default, ndmin=1, dtype=float)
        elif len(info_mse_metric_state_weight) == nx:
            self.info_mse_metric_state_weight = np.array(info_mse_metric_state_weight, ndmin=1, dtype=float)
        else:
            raise ValueError('[ERROR] in Quadrotor.__init__(), wrong info_mse_metric_state_weight argument size.')
        # URDF constants (base_aviary.py:76-131)
        self.GRAVITY_ACC = 9.8
        self.MASS, self.L = CF2X['mass'], CF2X['arm']
        self.THRUST2WEIGHT_RATIO = CF2X['thrust2weight']
        self.J = np.diag([CF2X['ixx'], CF2X['iyy'], CF2X['izz']])
        self.J_INV = np.linalg.inv(self.J)
        self.KF, self.KM = CF2X['kf'], CF2X['km']
        self.PWM2RPM_SCALE, self.PWM2RPM_CONST = CF2X['pwm2rpm_scale'], CF2X['pwm2rpm_const']
        self.MIN_PWM, self.MAX_PWM = CF2X['pwm_min'], CF2X['pwm_max']
        self.COLLISION_H, self.COLLISION_R, self.COLLISION_Z_OFFSET = CF2X['collision_h'], CF2X['collision_r'], CF2X['collision_z_offset']
        self.GROUND_PLANE_Z = -0.05
        self.GRAVITY = self.GRAVITY_ACC * self.MASS
        self.HOVER_RPM = np.sqrt(self.GRAVITY / (4 * self.KF))
        self.MAX_RPM = np.sqrt((self.THRUST2WEIGHT_RATIO * self.GRAVITY) / (4 * self.KF))
        self.MAX_THRUST = 4 * self.KF * self.MAX_RPM**2
        self.GND_EFF_COEFF, self.PROP_RADIUS = CF2X['gnd_eff_coeff'], CF2X['prop_radius']
        self.DRAG_COEFF = np.array([CF2X['drag_coeff_xy'], CF2X['drag_coeff_xy'], CF2X['drag_coeff_z']])
        self.DW_COEFF_1, self.DW_COEFF_2, self.DW_COEFF_3 = CF2X['dw_coeff_1'], CF2X['dw_coeff_2'], CF2X['dw_coeff_3']
        self.GND_EFF_H_CLIP = 0.25 * self.PROP_RADIUS * np.sqrt((15 * self.MAX_RPM**2 * self.KF * self.GND_EFF_COEFF) / self.MAX_THRUST)   # base_aviary.py:130-131

        super().__init__(init_state=init_state, inertial_prop=inertial_prop, **kwargs)

        # initial state (quadrotor.py:207-226).  QUIRK kept on purpose: the reference overwrites whatever
        # `init_state_randomization_info` / `inertial_prop_randomization_info` the caller passed with the class
        # defaults (quadrotor.py:208 and :233 run AFTER BenchmarkEnv.__init__ stored the kwargs), so e.g. the ranges in
        # examples/rl/config_overrides/quadrotor_2D/quadrotor_2D_track.yaml:19-43 never take effect.
        # `honor_randomization_info=True` (engine-only kwarg) uses the caller's ranges instead.
        user_init_info = kwargs.get('init_state_randomization_info')
        user_inertial_info = kwargs.get('inertial_prop_randomization_info')
        if not (honor_randomization_info and user_init_info is not None):
            self.INIT_STATE_RAND_INFO = deepcopy(self.BASE_INIT_STATE_RAND_INFO)
        labels = self.INIT_STATE_LABELS[self.QUAD_TYPE]
        if init_state is None:
            for name in self.BASE_INIT_STATE_RAND_INFO:
                self.__dict__[name.upper()] = 0.
        elif isinstance(init_state, np.ndarray):
            for i, name in enumerate(labels):
                self.__dict__[name.upper()] = init_state[i]
        elif isinstance(init_state, dict):
            for name in labels:
                self.__dict__[name.upper()] = init_state.get(name, 0.)
        else:
            raise ValueError('[ERROR] in Quadrotor.__init__(), init_state incorrect format.')
        self.INIT_STATE_RAND_INFO = {k: v for k, v in self.INIT_STATE_RAND_INFO.items() if k in labels}
        # inertial randomisation info restricted to the quad type (quadrotor.py:232-242)
        if honor_randomization_info and user_inertial_info is not None:
            self.INERTIAL_PROP_RAND_INFO = deepcopy(user_inertial_info)
        else:
            self.INERTIAL_PROP_RAND_INFO = deepcopy(self.BASE_INERTIAL_PROP_RAND_INFO)
        if self.QUAD_TYPE == QuadType.ONE_D:
            for k in ('Ixx', 'Iyy', 'Izz'):
                self.INERTIAL_PROP_RAND_INFO.pop(k, None)
        elif self.QUAD_TYPE == QuadType.TWO_D:
            for k in ('Ixx', 'Izz'):
                self.INERTIAL_PROP_RAND_INFO.pop(k, None)
        # inertial overrides (quadrotor.py:244-259)
        if inertial_prop is None:
            pass
        elif isinstance(inertial_prop, dict):
            self.MASS = inertial_prop.get('M', self.MASS)
            self.J[0, 0] = inertial_prop.get('Ixx', self.J[0, 0])
            self.J[1, 1] = inertial_prop.get('Iyy', self.J[1, 1])
            self.J[2, 2] = inertial_prop.get('Izz', self.J[2, 2])
        elif self.QUAD_TYPE == QuadType.ONE_D and np.array(inertial_prop).shape == (1,):
            self.MASS = inertial_prop[0]
        elif self.QUAD_TYPE == QuadType.TWO_D and np.array(inertial_prop).shape == (2,):
            self.MASS, self.J[1, 1] = inertial_prop
        elif self.QUAD_TYPE == QuadType.THREE_D and np.array(inertial_prop).shape == (4,):
            self.MASS, self.J[0, 0], self.J[1, 1], self.J[2, 2] = inertial_prop
        else:
            raise ValueError('[ERROR] in Quadrotor.__init__(), inertial_prop incorrect format.')
        # references (quadrotor.py:261-323).  U_GOAL uses the overridden nominal mass, hover_thrust the URDF mass.
        self.U_GOAL = np.ones(self.action_dim) * self.MASS * self.GRAVITY_ACC / self.action_dim
        if self.TASK == Task.STABILIZATION:
            goal = self.TASK_INFO['stabilization_goal']
            if self.QUAD_TYPE == QuadType.ONE_D:
                self.X_GOAL = np.hstack([goal[1], 0.0])
            elif self.QUAD_TYPE == QuadType.TWO_D:
                self.X_GOAL = np.hstack([goal[0], 0.0, goal[1], 0.0, 0.0, 0.0])
            else:
                self.X_GOAL = np.hstack([goal[0], 0.0, goal[1], 0.0, goal[2], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        else:
            pos, vel, _ = self._generate_trajectory(traj_type=self.TASK_INFO['trajectory_type'],
                                                    traj_length=self.EPISODE_LEN_SEC,
                                                    num_cycles=self.TASK_INFO['num_cycles'],
                                                    traj_plane=self.TASK_INFO['trajectory_plane'],
                                                    position_offset=self.TASK_INFO['trajectory_position_offset'],
                                                    scaling=self.TASK_INFO['trajectory_scale'],
                                                    sample_time=self.CTRL_TIMESTEP)
            zeros = np.zeros(pos.shape[0])
            if self.QUAD_TYPE == QuadType.ONE_D:
                cols = [pos[:, 2], vel[:, 2]]
            elif self.QUAD_TYPE == QuadType.TWO_D:
                cols = [pos[:, 0], vel[:, 0], pos[:, 2], vel[:, 2], zeros, zeros]
            else:
                pos, vel = project_onto_plane(pos, vel, self.TASK_INFO['proj_point'], self.TASK_INFO['proj_normal'])
                cols = [pos[:, 0], vel[:, 0], pos[:, 1], vel[:, 1], pos[:, 2], vel[:, 2]] + [zeros] * 6
            self.X_GOAL = np.vstack(cols).transpose()
        nu = self.action_dim
        self.Q = get_cost_weight_matrix(self.rew_state_weight, nx)
        self.R = get_cost_weight_matrix(self.rew_act_weight, nu)

    # -- spaces (quadrotor.py:606-712) ----------------------------------------------------------
    def _set_action_space(self):
        action_dim = {QuadType.ONE_D: 1, QuadType.TWO_D: 2, QuadType.THREE_D: 4}[self.QUAD_TYPE]
        self.ACTION_LABELS = {1: ['T'], 2: ['T1', 'T2'], 4: ['T1', 'T2', 'T3', 'T4']}[action_dim]
        self.ACTION_UNITS = (['N'] if not self.NORMALIZED_RL_ACTION_SPACE else ['-']) * action_dim
        n_mot = 4 / action_dim
        a_low = self.KF * n_mot * (self.PWM2RPM_SCALE * self.MIN_PWM + self.PWM2RPM_CONST)**2
        a_high = self.KF * n_mot * (self.PWM2RPM_SCALE * self.MAX_PWM + self.PWM2RPM_CONST)**2
        self.physical_action_bounds = (np.full(action_dim, a_low, np.float32), np.full(action_dim, a_high, np.float32))
        if self.NORMALIZED_RL_ACTION_SPACE:
            self.hover_thrust = self.GRAVITY_ACC * self.MASS / action_dim     # URDF mass: runs before the inertial override
            self.action_space = Box(low=-np.ones(action_dim), high=np.ones(action_dim), dtype=np.float32)
        else:
            self.action_space = Box(low=self.physical_action_bounds[0], high=self.physical_action_bounds[1], dtype=np.float32)

    def _set_observation_space(self):
        self.x_threshold = self.y_threshold = self.z_threshold = 2
        self.x_dot_threshold = self.y_dot_threshold = self.z_dot_threshold = 30
        self.phi_threshold_radians = self.theta_threshold_radians = 85 * math.pi / 180
        self.psi_threshold_radians = 180 * math.pi / 180
        self.phi_dot_threshold_radians = self.theta_dot_threshold_radians = self.psi_dot_threshold_radians = 500 * math.pi / 180
        if self.QUAD_TYPE == QuadType.ONE_D:
            low = np.array([self.GROUND_PLANE_Z, -self.z_dot_threshold])
            high = np.array([self.z_threshold, self.z_dot_threshold])
            self.STATE_LABELS, self.STATE_UNITS = ['z', 'z_dot'], ['m', 'm/s']
        elif self.QUAD_TYPE == QuadType.TWO_D:
            low = np.array([-self.x_threshold, -self.x_dot_threshold, self.GROUND_PLANE_Z, -self.z_dot_threshold,
                            -self.theta_threshold_radians, -self.theta_dot_threshold_radians])
            high = np.array([self.x_threshold, self.x_dot_threshold, self.z_threshold, self.z_dot_threshold,
                             self.theta_threshold_radians, self.theta_dot_threshold_radians])
            self.STATE_LABELS = ['x', 'x_dot', 'z', 'z_dot', 'theta', 'theta_dot']
            self.STATE_UNITS = ['m', 'm/s', 'm', 'm/s', 'rad', 'rad/s']
        else:
            low = np.array([-self.x_threshold, -self.x_dot_threshold, -self.y_threshold, -self.y_dot_threshold,
                            self.GROUND_PLANE_Z, -self.z_dot_threshold,
                            -self.phi_threshold_radians, -self.theta_threshold_radians, -self.psi_threshold_radians,
                            -self.phi_dot_threshold_radians, -self.theta_dot_threshold_radians, -self.psi_dot_threshold_radians])
            high = np.array([self.x_threshold, self.x_dot_threshold, self.y_threshold, self.y_dot_threshold,
                             self.z_threshold, self.z_dot_threshold,
                             self.phi_threshold_radians, self.theta_threshold_radians, self.psi_threshold_radians,
                             self.phi_dot_threshold_radians, self.theta_dot_threshold_radians, self.psi_dot_threshold_radians])
            self.STATE_LABELS = ['x', 'x_dot', 'y', 'y_dot', 'z', 'z_dot', 'phi', 'theta', 'psi', 'p', 'q', 'r']
            self.STATE_UNITS = ['m', 'm/s', 'm', 'm/s', 'm', 'm/s', 'rad', 'rad', 'rad', 'rad/s', 'rad/s', 'rad/s']
        self.state_space = Box(low=low, high=high, dtype=np.float32)
        if self.COST == Cost.RL_REWARD and self.TASK == Task.TRAJ_TRACKING and self.obs_goal_horizon > 0:
            mul = 1 + self.obs_goal_horizon
            low, high = np.concatenate([low] * mul), np.concatenate([high] * mul)
        elif self.COST == Cost.RL_REWARD and self.TASK == Task.STABILIZATION and self.obs_goal_horizon > 0:
            low, high = np.concatenate([low] * 2), np.concatenate([high] * 2)
        self.observation_space = Box(low=low, high=high, dtype=np.float32)

    def _setup_disturbances(self):
        # quadrotor.py:714-720.  The observation disturbance is declared with dim = obs_dim but added to the
        # state_dim-long state (quadrotor.py:805-807): in the reference that only works when they coincide.
        self.DISTURBANCE_MODES = deepcopy(self.DISTURBANCE_MODES)
        self.DISTURBANCE_MODES['observation']['dim'] = self.obs_dim
        self.DISTURBANCE_MODES['action']['dim'] = self.action_dim
        self.DISTURBANCE_MODES['dynamics']['dim'] = int(self.QUAD_TYPE)
        if self.DISTURBANCES is not None and 'observation' in self.DISTURBANCES and self.obs_dim != self.state_dim:
            raise ValueError('operands could not be broadcast together: observation disturbances need obs_dim == state_dim '
                             '(obs_goal_horizon = 0), as in the reference')
        super()._setup_disturbances()

    def normalize_action(self, action):
        if self.NORMALIZED_RL_ACTION_SPACE:
            action = (action / self.hover_thrust - 1) / self.norm_act_scale
        return action

    def denormalize_action(self, action):
        if self.NORMALIZED_RL_ACTION_SPACE:
            action = (1 + self.norm_act_scale * action) * self.hover_thrust
        return action

    # -- engine -----------------------------------------------------------------------------------
    def _inertial_keys(self):
        return ['M', 'Ixx', 'Iyy', 'Izz']                       # dict order of quadrotor.py:346-351

    def _init_keys(self):
        return self.INIT_STATE_LABELS[self.QUAD_TYPE]

    def _fill_system_config(self, cfg):
        cfg.norm_act_scale = float(self.norm_act_scale)
        cfg.hover_thrust = float(getattr(self, 'hover_thrust', 0.0))
        cfg.gravity = float(self.GRAVITY_ACC)
        cfg.kf, cfg.km = float(self.KF), float(self.KM)
        cfg.arm = float(CF2X['prop_offset'])
        cfg.pwm2rpm_scale, cfg.pwm2rpm_const = float(self.PWM2RPM_SCALE), float(self.PWM2RPM_CONST)
        cfg.pwm_min, cfg.pwm_max = float(self.MIN_PWM), float(self.MAX_PWM)
        cfg.inertial_nominal[0] = float(self.MASS)
        for j in range(3):
            cfg.inertial_nominal[1 + j] = float(self.J[j, j])
        for j, name in enumerate(self._init_keys()):
            cfg.init_nominal[j] = float(self.__dict__[name.upper()])
        cfg.physics = _abi.PHYSICS[self.PHYSICS]
        cfg.gnd_eff_coeff, cfg.prop_radius, cfg.gnd_eff_h_clip = float(self.GND_EFF_COEFF), float(self.PROP_RADIUS), float(self.GND_EFF_H_CLIP)
        cfg.drag_coeff_xy, cfg.drag_coeff_z = float(self.DRAG_COEFF[0]), float(self.DRAG_COEFF[2])

    def _check_inertial(self, additive):
        vals = [self.MASS + additive[0], self.J[0, 0] + additive[1], self.J[1, 1] + additive[2], self.J[2, 2] + additive[3]]
        if any(v < 0 for v in vals):
            raise ValueError('[ERROR] in Quadrotor.reset(), negative randomized inertial properties.')

    def _get_reset_info(self, inertial_additive):
        info = {
            'symbolic_model': self.symbolic,
            'physical_parameters': {
                'quadrotor_mass': self.MASS + inertial_additive[0],
                'quadrotor_inertia': [self.J[0, 0] + inertial_additive[1], self.J[1, 1] + inertial_additive[2],
                                      self.J[2, 2] + inertial_additive[3]],
            },
            'x_reference': self.X_GOAL,
            'u_reference': self.U_GOAL,
        }
        if self.constraints is not None:
            info['symbolic_constraints'] = self.constraints.get_all_symbolic_models()
        return info
