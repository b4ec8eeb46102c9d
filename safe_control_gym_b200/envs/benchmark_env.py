'''BenchmarkEnv: the reference's env base class re-hosted on the CUDA engine.

API surface of safe_control_gym/envs/benchmark_env.py:42-502 (constructor kwargs = YAML keys, `seed`,
`reset`, `step` 4-tuple, `X_GOAL`, spaces, `constraints`, `disturbances`, action normalisation, the info
dict keys).  What differs is *where* the work happens: the constructor only resolves the configuration
into a flat `scg_config` + tables (reference trajectory, constraint rows, randomisation tables); `reset()`
and `step()` are C-ABI calls (`scg_reset` / `scg_step`) that run the hand-written sm_100a kernels on a
batch of one environment.  Batched use goes through `CudaVecEnv` / the fused PPO collector, which build a
handle for N envs from the same resolved configuration.

There is no CPU implementation of `step()`: without the CUDA library / a GPU, `reset()` raises.
'''

import copy
import os
from enum import Enum

import numpy as np

from safe_control_gym_b200 import _abi
from safe_control_gym_b200.envs import disturbances as dist_mod
from safe_control_gym_b200.envs.constraints import create_constraint_list
from safe_control_gym_b200.envs.disturbances import create_disturbance_list
from safe_control_gym_b200.spaces import Box


class Cost(str, Enum):
    RL_REWARD = 'rl_reward'
    QUADRATIC = 'quadratic'


class Task(str, Enum):
    STABILIZATION = 'stabilization'
    TRAJ_TRACKING = 'traj_tracking'


class Environment(str, Enum):
    CARTPOLE = 'cartpole'
    QUADROTOR = 'quadrotor'


def get_cost_weight_matrix(weights, dim):
    '''safe_control_gym/controllers/lqr/lqr_utils.py:77-99.'''
    if len(weights) == dim:
        return np.diag(weights)
    if len(weights) == 1:
        if isinstance(weights, list):
            return np.diag(weights * dim)
        if isinstance(weights, np.ndarray):
            return np.diag(weights.item() * np.ones(dim))
        raise Exception('Wrong type for cost weights.')
    raise Exception('Wrong dimension for cost weights.')


class BenchmarkEnv:
    '''Benchmark environment base class (engine-backed).'''

    _count = 0
    NAME = 'base'
    SYSTEM = None                     # SCG_SYS_*
    AVAILABLE_CONSTRAINTS = None
    DISTURBANCE_MODES = None
    INERTIAL_PROP_RAND_INFO = None
    INIT_STATE_RAND_INFO = None
    TASK_INFO = None

    def __init__(self,
                 output_dir=None, seed=None, gui=False, verbose=False, normalized_rl_action_space=False,
                 task=Task.STABILIZATION, task_info=None, cost=Cost.RL_REWARD,
                 pyb_freq=50, ctrl_freq=50, episode_len_sec=5,
                 init_state=None, randomized_init=True, init_state_randomization_info=None,
                 prior_prop=None, inertial_prop=None, randomized_inertial_prop=False, inertial_prop_randomization_info=None,
                 constraints=None, done_on_violation=False, use_constraint_penalty=False, constraint_penalty=1.0,
                 disturbances=None, adversary_disturbance=None, adversary_disturbance_offset=0.0, adversary_disturbance_scale=0.01,
                 # engine options (new)
                 device=0, precision='f64',
                 **kwargs):
        self.idx = self.__class__._count
        self.__class__._count += 1
        self.output_dir = os.getcwd() if output_dir is None else output_dir
        if gui:
            raise NotImplementedError('the CUDA engine is headless (gui=True is PyBullet-only)')
        self.GUI = False
        self.VERBOSE = verbose
        self.TASK = Task(task)
        if task_info is not None:
            self.TASK_INFO = task_info
        # timing constants (benchmark_env.py:139-148)
        self.CTRL_FREQ = ctrl_freq
        self.PYB_FREQ = pyb_freq
        if self.PYB_FREQ % self.CTRL_FREQ != 0:
            raise ValueError('[ERROR] in BenchmarkEnv.__init__(), pyb_freq is not divisible by env_freq.')
        self.PYB_STEPS_PER_CTRL = int(self.PYB_FREQ / self.CTRL_FREQ)
        self.CTRL_TIMESTEP = 1. / self.CTRL_FREQ
        self.PYB_TIMESTEP = 1. / self.PYB_FREQ
        self.EPISODE_LEN_SEC = episode_len_sec
        self.CTRL_STEPS = self.EPISODE_LEN_SEC * self.CTRL_FREQ
        self.INIT_STATE = init_state
        self.RANDOMIZED_INIT = randomized_init
        if init_state_randomization_info is not None:
            self.INIT_STATE_RAND_INFO = init_state_randomization_info
        self.PRIOR_PROP = prior_prop
        self.INERTIAL_PROP = inertial_prop
        self.RANDOMIZED_INERTIAL_PROP = randomized_inertial_prop
        if inertial_prop_randomization_info is not None:
            self.INERTIAL_PROP_RAND_INFO = inertial_prop_randomization_info
        self.NORMALIZED_RL_ACTION_SPACE = normalized_rl_action_space
        self.COST = Cost(cost)
        self._set_action_space()
        self._set_observation_space()
        self.action_dim = self.action_space.shape[0]
        self.obs_dim = self.observation_space.shape[0]
        self.state_dim = self.state_space.shape[0] if hasattr(self, 'state_space') else self.obs_dim
        self.CONSTRAINTS = constraints
        self.DONE_ON_VIOLATION = done_on_violation
        self.use_constraint_penalty = use_constraint_penalty
        self.constraint_penalty = constraint_penalty
        self._setup_constraints()
        self.DISTURBANCES = disturbances
        self.adversary_disturbance = adversary_disturbance       # None | 'action' | 'dynamics' (benchmark_env.py:184-186)
        self.adv_action = None
        self.adversary_disturbance_offset = adversary_disturbance_offset
        self.adversary_disturbance_scale = adversary_disturbance_scale
        self._setup_disturbances()
        # engine
        self.device = int(str(device).replace('cuda:', '')) if not isinstance(device, int) else device
        if precision not in ('f64', 'f32'):
            raise ValueError('precision must be \'f64\' or \'f32\'')
        self.precision = precision
        self._handle = None
        self.SEED = seed
        self.seed(seed)
        self.initial_reset = False
        self.at_reset = False
        self.state = None
        self.current_raw_action = None
        self.current_physical_action = None
        self.current_noisy_physical_action = None
        self.current_clipped_action = None
        self.symbolic = None   # CasADi prior model: out of scope

    # ------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        '''benchmark_env.py:193-214: one numpy Generator shared by env and disturbances.'''
        if seed is not None and not (isinstance(seed, (int, np.integer)) and seed >= 0):
            raise ValueError(f'Seed must be a non-negative integer or omitted, not {seed}')
        seed_seq = np.random.SeedSequence(seed)                      # gymnasium.utils.seeding.np_random
        self.np_random = np.random.Generator(np.random.PCG64(seed_seq))
        np_seed = seed_seq.entropy
        self.action_space.seed(np_seed if seed is None else seed)
        for _, disturbs in self.disturbances.items():
            disturbs.seed(self)
        return [np_seed]

    def set_adversary_control(self, action):
        '''benchmark_env.py:216-228: the adversary's action for the next step (clip to the adversary action space, scale, offset -- done
        by the engine in float32 like the reference's numpy expression).'''
        if self.adversary_disturbance is None:
            raise RuntimeError('[ERROR] adversary_disturbance does not exist, env.set_adversary_control() cannot be called.')
        action = np.asarray(action, dtype=np.float32).reshape(1, -1)
        clipped = np.clip(action[0], self.adversary_action_space.low, self.adversary_action_space.high)
        self.adv_action = clipped * self.adversary_disturbance_scale + self.adversary_disturbance_offset
        self._engine().set_adversary_control(action)

    def _check_initial_reset(self):
        if not self.initial_reset:
            raise RuntimeError('[ERROR] You must call env.reset() at least once before using env.step().')

    def _randomize_values_by_info(self, original_values, randomization_info):
        '''benchmark_env.py:237-268 (additive).  Returns (randomized_values, additive_samples).'''
        randomized_values = copy.deepcopy(original_values)
        rand_info_copy = copy.deepcopy(randomization_info)
        samples = {k: 0.0 for k in original_values}
        for key in original_values:
            if key in rand_info_copy:
                distrib = getattr(self.np_random, rand_info_copy[key].pop('distrib'))
                d_args = rand_info_copy[key].pop('args', [])
                d_kwargs = rand_info_copy[key]
                s = distrib(*d_args, **d_kwargs)
                samples[key] = float(s)
                randomized_values[key] += s
        return randomized_values, samples

    def _setup_disturbances(self):
        self.disturbances = {}
        self._dist_reset_order = []
        if self.DISTURBANCES is not None:
            for mode, disturb_specs in self.DISTURBANCES.items():
                assert mode in self.DISTURBANCE_MODES, '[ERROR] in BenchmarkEnv._setup_disturbances(), disturbance mode not available.'
                self.disturbances[mode] = create_disturbance_list(disturb_specs, self.DISTURBANCE_MODES[mode], self)
                self._dist_reset_order.append(mode)
        # adversary disturbance, set from outside the env (benchmark_env.py:288-295): its action space is [-1, 1]^dim of the mode it drives
        if self.adversary_disturbance is not None:
            if self.adversary_disturbance not in ('action', 'dynamics'):
                raise NotImplementedError('adversary_disturbance must be \'action\' or \'dynamics\' (the observation mode is never read by the reference envs)')
            dim = self.DISTURBANCE_MODES[self.adversary_disturbance]['dim']
            self.adversary_action_space = Box(low=-np.ones(dim), high=np.ones(dim), dtype=np.float32)
            self.adversary_observation_space = self.observation_space

    def _setup_constraints(self):
        self.constraints = None
        self.num_constraints = 0
        if self.CONSTRAINTS is not None:
            self.constraints = create_constraint_list(self.CONSTRAINTS, self.AVAILABLE_CONSTRAINTS, self)
            self.num_constraints = self.constraints.num_constraints

    # ------------------------------------------------------------------------------------------
    # reference trajectories (benchmark_env.py:504-713)
    # ------------------------------------------------------------------------------------------
    def _generate_trajectory(self, traj_type='figure8', traj_length=10.0, num_cycles=1, traj_plane='xy',
                             position_offset=np.array([0, 0]), scaling=1.0, sample_time=0.01):
        if traj_type not in ('circle', 'square', 'figure8'):
            raise ValueError('Trajectory type should be one of [circle, square, figure8].')
        axes = ['x', 'y', 'z']
        if not (len(traj_plane) == 2 and traj_plane[0] in axes and traj_plane[1] in axes and traj_plane[0] != traj_plane[1]):
            raise ValueError('Trajectory plane should be in form of ab, where a and b can be {x, y, z}.')
        ia, ib = axes.index(traj_plane[0]), axes.index(traj_plane[1])
        period = traj_length / num_cycles
        times = np.arange(0, traj_length + sample_time, sample_time)   # one sample longer than the episode (:543)
        pos = np.zeros((len(times), 3))
        vel = np.zeros((len(times), 3))
        shape_fn = {'figure8': self._figure8, 'circle': self._circle, 'square': self._square}[traj_type]
        for i, t in enumerate(times):
            a, b, a_dot, b_dot = shape_fn(t, period, scaling)
            pos[i, ia], pos[i, ib] = a + position_offset[0], b + position_offset[1]
            vel[i, ia], vel[i, ib] = a_dot, b_dot
        speed = np.linalg.norm(vel, axis=1, keepdims=True)
        return pos, vel, speed

    @staticmethod
    def _figure8(t, traj_period, scaling):
        w = 2.0 * np.pi / traj_period
        s, c = np.sin(w * t), np.cos(w * t)
        return scaling * s, scaling * s * c, scaling * w * c, scaling * w * (c**2 - s**2)

    @staticmethod
    def _circle(t, traj_period, scaling):
        w = 2.0 * np.pi / traj_period
        s, c = np.sin(w * t), np.cos(w * t)
        return scaling * c, scaling * s, -scaling * w * s, scaling * w * c

    @staticmethod
    def _square(t, traj_period, scaling):
        seg_period = traj_period / 4.0
        speed = scaling / seg_period
        cycle_time = t % traj_period
        seg_time = cycle_time % seg_period
        seg = int(np.floor(cycle_time / seg_period))
        p = speed * seg_time
        if seg == 0:
            return 0.0, p, 0.0, speed
        if seg == 1:
            return -p, scaling, -speed, 0.0
        if seg == 2:
            return -scaling, scaling - p, 0.0, -speed
        return -scaling + p, 0.0, speed, 0.0

    # ------------------------------------------------------------------------------------------
    # configuration -> engine
    # ------------------------------------------------------------------------------------------
    def _x_goal_table(self):
        xg = np.atleast_2d(np.asarray(self.X_GOAL, dtype=np.float64))
        return np.ascontiguousarray(xg)

    def _fill_system_config(self, cfg):
        '''System-specific fields of scg_config (implemented by CartPole / Quadrotor).'''
        raise NotImplementedError

    def _inertial_keys(self):
        raise NotImplementedError

    def _init_keys(self):
        raise NotImplementedError

    def build_config(self, num_envs=1, auto_reset=False, env_index_offset=0, rng_mode='table', seed=0, device=None, precision=None):
        '''Resolves this env into the flat POD the C-ABI takes.'''
        cfg = _abi.new_config()
        cfg.system = self.SYSTEM
        cfg.num_envs = int(num_envs)
        cfg.env_index_offset = int(env_index_offset)
        cfg.device = self.device if device is None else int(device)
        cfg.precision = _abi.SCG_PREC_F64 if (precision or self.precision) == 'f64' else _abi.SCG_PREC_F32
        cfg.rng_mode = _abi.SCG_RNG_TABLE if rng_mode == 'table' else _abi.SCG_RNG_PHILOX
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.ctrl_freq, cfg.pyb_freq, cfg.ctrl_steps = int(self.CTRL_FREQ), int(self.PYB_FREQ), int(self.CTRL_STEPS)
        cfg.task = _abi.SCG_TASK_STABILIZATION if self.TASK == Task.STABILIZATION else _abi.SCG_TASK_TRAJ_TRACKING
        cfg.cost = _abi.SCG_COST_RL_REWARD if self.COST == Cost.RL_REWARD else _abi.SCG_COST_QUADRATIC
        cfg.obs_goal_horizon = int(self.obs_goal_horizon)
        cfg.obs_wrap_angle = int(bool(getattr(self, 'obs_wrap_angle', False)))
        cfg.normalized_action = int(bool(self.NORMALIZED_RL_ACTION_SPACE))
        cfg.rew_exponential = int(bool(self.rew_exponential))
        cfg.done_on_out_of_bound = int(bool(self.done_on_out_of_bound))
        cfg.done_on_violation = int(bool(self.DONE_ON_VIOLATION))
        cfg.use_constraint_penalty = int(bool(self.use_constraint_penalty))
        cfg.auto_reset = int(bool(auto_reset))
        cfg.constraint_penalty = float(self.constraint_penalty)
        cfg.goal_tolerance = float(self.TASK_INFO['stabilization_goal_tolerance']) if self.TASK == Task.STABILIZATION else 0.0
        nx, nu = self.state_dim, self.action_dim
        rsw = np.broadcast_to(self.rew_state_weight, (nx,)) if len(self.rew_state_weight) in (1, nx) else None
        raw = np.broadcast_to(self.rew_act_weight, (nu,)) if len(self.rew_act_weight) in (1, nu) else None
        if rsw is None or raw is None:
            raise ValueError('rew_state_weight / rew_act_weight have the wrong dimension')
        for k in range(nx):
            cfg.rew_state_weight[k] = float(rsw[k])
            cfg.mse_weight[k] = float(self.info_mse_metric_state_weight[k])
            cfg.Q[k] = float(self.Q[k, k])
            cfg.state_low[k] = float(self.state_space.low[k])      # float32 bounds promoted (quadrotor.py:697)
            cfg.state_high[k] = float(self.state_space.high[k])
        for k in range(nu):
            cfg.rew_act_weight[k] = float(raw[k])
            cfg.R[k] = float(self.R[k, k])
            cfg.u_goal[k] = float(np.atleast_1d(self.U_GOAL)[k])
            cfg.act_low[k] = float(self.physical_action_bounds[0][k])
            cfg.act_high[k] = float(self.physical_action_bounds[1][k])
        self._fill_system_config(cfg)
        cfg.randomized_init = int(bool(self.RANDOMIZED_INIT))
        cfg.randomized_inertial = int(bool(self.RANDOMIZED_INERTIAL_PROP))
        for j, key in enumerate(self._init_keys()):
            kind, a, b = self._philox_distribution(self.INIT_STATE_RAND_INFO, key, self.RANDOMIZED_INIT and rng_mode != 'table')
            cfg.init_rand_kind[j], cfg.init_rand_a[j], cfg.init_rand_b[j] = kind, a, b
        for j, key in enumerate(self._inertial_keys()):
            kind, a, b = self._philox_distribution(self.INERTIAL_PROP_RAND_INFO, key, self.RANDOMIZED_INERTIAL_PROP and rng_mode != 'table')
            cfg.inertial_rand_kind[j], cfg.inertial_rand_a[j], cfg.inertial_rand_b[j] = kind, a, b
        dist_mod.fill_config(cfg, self.disturbances)
        cfg.adversary_mode = {None: _abi.SCG_ADV_NONE, 'action': _abi.SCG_ADV_ACTION, 'dynamics': _abi.SCG_ADV_DYNAMICS}[self.adversary_disturbance]
        cfg.adversary_scale, cfg.adversary_offset = float(self.adversary_disturbance_scale), float(self.adversary_disturbance_offset)
        return cfg

    @staticmethod
    def _philox_distribution(info, key, active):
        '''Maps a reference randomisation entry onto the in-kernel distributions.'''
        if not active or info is None or key not in info:
            return _abi.SCG_RAND_NONE, 0.0, 0.0
        spec = dict(info[key])
        name = spec.pop('distrib')
        args = list(spec.pop('args', []))
        if name == 'uniform':
            low = args[0] if len(args) > 0 else spec.get('low', 0.0)
            high = args[1] if len(args) > 1 else spec.get('high', 1.0)
            return _abi.SCG_RAND_UNIFORM, float(low), float(high)
        if name == 'normal':
            loc = args[0] if len(args) > 0 else spec.get('loc', 0.0)
            scale = args[1] if len(args) > 1 else spec.get('scale', 1.0)
            return _abi.SCG_RAND_NORMAL, float(loc), float(scale)
        raise NotImplementedError(f'distribution \'{name}\' is only available with rng_mode=\'table\' (host-drawn tables)')

    def draw_episode_offsets(self):
        return dist_mod.draw_episode_offsets(self.disturbances, self._dist_reset_order, self.np_random)

    def draw_episode_row(self, offs=None):
        '''One episode's random draws from `self.np_random` in the reference's order
        (disturbance offsets -> inertial -> initial state; SURVEY.md appendix A "RNG streams").'''
        if offs is None:
            offs = self.draw_episode_offsets()
        inertial = [0.0] * _abi.SCG_MAX_INERTIAL
        if self.RANDOMIZED_INERTIAL_PROP:
            keys = self._inertial_keys()
            vals = {k: 0.0 for k in keys}
            _, smp = self._randomize_values_by_info(vals, self.INERTIAL_PROP_RAND_INFO)
            for j, k in enumerate(keys):
                inertial[j] = smp[k]
        init = [0.0] * self.state_dim
        if self.RANDOMIZED_INIT:
            keys = self._init_keys()
            vals = {k: 0.0 for k in keys}
            _, smp = self._randomize_values_by_info(vals, self.INIT_STATE_RAND_INFO)
            for j, k in enumerate(keys):
                init[j] = smp[k]
        return offs + inertial + init

    def draw_step_row(self):
        return dist_mod.draw_step_noise(self.disturbances, self.np_random)

    # ------------------------------------------------------------------------------------------
    # gym API on a batch of one
    # ------------------------------------------------------------------------------------------
    def _engine(self):
        if self._handle is None:
            from safe_control_gym_b200.engine import EngineHandle
            self._handle = EngineHandle(self, num_envs=1, auto_reset=False, rng_mode='table', seed=0)
        return self._handle

    def before_reset(self, seed=None):
        '''benchmark_env.py:320-341: disturbance reset draws happen BEFORE an optional re-seed.'''
        self._pending_offsets = self.draw_episode_offsets()
        self.initial_reset = True
        self.at_reset = True
        self.pyb_step_counter = 0
        self.ctrl_step_counter = 0
        self.current_raw_action = None
        self.current_physical_action = None
        self.current_noisy_physical_action = None
        self.current_clipped_action = None
        self.adv_action = None                                    # benchmark_env.py:338-339
        if seed is not None:
            self.seed(seed)

    def reset(self, seed=None):
        '''(Re-)initialises the environment; returns (obs, info) like quadrotor.py:328-392.'''
        self.before_reset(seed=seed)
        eng = self._engine()
        row = np.asarray(self.draw_episode_row(self._pending_offsets), dtype=np.float64)
        inertial = row[len(row) - self.state_dim - _abi.SCG_MAX_INERTIAL: len(row) - self.state_dim]
        self._check_inertial(inertial)
        eng.set_episode_table(row.reshape(1, 1, -1), restart=True)
        out = eng.reset()
        self.state = out['state_f64'][0].cpu().numpy().copy()
        obs = out['obs_f64'][0].cpu().numpy().copy()
        info = self._get_reset_info(inertial)
        info['current_step'] = 0
        if self.constraints is not None and not (self.constraints.state_constraints == []):
            info['constraint_values'] = self.constraints.get_values(self, only_state=True)
        self.at_reset = False
        return obs, info

    def _check_inertial(self, additive):
        return

    def before_step(self, action):
        self._check_initial_reset()
        action = np.atleast_1d(np.squeeze(action))
        if action.ndim != 1:
            raise ValueError('[ERROR]: The action returned by the controller must be 1 dimensional.')
        self.current_raw_action = action
        return action

    def step(self, action):
        '''Advances one control step on the GPU; returns (obs, rew, done, info) (quadrotor.py:394-445).'''
        action = self.before_step(action)
        eng = self._engine()
        if dist_mod.noise_width(self.disturbances) > 0:
            eng.set_noise_table(np.asarray(self.draw_step_row(), dtype=np.float64).reshape(1, 1, -1))
        out = eng.step(np.asarray(action, dtype=np.float64).reshape(1, -1))
        host = {k: v[0].cpu().numpy() for k, v in out.items()}
        self.state = host['state_f64'].copy()
        self.current_physical_action = self.denormalize_action(action)
        self.pyb_step_counter += self.PYB_STEPS_PER_CTRL
        self.ctrl_step_counter += 1
        if self.adversary_disturbance == 'dynamics':
            self.adv_action = None                                # consumed by the step (quadrotor.py:424-425, cartpole.py:550-551)
        info = {}
        if self.TASK == Task.STABILIZATION and self.COST == Cost.QUADRATIC:
            info['goal_reached'] = bool(host['goal_reached'])
        if self.done_on_out_of_bound:
            info['out_of_bounds'] = bool(host['out_of_bounds'])
        info['mse'] = float(host['mse'])
        info['current_step'] = int(host['current_step'])
        if self.constraints is not None:
            info['constraint_values'] = host['cvals'].copy()
        info['constraint_violation'] = int(host['cviol'])
        done = bool(host['done'])
        if self.ctrl_step_counter >= self.CTRL_STEPS:
            info['TimeLimit.truncated'] = bool(host['truncated'])
        return host['obs_f64'].copy(), float(host['rew_f64']), done, info

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def render(self, mode='human'):
        raise NotImplementedError('rendering is PyBullet-only; the CUDA engine is headless')
