'''CudaVecEnv: N environments stepped by one kernel launch, behind the reference's VecEnv API.

Replaces DummyVecEnv / SubprocVecEnv (safe_control_gym/envs/env_wrappers/vectorized_env/dummy_vec_env.py:29-48,
subproc_vec_env.py:20-183): `reset() -> (obs, {'n': infos})`, `step(actions) -> (obs, rew, done, {'n': infos})` with
auto-reset and `terminal_observation` / `terminal_info`.  obs / rew / done are torch.cuda tensors; `info['n']` is a
lazy sequence that materialises per-env dicts (one device->host copy) only when somebody indexes it.

Seeding contract of make_env_fn (vectorized_env/__init__.py:31-35): env i draws from numpy Generator(seed + i).
In table mode each env's reset draws follow exactly the reference stream of its own generator (as long as no
per-step noise interleaves with them); in philox mode the draws come from the in-kernel counter-based streams.
'''

import numpy as np
import torch

from safe_control_gym_b200.engine import EngineHandle
from safe_control_gym_b200.envs import disturbances as dist_mod
from safe_control_gym_b200.envs.benchmark_env import Cost, Task
from safe_control_gym_b200.envs.env_wrappers.vectorized_env.vec_env import VecEnv


class LazyInfos:
    '''Sequence of per-env info dicts backed by batched device tensors.'''

    def __init__(self, owner, tensors, at_reset=False):
        self._owner, self._t, self._at_reset, self._host = owner, tensors, at_reset, None

    def __len__(self):
        return self._owner.num_envs

    def _materialise(self):
        if self._host is None:
            self._host = {k: v.detach().cpu().numpy() for k, v in self._t.items()}
        return self._host

    def __getitem__(self, i):
        h, env = self._materialise(), self._owner.proto
        if self._at_reset:
            info = {'current_step': 0, 'x_reference': env.X_GOAL, 'u_reference': env.U_GOAL, 'symbolic_model': None}
            return info

        def base(prefix=''):
            d = {'current_step': int(h['current_step'][i]), 'mse': float(h['mse'][i]), 'constraint_violation': int(h['cviol'][i])}
            if env.constraints is not None:
                d['constraint_values'] = h['cvals'][i].copy()
            if env.done_on_out_of_bound:
                d['out_of_bounds'] = bool(h['out_of_bounds'][i])
            if env.TASK == Task.STABILIZATION and env.COST == Cost.QUADRATIC:
                d['goal_reached'] = bool(h['goal_reached'][i])
            if h['current_step'][i] >= env.CTRL_STEPS:
                d['TimeLimit.truncated'] = bool(h['truncated'][i])
            return d
        if h['done'][i]:
            # dummy_vec_env.py:33-38: the returned info is the RESET info carrying the terminal transition
            return {'current_step': 0, 'terminal_observation': h['terminal_obs_f64'][i].copy(), 'terminal_info': base()}
        return base()

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class CudaVecEnv(VecEnv):
    def __init__(self, env_func, num_envs=1, seed=None, rng_mode='table', episodes_in_table=64, device=None, precision=None,
                 env_index_offset=0, total_envs=None):
        self.proto = env_func(seed=seed) if seed is not None else env_func()
        VecEnv.__init__(self, num_envs, self.proto.observation_space, self.proto.action_space)
        self.seed_value = 0 if seed is None else int(seed)
        self.rng_mode = rng_mode
        self.env_index_offset = int(env_index_offset)
        self.engine = EngineHandle(self.proto, num_envs=num_envs, auto_reset=True, rng_mode=rng_mode, seed=self.seed_value,
                                   env_index_offset=env_index_offset, device=device, precision=precision)
        self.device = self.engine.device
        self.actions = None
        if rng_mode == 'table':
            self._upload_tables(episodes_in_table)
        self.last = None

    def _upload_tables(self, K):
        env, N = self.proto, self.num_envs
        rows = []
        for i in range(N):
            env.seed(self.seed_value + self.env_index_offset + i)     # vectorized_env/__init__.py:31
            rows.append([env.draw_episode_row() for _ in range(K)])
        self.engine.set_episode_table(np.asarray(rows, dtype=np.float64).transpose(1, 0, 2))
        if dist_mod.noise_width(env.disturbances) > 0:
            raise NotImplementedError('per-step noise with rng_mode=\'table\' needs an explicit noise table: use '
                                      'engine.set_noise_table(...) or rng_mode=\'philox\'')

    # -- VecEnv API -----------------------------------------------------------------------------
    def reset(self):
        out = self.engine.reset()
        self.last = out
        return out['obs'], {'n': LazyInfos(self, {}, at_reset=True)}

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        out = self.engine.step(self.actions)
        self.last = out
        return out['obs'], out['rew'], out['done'].bool(), {'n': LazyInfos(self, out), 'batch': out}

    def close_extras(self):
        self.engine.close()

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.proto, attr_name) for _ in self._get_indices(indices)]

    def set_attr(self, attr_name, values, indices=None):
        raise NotImplementedError('per-env attributes cannot be changed after the batch was created')

    def env_method(self, method_name, method_args=None, method_kwargs=None, indices=None):
        if method_name == 'set_adversary_control':
            # rarl.py:366-367: env_method('set_adversary_control', [[act_i] for act_i in act_adv]) -> ONE batched call into the engine
            if indices is not None:
                raise NotImplementedError('set_adversary_control addresses the whole batch')
            first = method_args[0][0] if isinstance(method_args, (list, tuple)) else None
            if torch.is_tensor(method_args):
                batch = method_args
            elif torch.is_tensor(first):
                batch = torch.stack([a[0] for a in method_args])
            else:
                batch = np.stack([np.asarray(a[0], dtype=np.float32) for a in method_args])
            self.engine.set_adversary_control(batch)
            return [None] * self.num_envs
        idx = list(self._get_indices(indices))
        method_args = method_args or [[]] * len(idx)
        method_kwargs = method_kwargs or [{}] * len(idx)
        return [getattr(self.proto, method_name)(*a, **k) for a, k in zip(method_args, method_kwargs)]

    def get_env_random_state(self):
        '''Device env state + host generator state (dummy_vec_env.py:68-74 keeps python/numpy/torch RNG states).'''
        return [{'engine': self.engine.get_state(), 'np_random': self.proto.np_random.bit_generator.state}]

    def set_env_random_state(self, worker_random_states):
        st = worker_random_states[0]
        self.engine.set_state(st['engine'])
        self.proto.np_random.bit_generator.state = st['np_random']
