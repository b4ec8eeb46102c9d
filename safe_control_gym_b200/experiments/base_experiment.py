'''Batched evaluation harness: `BaseExperiment.run_evaluation` + `MetricExtractor` on the fused collector.

Counterpart of safe_control_gym/experiments/base_experiment.py:60-165 (`run_evaluation`, `_execute_evaluations` with `n_episodes`) and
:380-492 (`MetricExtractor`).  The reference steps ONE env episode after episode on the host and a `RecordDataWrapper` appends every
step's `info` dict to python lists; here the `n_episodes` evaluation episodes are `n_episodes` envs of one deterministic fused rollout
(`scg_rollout`, `deterministic = 1`: `ac.act()` = the mean action, ppo.py:233-238) of `CTRL_STEPS` steps, the per-step streams stay on
the device as `[T][n_episodes]` tensors with an "alive" mask (steps up to and including the first `done`), and the metrics are reductions
over them.  `trajs_data` keeps the reference's keys (`length`, `reward`, `mse`, `constraint_violation`, `obs`, `action`) as lists of
per-episode numpy arrays, so `MetricExtractor.compute_metrics(trajs_data)` also accepts data recorded by the reference itself.
'''

import numpy as np
import torch


def compute_cvar(data, alpha, lower_range=True):
    '''CVaR as the mean of the lower / upper alpha tail (math_and_models/metrics/performance_metrics.py:6-33).'''
    data = np.atleast_2d(data)
    _, n = data.shape
    s = np.sort(data)
    cvar = s[:, :int(alpha * n)].mean() if lower_range else s[:, -int(alpha * n):].mean()
    if np.all(np.isnan(cvar)):
        raise ValueError(f'Not enough samples to compute {alpha} CVaR from {data}')
    return cvar


class MetricExtractor:
    '''Metrics of recorded evaluation episodes (base_experiment.py:380-492): lengths, returns, RMSE from `info['mse']`, failure rate and
    number of constraint-violation steps.'''

    def compute_metrics(self, data, verbose=False):
        self.data, self.verbose = data, verbose
        lengths, rmse, viol = self.get_episode_lengths(), np.asarray(self.get_episode_rmse()), self.get_episode_constraint_violation_steps()
        return {
            'average_length': np.asarray(lengths).mean(),
            'length': lengths if len(lengths) > 1 else lengths[0],
            'average_return': np.asarray(self.get_episode_returns()).mean(),
            'average_rmse': rmse.mean(),
            'rmse': rmse if len(rmse) > 1 else rmse[0],
            'rmse_std': rmse.std(),
            'worst_case_rmse_at_0.5': compute_cvar(rmse, 0.5, lower_range=False),
            'failure_rate': np.asarray(self.get_episode_constraint_violations()).mean(),
            'average_constraint_violation': np.asarray(viol).mean(),
            'constraint_violation_std': np.asarray(viol).std(),
            'constraint_violation': np.asarray(viol) if len(viol) > 1 else viol[0],
        }

    def get_episode_data(self, key, postprocess_func=lambda x: x):
        if key in self.data:
            return [postprocess_func(v) for v in self.data[key]]
        if 'info' in self.data and key in self.data['info'][0][-1]:          # the reference's layout: lists of per-step info dicts
            return [postprocess_func([info[key] for info in ep if key in info]) for ep in self.data['info']]
        raise KeyError(f'Given data key \'{key}\' does not exist in recorded trajectory data.')

    def get_episode_lengths(self):
        return self.get_episode_data('length', postprocess_func=lambda x: int(np.sum(x)))

    def get_episode_returns(self):
        return self.get_episode_data('reward', postprocess_func=lambda x: float(np.sum(x)))

    def get_episode_rmse(self):
        return self.get_episode_data('mse', postprocess_func=lambda x: float(np.sqrt(np.mean(x))))

    def get_episode_constraint_violations(self):
        return self.get_episode_data('constraint_violation', postprocess_func=lambda x: float(np.any(x)))

    def get_episode_constraint_violation_steps(self):
        return self.get_episode_data('constraint_violation', postprocess_func=lambda x: float(np.sum(x)))


class BaseExperiment:
    '''Evaluation of an RL controller on the batched engine (base_experiment.py:19-165).

        exp = BaseExperiment(ctrl)                          # a safe_control_gym_b200 PPO controller (its env_func defines the task)
        trajs_data, metrics = exp.run_evaluation(n_episodes=64)
    '''

    def __init__(self, ctrl, seed=None):
        self.ctrl = ctrl
        self.metric_extractor = MetricExtractor()
        self.seed = ctrl.seed * 111 if seed is None else seed          # ppo.py:52: the evaluation env is seeded with seed * 111
        self._env = None

    def _eval_env(self, n):
        from safe_control_gym_b200.envs.env_wrappers.vectorized_env import make_vec_envs
        if self._env is None or self._env.num_envs != n:
            if self._env is not None:
                self._env.close()
            c = self.ctrl
            self._env = make_vec_envs(c.env_func, None, n, 1, self.seed, rng_mode=c.env.unwrapped.rng_mode, device=c.device.index, precision=c.precision)
        return self._env

    def run_evaluation(self, n_episodes=10, verbose=False, **kwargs):
        '''One deterministic fused rollout of `n_episodes` envs; returns (trajs_data, metrics) like base_experiment.py:49-88.'''
        c = self.ctrl
        c.agent.eval()
        env = self._eval_env(n_episodes)
        obs, _ = env.reset()
        blob, dims = c._policy_blob()
        T = int(env.proto.CTRL_STEPS)
        kw = {}
        if c.norm_obs:
            c.obs_normalizer.set_read_only()
            norm = c._norm_blob()
            norm[2 * dims[0] + 4] = 0.0                      # evaluation reports raw rewards
            kw = dict(norm=norm, obs0=c.obs_normalizer.normalize(obs).to(torch.float32).contiguous())
        out = env.engine.rollout(blob, dims, T, activation=c.activation, deterministic=True, want=('cviol', 'mse'), **kw)
        done = out['mask'][..., 0] == 0
        first = torch.where(done.any(0), done.float().argmax(0), torch.full((n_episodes,), T - 1, device=done.device))
        host = {k: out[k].cpu().numpy() for k in ('obs', 'act', 'rew', 'mse', 'cviol')}
        first = first.cpu().numpy()
        data = {'length': [], 'reward': [], 'mse': [], 'constraint_violation': [], 'obs': [], 'action': []}
        for i in range(n_episodes):
            L = int(first[i]) + 1
            data['length'].append(np.ones(L, dtype=np.int64))
            data['reward'].append(host['rew'][:L, i, 0].astype(np.float64))
            data['mse'].append(host['mse'][:L, i, 0].astype(np.float64))
            data['constraint_violation'].append(host['cviol'][:L, i, 0].astype(np.float64))
            data['obs'].append(host['obs'][:L, i])
            data['action'].append(host['act'][:L, i])
        metrics = self.metric_extractor.compute_metrics(data, verbose=verbose)
        if verbose:
            for k, v in metrics.items():
                print(f'{k}: {np.round(v, 3)}')
        return data, metrics

    def compute_metrics(self, trajs_data):
        return self.metric_extractor.compute_metrics(trajs_data)

    def close(self):
        if self._env is not None:
            self._env.close()
            self._env = None
