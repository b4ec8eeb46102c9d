'''Episode statistics for a CudaVecEnv (record_episode_statistics.py:92-166), accumulated on the device.'''

from collections import deque

import torch

from safe_control_gym_b200.envs.env_wrappers.vectorized_env.vec_env import VecEnvWrapper


class VecRecordEpisodeStatistics(VecEnvWrapper):
    def __init__(self, venv, deque_size=None, **kwargs):
        super().__init__(venv)
        self.deque_size = deque_size
        dev = venv.device
        self.episode_return = torch.zeros(self.num_envs, dtype=torch.float64, device=dev)
        self.episode_length = torch.zeros(self.num_envs, dtype=torch.float64, device=dev)
        self.return_queue = deque(maxlen=deque_size)
        self.length_queue = deque(maxlen=deque_size)
        self.episode_stats, self.accumulated_stats, self.queued_stats = {}, {}, {}

    def add_tracker(self, name, init_value, mode='accumulate'):
        self.episode_stats[name] = torch.zeros(self.num_envs, dtype=torch.float64, device=self.venv.device)
        if mode == 'accumulate':
            self.accumulated_stats[name] = init_value
        elif mode == 'queue':
            self.queued_stats[name] = deque(maxlen=self.deque_size)
        else:
            raise Exception('Tracker mode not implemented.')

    def reset(self, **kwargs):
        self.episode_return.zero_()
        self.episode_length.zero_()
        for v in self.episode_stats.values():
            v.zero_()
        return self.venv.reset(**kwargs)

    def step_wait(self):
        obs, rew, done, info = self.venv.step_wait()
        batch = info['batch']
        self.episode_return += rew.double()
        self.episode_length += 1
        tracked = {'constraint_violation': batch['cviol'].double(), 'mse': batch['mse']}
        for key, acc in self.episode_stats.items():
            if key in tracked:
                acc += tracked[key]
        if bool(done.any()):
            idx = torch.nonzero(done).flatten()
            self.return_queue.extend(self.episode_return[idx].tolist())
            self.length_queue.extend(self.episode_length[idx].tolist())
            for key, acc in self.episode_stats.items():
                vals = acc[idx].tolist()
                if key in self.accumulated_stats:
                    self.accumulated_stats[key] += sum(vals)
                if key in self.queued_stats:
                    self.queued_stats[key].extend(vals)
                acc[idx] = 0
            self.episode_return[idx] = 0
            self.episode_length[idx] = 0
        return obs, rew, done, info

    def record_rollout(self, out):
        '''Feeds the per-env episode sums of a fused rollout (scg_rollout_out.d_ep_*) into the trackers.'''
        cnt = out['ep_count']
        n = int(cnt.sum())
        if n == 0:
            return
        nz = cnt > 0
        c = cnt[nz].double()
        self.return_queue.extend((out['ep_return_sum'][nz].double() / c).tolist())
        self.length_queue.extend((out['ep_length_sum'][nz].double() / c).tolist())
        total_cv = float(out['ep_cviol_sum'].sum())
        if 'constraint_violation' in self.accumulated_stats:
            self.accumulated_stats['constraint_violation'] += total_cv
        if 'constraint_violation' in self.queued_stats:
            self.queued_stats['constraint_violation'].extend((out['ep_cviol_sum'][nz].double() / c).tolist())
