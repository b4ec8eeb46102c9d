'''Normalisers of the RL controllers (safe_control_gym/math_and_models/normalization.py:13-169), device-resident.

`RunningMeanStd` keeps float64 mean / var / count and merges batches with the parallel-variance update of the reference
(normalization.py:42-56).  Inside the fused collector the statistics are FROZEN for the duration of one rollout (the kernel applies
`clip((x - mean) / sqrt(var + eps))` / `clip(r / sqrt(var + eps))` with the values it was launched with); after the rollout the
controller feeds the raw streams the kernel recorded:

* observations: one merge of all T x N raw observations.  The reference merges T batches of N, one per step; the merge is exactly
  associative, so the statistics after the rollout are the same up to float64 rounding.
* rewards: `RewardStdNormalizer` tracks the variance of the running discounted return `ret = ret * gamma + rew` of every env
  (normalization.py:148-159), INCLUDING the reference's quirk `self.ret[dones.astype(np.long)] = 0`, which indexes with 0 / 1 instead of
  masking: env 0 is zeroed whenever some env is not done and env 1 whenever some env is done, no other env is ever reset
  (SURVEY.md appendix B.15).  `update_from_rollout` reproduces that exactly: the untouched envs are a discounted cumulative sum
  (vectorised), envs 0 / 1 are stepped explicitly, and the T per-step batch moments are merged in order.
'''

import numpy as np
import torch


def normalize_angle(x):
    '''Wraps an angle to [-pi, pi) (normalization.py:8-10).'''
    return ((x + np.pi) % (2 * np.pi)) - np.pi


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=(), device=None):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = float(epsilon)

    def update(self, arr):
        arr = torch.as_tensor(arr, dtype=torch.float64, device=self.mean.device)
        self.update_from_moments(arr.mean(0), arr.var(0, unbiased=False), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + delta.square() * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class BaseNormalizer:
    def __init__(self, read_only=False):
        self.read_only = read_only

    def set_read_only(self):
        self.read_only = True

    def unset_read_only(self):
        self.read_only = False

    def __call__(self, x, *args, **kwargs):
        return x

    def state_dict(self):
        return {}

    def load_state_dict(self, _):
        pass


class MeanStdNormalizer(BaseNormalizer):
    def __init__(self, shape=(), read_only=False, clip=10.0, epsilon=1e-8, device=None):
        super().__init__(read_only)
        self.rms = RunningMeanStd(shape=shape, device=device)
        self.clip, self.epsilon = clip, epsilon

    def normalize(self, x):
        x64 = torch.as_tensor(x, dtype=torch.float64, device=self.rms.mean.device)
        return torch.clamp((x64 - self.rms.mean) / torch.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)

    def __call__(self, x):
        if not self.read_only:
            self.rms.update(x)
        return self.normalize(x)

    def state_dict(self):
        return {'mean': self.rms.mean.cpu().numpy(), 'var': self.rms.var.cpu().numpy()}      # numpy, like the reference's checkpoints

    def load_state_dict(self, saved):
        if saved:
            dev = self.rms.mean.device
            self.rms.mean = torch.as_tensor(np.asarray(saved['mean']), dtype=torch.float64, device=dev)
            self.rms.var = torch.as_tensor(np.asarray(saved['var']), dtype=torch.float64, device=dev)


class RewardStdNormalizer(MeanStdNormalizer):
    def __init__(self, gamma=0.99, read_only=False, clip=10.0, epsilon=1e-8, device=None):
        super().__init__((), read_only, clip, epsilon, device)
        self.gamma = gamma
        self.ret = None

    def scale(self):
        return 1.0 / float(torch.sqrt(self.rms.var + self.epsilon))

    def update_from_rollout(self, rew, done):
        '''rew, done: [T][N] device tensors of one rollout (raw rewards, done flags 0/1).'''
        if self.read_only:
            return
        T, N = rew.shape
        dev = rew.device
        r = rew.to(torch.float64)
        if self.ret is None:
            self.ret = torch.zeros(N, dtype=torch.float64, device=dev)
        g = self.gamma
        # envs the quirk never touches: ret_t = g^(t+1) ret_prev + sum_{s<=t} g^(t-s) r_s, as a scaled cumulative sum (g^-T stays small: T ~ 1e2-1e3)
        pw = torch.pow(torch.tensor(g, dtype=torch.float64, device=dev), torch.arange(T, dtype=torch.float64, device=dev))      # g^t
        rets = pw[:, None] * (torch.cumsum(r / pw[:, None], 0) + g * self.ret[None, :])
        # envs 0 and 1: explicit recursion with the reference's integer-indexing reset
        any_done = (done != 0).any(1).cpu().numpy()
        any_live = (done == 0).any(1).cpu().numpy()
        r01 = r[:, :2].cpu().numpy()
        cur = self.ret[:2].cpu().numpy().copy()
        fix = np.zeros((T, min(N, 2)))
        for t in range(T):
            cur = cur * g + r01[t]
            fix[t] = cur
            if any_live[t]:
                cur[0] = 0.0
            if N > 1 and any_done[t]:
                cur[1] = 0.0
        rets[:, :fix.shape[1]] = torch.as_tensor(fix, dtype=torch.float64, device=dev)
        means, vars_ = rets.mean(1).cpu().numpy(), rets.var(1, unbiased=False).cpu().numpy()
        for t in range(T):              # T scalar merges, in order
            self.rms.update_from_moments(torch.tensor(means[t], dtype=torch.float64, device=dev), torch.tensor(vars_[t], dtype=torch.float64, device=dev), N)
        self.ret = rets[-1].clone()
        self.ret[:fix.shape[1]] = torch.as_tensor(cur, dtype=torch.float64, device=dev)

    def __call__(self, x, dones):
        '''Step-wise use (gym-style collector): one batch of rewards [N] and done flags.'''
        x = torch.as_tensor(x, dtype=torch.float64, device=self.rms.mean.device)
        if not self.read_only:
            self.update_from_rollout(x[None].to(torch.float32).to(torch.float64), torch.as_tensor(dones, device=x.device)[None])
        return torch.clamp(x / torch.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)
