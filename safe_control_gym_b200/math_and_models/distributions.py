'''Diagonal Gaussian policy head (safe_control_gym/math_and_models/distributions.py:9-33).'''

import torch


class Normal(torch.distributions.Normal):
    def log_prob(self, actions):
        return super().log_prob(actions).sum(-1, keepdim=True)

    def entropy(self):
        return super().entropy().sum(-1)

    def mode(self):
        return self.mean
