'''MLP used by the actor and the critic (safe_control_gym/math_and_models/neural_networks.py:18-54).

The parameter names (`fcs.{i}.weight` / `fcs.{i}.bias`) are part of the checkpoint format (ppo.py:112-148), so
the shipped `.pt` models load unchanged.
'''

import torch.nn as nn
import torch.nn.functional as F


def get_activation(name):
    return getattr(F, name) if name else (lambda x: x)


class MLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dims=(), act='relu', output_act=None, init_weights=False, **kwargs):
        super().__init__()
        dims = [input_dim] + list(hidden_dims) + [output_dim]
        self.fcs = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        if init_weights:
            for fc in self.fcs:
                nn.init.orthogonal_(fc.weight.data, gain=1)
                nn.init.constant_(fc.bias.data, 0)
        self.act_name = act
        self.act = get_activation(act)
        self.output_act = get_activation(output_act)

    def forward(self, x):
        for fc in self.fcs[:-1]:
            x = self.act(fc(x))
        return self.output_act(self.fcs[-1](x))
