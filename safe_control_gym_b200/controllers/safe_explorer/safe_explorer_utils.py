'''Safety layer of the SafeExplorer controller (safe_control_gym/controllers/safe_explorer/safe_explorer_utils.py:15-299), device-resident.

One small MLP per state-constraint row predicts the sensitivity g_i(obs) of that constraint to the action
(`c_i' ~ c_i + g_i(obs) . a`, trained by regression, :93-118); at run time the mean action of the policy is projected with the closed-form
single-active-constraint solution of Dalal et al. 2018 (:120-176).  The projection is evaluated INSIDE the fused collector
(`scg_rollout_args.safety_*`, csrc/scg_kernels.cuh: safety_models / safety_apply); the torch version below is what the PPO update
differentiates through and what `select_action` uses.
'''

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from safe_control_gym_b200.math_and_models.neural_networks import MLP


class SafetyLayer:
    def __init__(self, obs_space, act_space, hidden_dim=64, num_constraints=1, lr=0.001, slack=None, device='cpu', **kwargs):
        self.num_constraints, self.device = num_constraints, device
        if isinstance(hidden_dim, int):
            hidden_dims = [hidden_dim]
        elif isinstance(hidden_dim, list):
            hidden_dims = hidden_dim
        else:
            raise ValueError('hidden_dim can only be int or list.')
        if len(hidden_dims) != 1:
            raise NotImplementedError('the fused collector evaluates constraint models with ONE hidden layer (the reference default)')
        self.obs_dim, self.act_dim, self.hidden = obs_space.shape[0], act_space.shape[0], hidden_dims[0]
        self.constraint_models = nn.ModuleList([MLP(self.obs_dim, self.act_dim, hidden_dims=hidden_dims) for _ in range(num_constraints)])
        assert slack is not None and isinstance(slack, (int, float, list))
        if isinstance(slack, (int, float)):
            slack = [slack] * num_constraints
        self.slack = np.array(slack, dtype=np.float64)
        self.optimizers = [torch.optim.Adam(m.parameters(), lr=lr) for m in self.constraint_models]

    def to(self, device):
        self.constraint_models.to(device)
        self.device = device

    def train(self):
        self.constraint_models.train()

    def eval(self):
        self.constraint_models.eval()

    def state_dict(self):
        return {'constraint_models': self.constraint_models.state_dict(), 'optimizers': [opt.state_dict() for opt in self.optimizers]}

    def load_state_dict(self, state_dict):
        self.constraint_models.load_state_dict(state_dict['constraint_models'])
        for i, sd in enumerate(state_dict['optimizers']):
            self.optimizers[i].load_state_dict(sd)

    def compute_loss(self, batch):
        '''L2 loss of the one-step constraint prediction per constraint (:93-107).'''
        obs, act, c, c_next = (batch[k].to(self.device) for k in ('obs', 'act', 'c', 'c_next'))
        losses = []
        for i, model in enumerate(self.constraint_models):
            g = model(obs)
            pred = c[:, i] + (g * act).sum(-1)
            losses.append(torch.mean((c_next[:, i] - pred) ** 2))
        return losses

    def update(self, batch):
        losses = self.compute_loss(batch)
        for loss, opt in zip(losses, self.optimizers):
            opt.zero_grad()
            loss.backward()
            opt.step()
        return {f'constraint_{i}_loss': float(loss.detach()) for i, loss in enumerate(losses)}

    def get_safe_action(self, obs, act, c):
        '''Action projection (:120-176), batched torch: obs (B,O), act (B,A), c (B,C) -> (B,A).'''
        g = torch.stack([model(obs) for model in self.constraint_models], 1)                 # (B,C,A)
        slack = torch.as_tensor(self.slack, dtype=act.dtype, device=act.device)
        numer = (g * act.unsqueeze(1)).sum(-1) + c + slack                                    # (B,C)
        denom = (g * g).sum(-1) + 1e-8
        mult = F.relu(numer / denom)
        max_mult, max_idx = torch.topk(mult, 1, dim=-1)                                        # (B,1)
        max_g = torch.gather(g, 1, max_idx.unsqueeze(-1).expand(-1, 1, g.shape[-1])).squeeze(1)
        return act - max_mult * max_g

    def pack(self, device):
        '''Constraint models -> the float32 blob of scg_rollout_args.d_safety ({W1[H][obs], b1[H], W2[A][H], b2[A]} per model).'''
        parts = []
        for m in self.constraint_models:
            parts += [m.fcs[0].weight.detach().reshape(-1), m.fcs[0].bias.detach(), m.fcs[1].weight.detach().reshape(-1), m.fcs[1].bias.detach()]
        return {'blob': torch.cat(parts).to(device=device, dtype=torch.float32).contiguous(), 'C': self.num_constraints, 'H': self.hidden,
                'slack': self.slack.tolist()}


class ConstraintBuffer:
    '''Replay storage of (obs, act, c, c_next) for the safety-layer regression (:179-299), on the device.'''

    def __init__(self, obs_dim, act_dim, num_constraints, max_size, device):
        self.max_size, self.device = int(max_size), device
        self.obs = torch.zeros(self.max_size, obs_dim, device=device)
        self.act = torch.zeros(self.max_size, act_dim, device=device)
        self.c = torch.zeros(self.max_size, num_constraints, device=device)
        self.c_next = torch.zeros(self.max_size, num_constraints, device=device)
        self.reset()

    def reset(self):
        self.pos, self.buffer_size = 0, 0

    def __len__(self):
        return self.buffer_size

    def push(self, batch):
        n = batch['obs'].shape[0]
        idx = (self.pos + torch.arange(n, device=self.device)) % self.max_size
        for k in ('obs', 'act', 'c', 'c_next'):
            getattr(self, k)[idx] = batch[k].to(self.device, torch.float32).reshape(n, -1)
        self.buffer_size = min(self.max_size, self.buffer_size + n)
        self.pos = (self.pos + n) % self.max_size

    def sampler(self, batch_size, generator=None, drop_last=True):
        perm = torch.randperm(len(self), device=self.device, generator=generator)
        n_full = len(self) // batch_size
        for i in range(n_full):
            idx = perm[i * batch_size:(i + 1) * batch_size]
            yield {k: getattr(self, k)[idx] for k in ('obs', 'act', 'c', 'c_next')}
        if not drop_last and len(self) % batch_size:
            idx = perm[n_full * batch_size:]
            yield {k: getattr(self, k)[idx] for k in ('obs', 'act', 'c', 'c_next')}
