'''SAC agent pieces (safe_control_gym/controllers/sac/sac_utils.py) with everything resident on the GPU.

Module trees keep the reference's parameter names (`actor.net.fcs.*`, `actor.mu_layer`, `actor.log_std_layer`, `q1.q_net.fcs.*`,
`q2.q_net.fcs.*`), so checkpoints are interchangeable.  The replay buffer is the SACBuffer scheme (sac_utils.py:283-330: obs, next_obs,
act, rew, mask) as torch.cuda tensors: `push` is one indexed store per key, `sample` one device-side gather -- no host copy on either side
(the reference keeps numpy arrays on the host and copies every sampled batch to the device).
'''

import copy
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from safe_control_gym_b200.math_and_models.neural_networks import MLP


class DeviceReplayBuffer:
    '''Ring buffer in HBM with the SACBuffer / DDPGBuffer scheme.'''

    KEYS = ('obs', 'next_obs', 'act', 'rew', 'mask')

    def __init__(self, obs_space, act_space, max_size, batch_size=None, device='cuda'):
        self.max_size, self.batch_size, self.device = int(max_size), batch_size, torch.device(device)
        widths = {'obs': obs_space.shape[0], 'next_obs': obs_space.shape[0], 'act': act_space.shape[0], 'rew': 1, 'mask': 1}
        self.data = {k: torch.zeros(self.max_size, w, dtype=torch.float32, device=self.device) for k, w in widths.items()}
        self.data['mask'].fill_(1.0)                                  # sac_utils.py:313-316
        self.pos, self.buffer_size = 0, 0

    def __len__(self):
        return self.buffer_size

    def push(self, batch):
        '''batch: {key: [n][width] device tensors}; wraps around like sac_utils.py:343-362.'''
        n = next(iter(batch.values())).shape[0]
        idx = (self.pos + torch.arange(n, device=self.device)) % self.max_size
        for k, v in batch.items():
            self.data[k][idx] = v.to(self.device, torch.float32).reshape(n, -1)
        self.buffer_size = min(self.max_size, max(self.buffer_size, self.pos + n))
        self.pos = (self.pos + n) % self.max_size

    def sample(self, batch_size=None, generator=None):
        bs = batch_size or self.batch_size
        idx = torch.randint(0, len(self), (bs,), device=self.device, generator=generator)
        return {k: v[idx] for k, v in self.data.items()}

    def state_dict(self):
        st = {'pos': self.pos, 'buffer_size': self.buffer_size}
        st.update({k: v.cpu().numpy() for k, v in self.data.items()})  # the reference stores numpy arrays (sac_utils.py:332-341)
        return st

    def load_state_dict(self, state):
        self.pos, self.buffer_size = int(state['pos']), int(state['buffer_size'])
        for k in self.KEYS:
            self.data[k].copy_(torch.as_tensor(state[k]).reshape(self.data[k].shape))


def soft_update(source, target, tau):
    '''Polyak averaging of the target networks (sac_utils.py:369-371) as ONE fused multi-tensor op.'''
    with torch.no_grad():
        tp, sp = list(target.parameters()), list(source.parameters())
        torch._foreach_mul_(tp, 1.0 - tau)
        torch._foreach_add_(tp, sp, alpha=tau)


class SquashedGaussianActor(nn.Module):
    '''sac_utils.py:149-181: trunk -> (mu, log_std clamped to [-20, 2]) -> tanh-squashed reparametrised sample -> affine map to the action box.'''

    def __init__(self, obs_dim, act_dim, hidden_dims, activation, low, high):
        super().__init__()
        self.net = MLP(obs_dim, hidden_dims[-1], hidden_dims[:-1], activation)
        self.mu_layer = nn.Linear(hidden_dims[-1], act_dim)
        self.log_std_layer = nn.Linear(hidden_dims[-1], act_dim)
        self.register_buffer('act_low', torch.as_tensor(low, dtype=torch.float32), persistent=False)
        self.register_buffer('act_high', torch.as_tensor(high, dtype=torch.float32), persistent=False)

    def forward(self, obs, deterministic=False, with_logprob=True):
        feat = self.net(obs)                      # MLP applies no activation after its last layer (neural_networks.py:48-54)
        mu, log_std = self.mu_layer(feat), torch.clamp(self.log_std_layer(feat), -20, 2)
        u = mu if deterministic else mu + log_std.exp() * torch.randn_like(mu)
        logp = None
        if with_logprob:
            std = log_std.exp()
            logp = (-((u - mu) ** 2) / (2 * std ** 2) - log_std - 0.5 * math.log(2 * math.pi)).sum(-1, keepdim=True)
            logp = logp - (2 * (math.log(2) - u - F.softplus(-2 * u))).sum(-1, keepdim=True)
        a = torch.tanh(u)
        return self.act_low + 0.5 * (a + 1.0) * (self.act_high - self.act_low), logp


class QFunction(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden_dims, activation):
        super().__init__()
        self.q_net = MLP(obs_dim + act_dim, 1, hidden_dims, activation)

    def forward(self, obs, act):
        return self.q_net(torch.cat([obs, act], dim=-1))


class SACActorCritic(nn.Module):
    def __init__(self, obs_space, act_space, hidden_dims=(256, 256), activation='relu'):
        super().__init__()
        od, ad = obs_space.shape[0], act_space.shape[0]
        self.actor = SquashedGaussianActor(od, ad, list(hidden_dims), activation, act_space.low, act_space.high)
        self.q1 = QFunction(od, ad, list(hidden_dims), activation)
        self.q2 = QFunction(od, ad, list(hidden_dims), activation)

    @torch.no_grad()
    def act(self, obs, deterministic=False):
        return self.actor(obs, deterministic, False)[0]              # stays on the device (the reference returns numpy)


class SACAgent:
    '''sac_utils.py:19-143.'''

    def __init__(self, obs_space, act_space, hidden_dim=256, gamma=0.99, tau=0.005, init_temperature=0.2, use_entropy_tuning=False, target_entropy=None,
                 actor_lr=0.001, critic_lr=0.001, entropy_lr=0.001, activation='relu', **kwargs):
        self.gamma, self.tau, self.use_entropy_tuning = gamma, tau, use_entropy_tuning
        self.ac = SACActorCritic(obs_space, act_space, [hidden_dim] * 2, activation)
        self.log_alpha = torch.tensor(np.log(init_temperature), dtype=torch.float32)
        self.target_entropy = (-float(np.prod(act_space.shape)) if target_entropy is None else target_entropy) if use_entropy_tuning else None
        self.ac_targ = copy.deepcopy(self.ac)
        for p in self.ac_targ.parameters():
            p.requires_grad = False
        self.actor_opt = torch.optim.Adam(self.ac.actor.parameters(), actor_lr)
        self.critic_opt = torch.optim.Adam(list(self.ac.q1.parameters()) + list(self.ac.q2.parameters()), critic_lr)
        self._entropy_lr = entropy_lr
        self.alpha_opt = torch.optim.Adam([self.log_alpha], entropy_lr)

    @property
    def alpha(self):
        return self.log_alpha.exp()

    def to(self, device):
        self.ac.to(device); self.ac_targ.to(device)
        self.log_alpha = self.log_alpha.to(device).requires_grad_(self.use_entropy_tuning)
        self.alpha_opt = torch.optim.Adam([self.log_alpha], self._entropy_lr)

    def train(self):
        self.ac.train()

    def eval(self):
        self.ac.eval()

    def state_dict(self):
        return {'ac': self.ac.state_dict(), 'log_alpha': self.log_alpha.detach().clone(), 'ac_targ': self.ac_targ.state_dict(),
                'actor_opt': self.actor_opt.state_dict(), 'critic_opt': self.critic_opt.state_dict(), 'alpha_opt': self.alpha_opt.state_dict()}

    def load_state_dict(self, sd):
        self.ac.load_state_dict(sd['ac']); self.ac_targ.load_state_dict(sd['ac_targ'])
        with torch.no_grad():
            self.log_alpha.copy_(torch.as_tensor(sd['log_alpha']).to(self.log_alpha.device))
        self.actor_opt.load_state_dict(sd['actor_opt']); self.critic_opt.load_state_dict(sd['critic_opt'])
        if sd.get('alpha_opt', {}).get('state'):
            self.alpha_opt.load_state_dict(sd['alpha_opt'])

    def update(self, batch):
        '''One gradient step of actor, (temperature,) critics, then the Polyak update; loss values stay device tensors (no .item() sync).'''
        obs, act, rew, next_obs, mask = (batch[k] for k in ('obs', 'act', 'rew', 'next_obs', 'mask'))
        pi, logp = self.ac.actor(obs)
        q = torch.min(self.ac.q1(obs, pi), self.ac.q2(obs, pi))
        policy_loss = (self.alpha.detach() * logp - q).mean()                                       # sac_utils.py:86-96
        self.actor_opt.zero_grad(set_to_none=True)
        policy_loss.backward()
        self.actor_opt.step()
        entropy_loss = torch.zeros((), device=obs.device)
        if self.use_entropy_tuning:
            entropy_loss = -(self.log_alpha * (logp + self.target_entropy).detach()).mean()
            self.alpha_opt.zero_grad(set_to_none=True)
            entropy_loss.backward()
            self.alpha_opt.step()
        with torch.no_grad():                                                                       # sac_utils.py:103-113
            na, nlogp = self.ac.actor(next_obs)
            q_targ = rew + self.gamma * mask * (torch.min(self.ac_targ.q1(next_obs, na), self.ac_targ.q2(next_obs, na)) - self.alpha * nlogp)
        critic_loss = (self.ac.q1(obs, act) - q_targ).pow(2).mean() + (self.ac.q2(obs, act) - q_targ).pow(2).mean()
        self.critic_opt.zero_grad(set_to_none=True)
        critic_loss.backward()
        self.critic_opt.step()
        soft_update(self.ac, self.ac_targ, self.tau)
        return {'policy_loss': policy_loss.detach(), 'critic_loss': critic_loss.detach(), 'entropy_loss': entropy_loss.detach()}
