'''DDPG agent pieces (safe_control_gym/controllers/ddpg/ddpg_utils.py) on the GPU: deterministic tanh actor mapped onto the action box, one
Q function, Polyak targets, and the exploration noise of ddpg.yaml (Ornstein-Uhlenbeck process with a scheduled std) kept as a device tensor
for the whole batch (the reference draws one numpy sample per env per step, ddpg.py:278-280).'''

import copy

import numpy as np
import torch
import torch.nn as nn

from safe_control_gym_b200.controllers.sac.sac_utils import QFunction, soft_update
from safe_control_gym_b200.math_and_models.neural_networks import MLP


class DeterministicActor(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden_dims, activation, low, high):
        super().__init__()
        self.net = MLP(obs_dim, act_dim, hidden_dims, activation)
        self.register_buffer('act_low', torch.as_tensor(low, dtype=torch.float32), persistent=False)
        self.register_buffer('act_high', torch.as_tensor(high, dtype=torch.float32), persistent=False)

    def forward(self, obs):
        return self.act_low + 0.5 * (torch.tanh(self.net(obs)) + 1.0) * (self.act_high - self.act_low)      # ddpg_utils.py:121-124,163-164


class DDPGActorCritic(nn.Module):
    def __init__(self, obs_space, act_space, hidden_dims=(256, 256), activation='relu'):
        super().__init__()
        od, ad = obs_space.shape[0], act_space.shape[0]
        self.actor = DeterministicActor(od, ad, list(hidden_dims), activation, act_space.low, act_space.high)
        self.q = QFunction(od, ad, list(hidden_dims), activation)

    @torch.no_grad()
    def act(self, obs, **kwargs):
        return self.actor(obs)


class DDPGAgent:
    '''ddpg_utils.py:17-108.'''

    def __init__(self, obs_space, act_space, hidden_dim=256, gamma=0.99, tau=0.005, actor_lr=0.001, critic_lr=0.001, activation='relu', **kwargs):
        self.gamma, self.tau = gamma, tau
        self.ac = DDPGActorCritic(obs_space, act_space, [hidden_dim] * 2, activation)
        self.ac_targ = copy.deepcopy(self.ac)
        for p in self.ac_targ.parameters():
            p.requires_grad = False
        self.actor_opt = torch.optim.Adam(self.ac.actor.parameters(), actor_lr)
        self.critic_opt = torch.optim.Adam(self.ac.q.parameters(), critic_lr)

    def to(self, device):
        self.ac.to(device); self.ac_targ.to(device)

    def train(self):
        self.ac.train()

    def eval(self):
        self.ac.eval()

    def state_dict(self):
        return {'ac': self.ac.state_dict(), 'ac_targ': self.ac_targ.state_dict(), 'actor_opt': self.actor_opt.state_dict(), 'critic_opt': self.critic_opt.state_dict()}

    def load_state_dict(self, sd):
        self.ac.load_state_dict(sd['ac']); self.ac_targ.load_state_dict(sd['ac_targ'])
        self.actor_opt.load_state_dict(sd['actor_opt']); self.critic_opt.load_state_dict(sd['critic_opt'])

    def update(self, batch):
        obs, act, rew, next_obs, mask = (batch[k] for k in ('obs', 'act', 'rew', 'next_obs', 'mask'))
        policy_loss = -self.ac.q(obs, self.ac.actor(obs)).mean()                                   # ddpg_utils.py:65-71
        self.actor_opt.zero_grad(set_to_none=True)
        policy_loss.backward()
        self.actor_opt.step()
        with torch.no_grad():                                                                      # ddpg_utils.py:77-81: online actor, target critic
            q_targ = rew + self.gamma * mask * self.ac_targ.q(next_obs, self.ac.actor(next_obs))
        critic_loss = (self.ac.q(obs, act) - q_targ).pow(2).mean()
        self.critic_opt.zero_grad(set_to_none=True)
        critic_loss.backward()
        self.critic_opt.step()
        soft_update(self.ac, self.ac_targ, self.tau)
        return {'policy_loss': policy_loss.detach(), 'critic_loss': critic_loss.detach()}


class OUNoise:
    '''Batched Ornstein-Uhlenbeck process (math_and_models/random_processes.py:30-44: x += theta (0 - x) dt + std sqrt(dt) N(0,1); std schedule.py:22-38),
    one state row per env, std from a constant or linearly decaying schedule.'''

    def __init__(self, batch, act_dim, std=0.2, theta=0.15, dt=1e-2, std_end=None, std_steps=None, device='cuda', generator=None):
        self.x = torch.zeros(batch, act_dim, device=device)
        self.std0, self.std_end, self.std_steps, self.theta, self.dt, self.gen, self.t = std, std_end, std_steps, theta, dt, generator, 0

    def std(self):
        if self.std_end is None or not self.std_steps:
            return self.std0
        f = min(1.0, self.t / self.std_steps)
        return self.std0 + (self.std_end - self.std0) * f

    def reset(self):
        self.x.zero_()

    def sample(self):
        n = torch.randn(self.x.shape, device=self.x.device, generator=self.gen)
        self.x = self.x - self.theta * self.x * self.dt + self.std() * np.sqrt(self.dt) * n
        self.t += 1
        return self.x


def make_action_noise_process(noise_config, batch, act_space, device, generator):
    '''ddpg.yaml `random_process: {func: OrnsteinUhlenbeckProcess | GaussianProcess, std: {func: LinearSchedule, args: v[, ...]}}`.'''
    if not noise_config:
        return None
    cfg = copy.deepcopy(dict(noise_config))
    func = cfg.pop('func', 'OrnsteinUhlenbeckProcess')
    std_cfg = dict(cfg.pop('std', {'args': 0.2}))
    std_cfg.pop('func', None)
    std0 = std_cfg.pop('args', 0.2)
    proc = OUNoise(batch, act_space.shape[0], std=float(std0), std_end=std_cfg.get('end'), std_steps=std_cfg.get('steps'), device=device, generator=generator)
    if func == 'GaussianProcess':
        proc.theta, proc.dt = 1.0, 1.0          # x <- std * N(0, 1): memoryless
    return proc
