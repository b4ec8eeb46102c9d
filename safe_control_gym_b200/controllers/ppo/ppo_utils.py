'''PPO agent pieces (safe_control_gym/controllers/ppo/ppo_utils.py) living entirely on the GPU.

`MLPActorCritic` keeps the reference's module tree (`actor.pi_net.fcs.*`, `actor.logstd`, `critic.v_net.fcs.*`) so
checkpoints are interchangeable.  The rollout buffer is a dict of torch.cuda tensors in the PPOBuffer scheme
(ppo_utils.py:266-295) filled by ONE fused kernel launch; `PPOAgent.update` consumes it in place (no host copy).
'''

from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn

from safe_control_gym_b200.math_and_models.distributions import Normal
from safe_control_gym_b200.math_and_models.neural_networks import MLP


class MLPActor(nn.Module):
    def __init__(self, obs_dim, act_dim, hidden_dims, activation):
        super().__init__()
        self.pi_net = MLP(obs_dim, act_dim, hidden_dims, activation)
        self.logstd = nn.Parameter(-0.5 * torch.ones(act_dim))      # ppo_utils.py:166

    def forward(self, obs, act=None):
        dist = Normal(self.pi_net(obs), self.logstd.exp())
        return dist, (dist.log_prob(act) if act is not None else None)


class MLPCritic(nn.Module):
    def __init__(self, obs_dim, hidden_dims, activation):
        super().__init__()
        self.v_net = MLP(obs_dim, 1, hidden_dims, activation)

    def forward(self, obs):
        return self.v_net(obs)


class MLPActorCritic(nn.Module):
    def __init__(self, obs_space, act_space, hidden_dims=(64, 64), activation='tanh'):
        super().__init__()
        obs_dim, act_dim = obs_space.shape[0], act_space.shape[0]
        self.activation = activation
        self.actor = MLPActor(obs_dim, act_dim, list(hidden_dims), activation)
        self.critic = MLPCritic(obs_dim, list(hidden_dims), activation)

    @torch.no_grad()
    def step(self, obs):
        '''Stochastic step (ppo_utils.py:224-231); returns device tensors.'''
        dist, _ = self.actor(obs)
        action = dist.sample()
        return action, self.critic(obs), dist.log_prob(action)

    @torch.no_grad()
    def act(self, obs):
        dist, _ = self.actor(obs)
        return dist.mode()


class PPOAgent:
    '''Models, optimisers and the clipped-surrogate update (ppo_utils.py:16-146).'''

    def __init__(self, obs_space, act_space, hidden_dim=64, use_clipped_value=False, clip_param=0.2, target_kl=0.01,
                 entropy_coef=0.01, actor_lr=0.0003, critic_lr=0.001, opt_epochs=10, mini_batch_size=64, activation='tanh', **kwargs):
        self.obs_space, self.act_space = obs_space, act_space
        self.use_clipped_value, self.clip_param, self.target_kl = use_clipped_value, clip_param, target_kl
        self.entropy_coef, self.opt_epochs, self.mini_batch_size, self.activation = entropy_coef, opt_epochs, mini_batch_size, activation
        self.ac = MLPActorCritic(obs_space, act_space, hidden_dims=[hidden_dim] * 2, activation=activation)
        self.actor_opt = torch.optim.Adam(self.ac.actor.parameters(), actor_lr)
        self.critic_opt = torch.optim.Adam(self.ac.critic.parameters(), critic_lr)

    def to(self, device):
        self.ac.to(device)

    def train(self):
        self.ac.train()

    def eval(self):
        self.ac.eval()

    def state_dict(self):
        return {'ac': self.ac.state_dict(), 'actor_opt': self.actor_opt.state_dict(), 'critic_opt': self.critic_opt.state_dict()}

    def load_state_dict(self, sd):
        self.ac.load_state_dict(sd['ac'])
        self.actor_opt.load_state_dict(sd['actor_opt'])
        self.critic_opt.load_state_dict(sd['critic_opt'])

    def compute_policy_loss(self, batch):
        dist, logp = self.ac.actor(batch['obs'], batch['act'])
        ratio = torch.exp(logp - batch['logp'])
        clip_adv = torch.clamp(ratio, 1 - self.clip_param, 1 + self.clip_param) * batch['adv']
        policy_loss = -torch.min(ratio * batch['adv'], clip_adv).mean()
        entropy_loss = -dist.entropy().mean()
        approx_kl = (batch['logp'] - logp).mean()
        return policy_loss, entropy_loss, approx_kl

    def compute_value_loss(self, batch):
        v_cur = self.ac.critic(batch['obs'])
        if self.use_clipped_value:
            v_old_clipped = batch['v'] + (v_cur - batch['v']).clamp(-self.clip_param, self.clip_param)
            return 0.5 * torch.max((v_cur - batch['ret']).pow(2), (v_old_clipped - batch['ret']).pow(2)).mean()
        return 0.5 * (v_cur - batch['ret']).pow(2).mean()

    def update(self, rollouts, device=None, generator=None):
        '''rollouts: dict of device tensors, [T][N][w] (one rank) or [G][T*n][w] views of the all-gathered record.
        Minibatches are index-gathered on the GPU (two-level index, so the strided gathered views are never copied).'''
        flat = {k: v.reshape(v.shape[0], -1, v.shape[-1]) if v.dim() == 3 else v for k, v in rollouts.items()
                if k in ('obs', 'act', 'logp', 'adv', 'ret', 'v')}
        d0, d1 = flat['obs'].shape[0], flat['obs'].shape[1]
        total = d0 * d1
        num_mini_batch = total // self.mini_batch_size
        assert num_mini_batch != 0, 'num_mini_batch is 0'
        sums = torch.zeros(4, device=flat['obs'].device)
        for _ in range(self.opt_epochs):
            perm = torch.randperm(total, device=flat['obs'].device, generator=generator)
            for i in range(num_mini_batch):
                idx = perm[i * self.mini_batch_size:(i + 1) * self.mini_batch_size]
                batch = {k: v[idx // d1, idx % d1] for k, v in flat.items()}
                policy_loss, entropy_loss, approx_kl = self.compute_policy_loss(batch)
                # update only when there is no KL constraint or it is satisfied (ppo_utils.py:128)
                if self.target_kl <= 0 or float(approx_kl.detach()) <= 1.5 * self.target_kl:
                    self.actor_opt.zero_grad(set_to_none=True)
                    (policy_loss + self.entropy_coef * entropy_loss).backward()
                    self.actor_opt.step()
                value_loss = self.compute_value_loss(batch)
                self.critic_opt.zero_grad(set_to_none=True)
                value_loss.backward()
                self.critic_opt.step()
                sums += torch.stack([policy_loss.detach(), value_loss.detach(), entropy_loss.detach(), approx_kl.detach()])
        n = self.opt_epochs * num_mini_batch
        p, v, e, k = (sums / n).tolist()
        return {'policy_loss': p, 'value_loss': v, 'entropy_loss': e, 'approx_kl': k}


def compute_returns_and_advantages(rews, vals, masks, terminal_vals=0, last_val=0, gamma=0.99, use_gae=False, gae_lambda=0.95):
    '''Torch restatement of ppo_utils.py:374-400 for [T][N][1] device tensors (the fused kernel does this in-kernel;
    this version serves the step-wise collector).'''
    T = rews.shape[0]
    rets, advs = torch.zeros_like(rews), torch.zeros_like(rews)
    ret, adv = last_val, torch.zeros_like(rews[0], dtype=torch.float64)
    vals = torch.cat([vals, last_val.unsqueeze(0)], 0)
    rews += gamma * terminal_vals
    for i in reversed(range(T)):
        ret = rews[i] + gamma * masks[i] * ret
        if not use_gae:
            adv = ret - vals[i]
        else:
            td = rews[i] + gamma * masks[i] * vals[i + 1] - vals[i]
            adv = adv * gae_lambda * gamma * masks[i] + td
        rets[i], advs[i] = ret, adv
    return rets, advs
