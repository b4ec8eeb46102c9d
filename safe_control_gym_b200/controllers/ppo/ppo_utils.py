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
                 entropy_coef=0.01, actor_lr=0.0003, critic_lr=0.001, opt_epochs=10, mini_batch_size=64, activation='tanh',
                 fused_update=None, **kwargs):
        self.obs_space, self.act_space = obs_space, act_space
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        # fused_update: None = use the fused sm_100a update kernel whenever the network shape is supported, True = require it,
        # False = the step-by-step torch restatement of ppo_utils.py:82-146 (kept as the parity reference of the kernel)
        self.fused_update, self.fused = fused_update, None
        self.use_clipped_value, self.clip_param, self.target_kl = use_clipped_value, clip_param, target_kl
        self.entropy_coef, self.opt_epochs, self.mini_batch_size, self.activation = entropy_coef, opt_epochs, mini_batch_size, activation
        self.ac = MLPActorCritic(obs_space, act_space, hidden_dims=[hidden_dim] * 2, activation=activation)
        self.actor_opt = torch.optim.Adam(self.ac.actor.parameters(), actor_lr)
        self.critic_opt = torch.optim.Adam(self.ac.critic.parameters(), critic_lr)

    def to(self, device):
        self.ac.to(device)
        device = torch.device(device)
        obs_dim, act_dim = self.obs_space.shape[0], self.act_space.shape[0]
        hid = self.ac.actor.pi_net.fcs[0].out_features
        supported = device.type == 'cuda' and hid in (32, 64) and obs_dim <= 32 and act_dim <= 4
        if self.fused_update and not supported:
            raise NotImplementedError(f'fused PPO update: unsupported shape/device (hidden {hid}, obs {obs_dim}, act {act_dim}, {device})')
        if supported and self.fused_update is not False:
            from safe_control_gym_b200.engine import FusedPPOState
            self.fused = FusedPPOState(self.ac, device)

    def train(self):
        self.ac.train()

    def eval(self):
        self.ac.eval()

    def _opt_groups(self):
        names = {id(p): n for n, p in self.ac.named_parameters()}
        return [(self.actor_opt, 0), (self.critic_opt, 1)], names

    def state_dict(self):
        '''Checkpoint format of the reference (ppo_utils.py:47-54): torch.optim.Adam state dicts, also when the fused kernel owns the moments.'''
        sd = {'ac': self.ac.state_dict(), 'actor_opt': self.actor_opt.state_dict(), 'critic_opt': self.critic_opt.state_dict()}
        if self.fused is not None:
            steps = self.fused.steps.tolist()
            groups, names = self._opt_groups()
            for opt, which in groups:
                state = {}
                if steps[which] > 0:
                    for i, p in enumerate(opt.param_groups[0]['params']):
                        n = names[id(p)]
                        state[i] = {'step': torch.tensor(float(steps[which])), 'exp_avg': self.fused.views(self.fused.exp_avg, n).clone(),
                                    'exp_avg_sq': self.fused.views(self.fused.exp_avg_sq, n).clone()}
                sd['actor_opt' if which == 0 else 'critic_opt']['state'] = state
        return sd

    def load_state_dict(self, sd):
        self.ac.load_state_dict(sd['ac'])          # copy_ into the existing parameters: the flat training vector stays the storage
        self.actor_opt.load_state_dict(sd['actor_opt'])
        self.critic_opt.load_state_dict(sd['critic_opt'])
        if self.fused is not None:
            groups, names = self._opt_groups()
            self.fused.exp_avg.zero_()
            self.fused.exp_avg_sq.zero_()
            steps = [0, 0]
            for opt, which in groups:
                st = sd['actor_opt' if which == 0 else 'critic_opt']['state']
                for i, p in enumerate(opt.param_groups[0]['params']):
                    if i in st:
                        n = names[id(p)]
                        self.fused.views(self.fused.exp_avg, n).copy_(st[i]['exp_avg'])
                        self.fused.views(self.fused.exp_avg_sq, n).copy_(st[i]['exp_avg_sq'])
                        steps[which] = int(st[i]['step'])
            self.fused.steps.copy_(torch.tensor(steps, dtype=torch.int64))

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

    def update(self, rollouts, device=None, generator=None, defer_stats=False):
        '''rollouts: dict of device tensors, [T][N][w] (one rank) or [G][T*n][w] views of the all-gathered record.
        With the fused kernel (default on sm_100a) the whole opt_epochs x minibatch sequence is ONE launch (csrc/scg_ppo.cu);
        `update_torch` is the step-by-step restatement of the reference and the kernel's parity reference.'''
        if self.fused is None:
            return self.update_torch(rollouts, device, generator)           # (finish_update passes plain dicts through)
        flat = {k: v.reshape(v.shape[0], -1, v.shape[-1]) if v.dim() != 3 else v for k, v in rollouts.items()
                if k in ('obs', 'act', 'logp', 'adv', 'ret', 'v')}
        dev = flat['obs'].device
        total = flat['obs'].shape[0] * flat['obs'].shape[1]
        assert total // self.mini_batch_size != 0, 'num_mini_batch is 0'
        perm = torch.stack([torch.randperm(total, device=dev, generator=generator) for _ in range(self.opt_epochs)]).to(torch.int32)
        self.fused.stats.zero_()
        group = self.actor_opt.param_groups[0]
        n_mb = self.fused.update(flat, perm, self.mini_batch_size, activation=self.activation, clip_param=self.clip_param,
                                 target_kl=self.target_kl, entropy_coef=self.entropy_coef, use_clipped_value=self.use_clipped_value,
                                 actor_lr=group['lr'], critic_lr=self.critic_opt.param_groups[0]['lr'], betas=group['betas'], eps=group['eps'])
        pending = {'_pending_fused_stats': self.opt_epochs * n_mb}
        return pending if defer_stats else self.finish_update(pending)

    def finish_update(self, results):
        '''Reads the loss sums of a fused update back (one host sync).  Separate from `update` so that the caller can queue more device
        work first.'''
        if '_pending_fused_stats' not in results:
            return results
        n = results['_pending_fused_stats']
        st = self.fused.stats.tolist()
        return {'policy_loss': st[0] / n, 'value_loss': st[1] / n, 'entropy_loss': st[2] / n, 'approx_kl': st[3] / n,
                'actor_steps': int(st[4]), 'minibatches': int(st[5])}

    def update_torch(self, rollouts, device=None, generator=None):
        '''Minibatches are index-gathered on the GPU (two-level index, so the strided gathered views are never copied).'''
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
