'''SafeExplorer PPO on the fused collector (safe_control_gym/controllers/safe_explorer/safe_ppo.py).

Two stages like the reference: (1) `pretraining: True` -- fit the safety layer's constraint models on transitions collected with random
actions (safe_ppo.py:218-289: `learn` -> `pretrain_step`); here the transitions come from the batched step API (`CudaVecEnv.step`, one
launch per step for the whole batch).  (2) `pretraining: False, pretrained: <path>` -- PPO whose actor mean is projected by the safety
layer before sampling (safe_ppo_utils.py:88-104).  The collector loop of stage 2 (safe_ppo.py:299-345) is ONE `scg_rollout` launch with the
safety layer evaluated in the kernel (`scg_rollout_args.safety_*`) and the constraint values `c` in the record; the update is the torch
restatement of `SafePPOAgent.update` (the projection sits inside the actor, so its gradient flows through it -- the fused update kernel
does not model that).
'''

import os
import time
from collections import defaultdict

import numpy as np
import torch

from safe_control_gym_b200.controllers.ppo.ppo import DEFAULT_CONFIG as PPO_DEFAULTS
from safe_control_gym_b200.controllers.ppo.ppo import PPO
from safe_control_gym_b200.controllers.safe_explorer.safe_explorer_utils import ConstraintBuffer, SafetyLayer
from safe_control_gym_b200.math_and_models.distributions import Normal

# defaults of safe_control_gym/controllers/safe_explorer/safe_ppo.yaml
DEFAULT_CONFIG = dict(PPO_DEFAULTS, pretraining=True, pretrained=None, constraint_hidden_dim=10, constraint_lr=0.0001, constraint_batch_size=256,
                      constraint_steps_per_epoch=6000, constraint_epochs=25, constraint_eval_steps=1500, constraint_eval_interval=5,
                      constraint_buffer_size=1000000, constraint_slack=None)


class SafeExplorerPPO(PPO):
    def __init__(self, env_func, training=True, checkpoint_path='model_latest.pt', output_dir='temp', use_gpu=True, seed=0, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(kwargs)
        cfg['fused_update'] = False                       # the projection is part of the actor: the update differentiates through it (torch)
        super().__init__(env_func, training=training, checkpoint_path=checkpoint_path, output_dir=output_dir, use_gpu=use_gpu, seed=seed, **cfg)
        proto = self.env.unwrapped.proto
        if proto.constraints is None:
            raise ValueError('SafeExplorerPPO needs an env with state constraints')
        self.num_constraints = proto.constraints.num_state_constraints           # safe_ppo.py:43
        slack = self.constraint_slack if self.constraint_slack is not None else 0.0
        self.safety_layer = SafetyLayer(self.env.observation_space, self.env.action_space, hidden_dim=self.constraint_hidden_dim,
                                        num_constraints=self.num_constraints, lr=self.constraint_lr, slack=slack)
        self.safety_layer.to(self.device)
        self._sample_gen = torch.Generator(device=self.device)
        self._sample_gen.manual_seed(int(seed) + 777)

    # ------------------------------------------------------------------------------------------
    def _state_constraint_values(self, batch):
        return batch['cvals'][:, :self.num_constraints].to(torch.float32)

    def reset(self):
        if self.training and self.pretraining:
            self.constraint_buffer = ConstraintBuffer(self.env.observation_space.shape[0], self.env.action_space.shape[0], self.num_constraints,
                                                      self.constraint_buffer_size, self.device)
            self.total_steps = 0
            obs, _ = self.env.reset()
            self.obs = obs
            self.c = None
            return
        if self.training:
            assert self.pretrained, 'Must provide a pre-trained model for adaptation.'
            path = os.path.join(self.pretrained, 'model_latest.pt') if os.path.isdir(self.pretrained) else self.pretrained
            state = torch.load(path, weights_only=False, map_location=self.device)
            self.safety_layer.load_state_dict(state['safety_layer'])
        super().reset()

    def save(self, path):
        super().save(path)
        if self.rank == 0:
            state = torch.load(path, weights_only=False)
            state['safety_layer'] = self.safety_layer.state_dict()
            torch.save(state, path)

    def load(self, path):
        super().load(path)
        state = torch.load(path, weights_only=False, map_location=self.device)
        if 'safety_layer' in state:
            self.safety_layer.load_state_dict(state['safety_layer'])

    # ------------------------------------------------------------------------------------------
    def learn(self, env=None, **kwargs):
        if not self.pretraining:
            return super().learn(env=env, **kwargs)
        final = self.constraint_epochs                                           # safe_ppo.py:218-244
        results = {}
        for epoch in range(final):
            results = self.pretrain_step()
            if self.rank == 0 and (epoch == final - 1 or (self.save_interval and epoch % self.save_interval == 0)):
                self.save(self.checkpoint_path)
            if self.log_interval and epoch % self.log_interval == 0:
                self.logger.info(f'pretrain epoch {epoch}: ' + ', '.join(f'{k} {v:.4g}' for k, v in results.items() if k.startswith('constraint_')))
        return results

    def _random_actions(self, n):
        sp = self.env.action_space
        lo = torch.as_tensor(sp.low, dtype=torch.float32, device=self.device)
        hi = torch.as_tensor(sp.high, dtype=torch.float32, device=self.device)
        return lo + (hi - lo) * torch.rand(n, lo.numel(), device=self.device, generator=self._sample_gen)

    def pretrain_step(self):
        '''One epoch of safety-layer pre-training (safe_ppo.py:246-289): collect transitions with uniformly random actions, then regress.'''
        start = time.time()
        N = self.env.num_envs
        steps = 0
        # `cvals` of the step API are the values of the state the step ENDED in (the terminal state when done, safe_ppo.py:442-447); the
        # values of a freshly reset state only arrive with the next step, so a transition that STARTS at a reset state is not pushed
        # (the reference reads them from the reset info; that is 1 transition per episode)
        if self.c is None:
            self.c = torch.zeros(N, self.num_constraints, device=self.device)
            self.fresh = torch.ones(N, dtype=torch.bool, device=self.device)
        while steps < self.constraint_steps_per_epoch:
            act = self._random_actions(N)
            obs_next, _, done, info = self.env.step(act)
            c_next = self._state_constraint_values(info['batch'])
            keep = ~self.fresh
            if bool(keep.any()):
                self.constraint_buffer.push({'obs': self.obs[keep], 'act': act[keep], 'c': self.c[keep], 'c_next': c_next[keep]})
            self.obs, self.c, self.fresh = obs_next, c_next, done.bool().reshape(-1)
            steps += N
        self.total_steps += steps
        results = defaultdict(list)
        self.safety_layer.train()
        for batch in self.constraint_buffer.sampler(self.constraint_batch_size, generator=self._sample_gen):
            for k, v in self.safety_layer.update(batch).items():
                results[k].append(v)
        self.constraint_buffer.reset()
        out = {k: sum(v) / len(v) for k, v in results.items()}
        out.update({'step': self.total_steps, 'elapsed_time': time.time() - start})
        return out

    # ------------------------------------------------------------------------------------------
    def select_action(self, obs, info=None):
        with torch.inference_mode():
            obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device).reshape(1, -1)
            c = torch.as_tensor(np.asarray(info['constraint_values'])[:self.num_constraints], dtype=torch.float32, device=self.device).reshape(1, -1)
            mu = self.agent.ac.actor.pi_net(obs)
            return self.safety_layer.get_safe_action(obs, mu, c)[0].cpu().numpy()

    def _collector_kwargs(self):
        if self.world > 1:
            raise NotImplementedError('SafeExplorerPPO: single-GPU collector')
        self.safety_layer.eval()
        return dict(safety=self.safety_layer.pack(self.device), mlp_engine='ffma2')

    def train_step(self):
        '''safe_ppo.py:299-361 with the collector loop fused; `SafePPOAgent.update` (safe_ppo_utils.py:24-64) in torch.'''
        self.agent.train()
        start = time.time()
        out = self.collect()
        self.obs = out['last_obs']
        self.total_steps += self.rollout_batch_size * self.rollout_steps
        self.env.record_rollout(out)
        adv = out['adv']
        roll = {k: out[k].reshape(-1, out[k].shape[-1]) for k in ('obs', 'act', 'logp', 'ret', 'v', 'c')}
        roll['adv'] = ((adv - adv.mean()) / (adv.std(unbiased=False) + 1e-6)).reshape(-1, 1)
        results = self._update(roll)
        results.update({'step': self.total_steps, 'elapsed_time': time.time() - start})
        return results

    def _update(self, roll):
        ag = self.agent
        total = roll['obs'].shape[0]
        n_mb = total // ag.mini_batch_size
        assert n_mb != 0, 'num_mini_batch is 0'
        sums = torch.zeros(4, device=self.device)
        for _ in range(ag.opt_epochs):
            perm = torch.randperm(total, device=self.device, generator=self._perm_gen)
            for i in range(n_mb):
                idx = perm[i * ag.mini_batch_size:(i + 1) * ag.mini_batch_size]
                b = {k: v[idx] for k, v in roll.items()}
                mu = ag.ac.actor.pi_net(b['obs'])
                mu_safe = self.safety_layer.get_safe_action(b['obs'], mu, b['c'])          # safe_ppo_utils.py:88-104
                dist = Normal(mu_safe, ag.ac.actor.logstd.exp())
                logp = dist.log_prob(b['act'])
                ratio = torch.exp(logp - b['logp'])
                clip_adv = torch.clamp(ratio, 1 - ag.clip_param, 1 + ag.clip_param) * b['adv']
                policy_loss = -torch.min(ratio * b['adv'], clip_adv).mean()
                entropy_loss = -dist.entropy().mean()
                approx_kl = (b['logp'] - logp).mean()
                if ag.target_kl <= 0 or float(approx_kl.detach()) <= 1.5 * ag.target_kl:
                    ag.actor_opt.zero_grad(set_to_none=True)
                    (policy_loss + ag.entropy_coef * entropy_loss).backward()
                    ag.actor_opt.step()
                value_loss = ag.compute_value_loss(b)
                ag.critic_opt.zero_grad(set_to_none=True)
                value_loss.backward()
                ag.critic_opt.step()
                sums += torch.stack([policy_loss.detach(), value_loss.detach(), entropy_loss.detach(), approx_kl.detach()])
        p, v, e, k = (sums / (ag.opt_epochs * n_mb)).tolist()
        return {'policy_loss': p, 'value_loss': v, 'entropy_loss': e, 'approx_kl': k}
