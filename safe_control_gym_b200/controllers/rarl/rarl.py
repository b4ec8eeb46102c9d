'''Robust adversarial RL (RARL) on the fused collector (safe_control_gym/controllers/rarl/rarl.py).

Protagonist and adversary are two PPO agents that see the same observation; the adversary's sampled action goes through
`env.set_adversary_control` (clip to [-1, 1], scale, offset) into the SAME control step the protagonist acts in.  The reference's collector
loop (rarl.py:349-418: two `ac.step` calls, `env_method('set_adversary_control', ...)`, `env.step`, buffer push of ONE of the two
players, terminal / bootstrap values from that player's critic, returns and advantages) is ONE `scg_rollout` launch of the extended
collector (`scg_rollout_args.d_adv_policy`, `adv_record`): both actor-critics are evaluated inside the kernel, the record that comes back
is the protagonist's (`collect_rollouts()`) or the adversary's (`collect_rollouts(adversary=True)`: its action, value, log-probability,
`-rew`).  The alternating update schedule (rarl.py:417-451) and the checkpoint keys are the reference's; both agents use the fused PPO
update kernel.
'''

import os
import time
from collections import defaultdict

import numpy as np
import torch

from safe_control_gym_b200.controllers.ppo.ppo import DEFAULT_CONFIG as PPO_DEFAULTS
from safe_control_gym_b200.controllers.ppo.ppo import PPO, PeerGatheredRollout, all_gather_rollout, rollout_flat_numel
from safe_control_gym_b200.controllers.ppo.ppo_utils import PPOAgent
from safe_control_gym_b200.engine import pack_policy
from safe_control_gym_b200.envs.env_wrappers.vectorized_env import make_vec_envs

# defaults of safe_control_gym/controllers/rarl/rarl.yaml
DEFAULT_CONFIG = dict(PPO_DEFAULTS, agent_iterations=10, adversary_iterations=10, pretrained=None, train_protagonist=True, train_adversary=True)


class RARL(PPO):
    def __init__(self, env_func, training=True, checkpoint_path='model_latest.pt', output_dir='temp', use_gpu=True, seed=0, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(kwargs)
        super().__init__(env_func, training=training, checkpoint_path=checkpoint_path, output_dir=output_dir, use_gpu=use_gpu, seed=seed, **cfg)
        proto = self.env.unwrapped.proto
        if proto.adversary_disturbance is None:
            self.close()
            raise RuntimeError('[ERROR] adversary_disturbance does not exist, env.set_adversary_control() cannot be called.')
        # adversary specs from the env (rarl.py:70-76)
        self.adv_obs_space = self.env.get_attr('adversary_observation_space')[0]
        self.adv_act_space = self.env.get_attr('adversary_action_space')[0]
        self.adversary = PPOAgent(self.adv_obs_space, self.adv_act_space, hidden_dim=self.hidden_dim, use_clipped_value=self.use_clipped_value,
                                  clip_param=self.clip_param, target_kl=self.target_kl, entropy_coef=self.entropy_coef, actor_lr=self.actor_lr,
                                  critic_lr=self.critic_lr, opt_epochs=self.opt_epochs, mini_batch_size=self.mini_batch_size, activation=self.activation,
                                  fused_update=cfg.get('fused_update'))
        self.adversary.to(self.device)
        self._sync_replica(self.adversary)
        self._adv_rollout_out, self._adv_peers = {}, {}

    # ------------------------------------------------------------------------------------------
    def _adversary_blob(self):
        if self.adversary.fused is not None:
            return self.adversary.fused.rollout_blob()
        return pack_policy(self.adversary.ac, self.device)

    def collect_rollouts(self, adversary=False, T=None):
        '''rarl.py:349-418 as one launch: the current protagonist and adversary collect `rollout_steps` x `rollout_batch_size` transitions;
        the record belongs to the protagonist, or to the adversary (`adversary=True`).'''
        T = self.rollout_steps if T is None else T
        blob, dims = self._policy_blob()
        adv_blob, adv_dims = self._adversary_blob()
        eng = self.env.unwrapped.engine
        cache, peers = (self._adv_rollout_out, self._adv_peers) if adversary else (self._rollout_out, self._peers)
        kw = {}
        cached = cache.get(T)
        if self.world > 1:
            if T not in peers:
                rec_act = adv_dims[1] if adversary else dims[1]
                peers[T] = PeerGatheredRollout(rollout_flat_numel(dims[0], rec_act, T, eng.N), self.device)
            self._peer = peers[T]
            if cached is None:
                kw.update(self._peer.rollout_kwargs())
            self._peer.begin()
        want = ('ret', 'adv', 'cviol', 'mse', 'last', 'episode')
        if self.norm_obs or self.norm_reward:
            want += ('raw',)
            kw.update(norm=self._norm_blob(), obs0=self.obs.contiguous() if self.norm_obs else None)
        out = eng.rollout(blob, dims, T, activation=self.activation, gamma=self.gamma, gae_lambda=self.gae_lambda, use_gae=self.use_gae,
                          out=cached, want=want, adversary=(adv_blob, adv_dims), adv_record=adversary, **kw)
        cache[T] = out
        self.obs = out['last_obs']
        self.total_steps += self.rollout_batch_size * T
        if self.norm_obs or self.norm_reward:
            self._update_normalizers(out)
        if self.world > 1:
            batch = self._peer.finish(out) if self._peer is not None else all_gather_rollout(out)
        else:
            batch = {k: v for k, v in out.items() if not k.startswith('_')}
        rollouts = dict(batch)
        adv = batch['adv']
        rollouts['adv'] = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-6)          # rarl.py:417
        return rollouts, out

    def _update_player(self, agent, iterations, adversary):
        results = defaultdict(list)
        for _ in range(iterations):
            rollouts, out = self.collect_rollouts(adversary=adversary)
            res = agent.update(rollouts, self.device, generator=self._perm_gen, defer_stats=True)
            self.env.record_rollout(out)
            res = agent.finish_update(res)
            for k, v in res.items():
                results[k].append(v)
        return {k: sum(v) / len(v) for k, v in results.items()}

    def update_agent(self):
        '''rarl.py:419-434'''
        self.agent.train(); self.adversary.eval()
        return self._update_player(self.agent, self.agent_iterations, adversary=False)

    def update_adversary(self):
        '''rarl.py:436-451'''
        self.agent.eval(); self.adversary.train()
        return {k + '_adv': v for k, v in self._update_player(self.adversary, self.adversary_iterations, adversary=True).items()}

    def train_step(self):
        '''rarl.py:267-283: the two players are updated by turns.'''
        self.obs_normalizer.unset_read_only()
        start = time.time()
        results = {}
        if self.train_protagonist:
            results.update(self.update_agent())
        if self.train_adversary:
            results.update(self.update_adversary())
        results.update({'step': self.total_steps, 'elapsed_time': time.time() - start})
        return results

    # ------------------------------------------------------------------------------------------
    def save(self, path):
        super().save(path)
        if self.rank == 0:                                            # rarl.py:127-132: + 'adversary'
            state = torch.load(path, weights_only=False)
            state['adversary'] = self.adversary.state_dict()
            torch.save(state, path)

    def load(self, path):
        super().load(path)
        state = torch.load(path, weights_only=False, map_location=self.device)
        self.adversary.load_state_dict(state['adversary'])

    def run(self, env=None, render=False, n_episodes=10, verbose=False, use_adv=False, **kwargs):
        '''Deterministic evaluation (rarl.py:210-265): the protagonist's mean action; `use_adv` lets the adversary push with ITS mean action,
        otherwise the adversary input is zero (clip(0) * scale + offset, as the reference passes np.zeros).'''
        self.agent.eval(); self.adversary.eval()
        if self.eval_env is None or self.eval_env.num_envs != n_episodes:
            if self.eval_env is not None:
                self.eval_env.close()
            self.eval_env = make_vec_envs(self.env_func, None, n_episodes, 1, self.seed * 111, rng_mode=self.env.unwrapped.rng_mode,
                                          device=self.device.index, precision=self.precision)
        obs, _ = self.eval_env.reset()
        blob, dims = self._policy_blob()
        if use_adv:
            adv = self._adversary_blob()
        else:                                                         # a zero adversary: all-zero weights and biases -> mean action 0
            adv_blob, adv_dims = self._adversary_blob()
            adv = (torch.zeros_like(adv_blob), adv_dims)
        T = int(self.eval_env.proto.CTRL_STEPS)
        kw = {}
        if self.norm_obs:
            self.obs_normalizer.set_read_only()
            kw.update(norm=self._norm_blob(), obs0=self.obs_normalizer.normalize(obs).to(torch.float32).contiguous())
            kw['norm'][2 * dims[0] + 4] = 0.0
        out = self.eval_env.engine.rollout(blob, dims, T, activation=self.activation, deterministic=True, want=('cviol', 'mse', 'episode'),
                                           adversary=adv, **kw)
        mask = out['mask'][..., 0]
        first_done = torch.where((mask == 0).any(0), (mask == 0).float().argmax(0), torch.full((n_episodes,), T - 1, device=self.device))
        live = (torch.arange(T, device=self.device)[:, None] <= first_done[None, :]).float()
        ep_lengths = first_done + 1
        return {'ep_returns': (out['rew'][..., 0] * live).sum(0).cpu().numpy(), 'ep_lengths': ep_lengths.cpu().numpy(),
                'constraint_violation': (out['cviol'][..., 0] * live).sum(0).cpu().numpy(),
                'mse': ((out['mse'][..., 0] * live).sum(0) / ep_lengths).cpu().numpy()}

    def log_step(self, results):
        step = results['step']
        self.logger.add_scalars({'step': step, 'time': results['elapsed_time'], 'progress': step / self.max_env_steps}, step, prefix='time')
        self.logger.add_scalars({k: results[k] for k in ['policy_loss', 'value_loss', 'entropy_loss'] if k in results}, step, prefix='loss')
        self.logger.add_scalars({k: results[k + '_adv'] for k in ['policy_loss', 'value_loss', 'entropy_loss'] if k + '_adv' in results}, step, prefix='loss_adv')
        if len(self.env.return_queue) > 0:
            r, l = np.asarray(self.env.return_queue), np.asarray(self.env.length_queue)
            self.logger.add_scalars({'ep_length': l.mean(), 'ep_return': r.mean(), 'ep_reward': (r / l).mean()}, step, prefix='stat')
        self.logger.dump_scalars()
