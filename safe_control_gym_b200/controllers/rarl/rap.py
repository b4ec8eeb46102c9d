'''RARL with an adversary population (RAP) on the fused collector (safe_control_gym/controllers/rarl/rap.py).

Every collect draws one adversary of the population per env (rap.py:354-356: sorted, so that the envs of one adversary are contiguous), resets the
envs (rap.py:360), rolls the protagonist against those adversaries and keeps BOTH players' records of the same trajectories; the protagonist is updated
on the whole batch, every adversary on the slice of envs it pushed (rap.py:266-276, 443-470).

The extended collector (`scg_rollout_args.d_adv_policy`) evaluates ONE adversary for all envs of a launch and records ONE player.  RAP is composed from
that launch without new device code: the trajectories are a pure function of (env state, policy blobs) -- every random draw is a counter-based Philox
draw keyed by (env, episode, step) -- so the env state after the reset is snapshotted (`scg_get_state`) and, for every adversary g that was drawn, the
launch runs twice from the snapshot (`scg_set_state`): once recording the protagonist, once the adversary (`adv_record`).  The columns of g's envs are
copied out of those two records.  Cost: 2 x |adversaries drawn| launches over all N envs per collect instead of one -- a functional composition, not a
fused RAP collector (that would need a per-env adversary index and a second record in the kernel: DESIGN.md 8); at the reference's batch sizes
(rollout_batch_size 4, two adversaries) the launches are microseconds.

STATUS: `collect_rollouts` is verified against the REFERENCE's own `RAP.collect_rollouts()` on the CPU: tests/test_rap_reference.py runs it on an adapter over
the oracle (same get_state / set_state / rollout contract as the engine) with the weights, adversary draw and normals of a recorded reference run, and gets
the protagonist's buffer and every adversary's columns the reference returned; tests/test_rap_host.py covers the bookkeeping against a stand-in engine.  Not yet
verified on a B200: the CUDA-side plumbing of this class (constructor, blobs of the fused update state, the engine calls, save / load / run) -- its GPU test
(tests/test_gpu_extended.py, opt-in with `SCG_RUN_UNVERIFIED=1`) ran once with the last GPU seconds of round 2 and stopped at a host-side unpacking bug,
fixed since.
'''

import time

import numpy as np
import torch

from safe_control_gym_b200.controllers.ppo.ppo_utils import PPOAgent
from safe_control_gym_b200.controllers.rarl.rarl import DEFAULT_CONFIG as RARL_DEFAULTS
from safe_control_gym_b200.controllers.rarl.rarl import RARL

# defaults of safe_control_gym/controllers/rarl/rap.yaml
DEFAULT_CONFIG = dict(RARL_DEFAULTS, num_adversaries=2)

PLAYER_KEYS = ('act', 'rew', 'v', 'logp', 'terminal_v', 'ret', 'adv')          # what differs between the two players' records
SHARED_KEYS = ('obs', 'mask')                                                  # rap.py:436-437: rollouts_adv.obs / .mask are the protagonist's
EPISODE_KEYS = ('ep_return_sum', 'ep_length_sum', 'ep_cviol_sum', 'ep_count')


def draw_adversaries(rng, num_adversaries, batch):
    '''rap.py:354-356: one adversary per env, sorted; -> (index per env, adversaries present, first env of each, one past its last env).'''
    idx = np.sort(rng.randint(num_adversaries, size=batch))
    groups, starts = np.unique(idx, return_index=True)
    ends = np.concatenate([starts[1:], [batch]])
    return idx, groups, starts, ends


def normalise_advantages(adv):
    '''rap.py:433, 459 (numpy std: ddof = 0)'''
    return (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-6)


def split_by_adversary(rollouts_adv, groups, starts, ends):
    '''rap.py:461-470: [[adversary index, that adversary's columns of every key], ...]'''
    return [[int(g), {k: v[:, int(s):int(e)].contiguous() for k, v in rollouts_adv.items()}] for g, s, e in zip(groups, starts, ends)]


class RAP(RARL):
    def __init__(self, env_func, training=True, checkpoint_path='model_latest.pt', output_dir='temp', use_gpu=True, seed=0, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(kwargs)
        super().__init__(env_func, training=training, checkpoint_path=checkpoint_path, output_dir=output_dir, use_gpu=use_gpu, seed=seed, **cfg)
        if self.world > 1:
            self.close()
            raise NotImplementedError('rap: one rank (the composition replays launches from a host snapshot of the env state)')
        if self.norm_obs or self.norm_reward:
            self.close()
            raise NotImplementedError('rap: norm_obs / norm_reward (rap.yaml default: off) are not wired into the replayed launches')
        self.num_adversaries = int(cfg['num_adversaries'])
        # rap.py:75-77: a population of identical-shape PPO agents (the first one is the agent RARL built)
        self.adversaries = [self.adversary]
        for _ in range(self.num_adversaries - 1):
            adv = PPOAgent(self.adv_obs_space, self.adv_act_space, hidden_dim=self.hidden_dim, use_clipped_value=self.use_clipped_value,
                           clip_param=self.clip_param, target_kl=self.target_kl, entropy_coef=self.entropy_coef, actor_lr=self.actor_lr,
                           critic_lr=self.critic_lr, opt_epochs=self.opt_epochs, mini_batch_size=self.mini_batch_size, activation=self.activation,
                           fused_update=cfg.get('fused_update'))
            adv.to(self.device)
            self.adversaries.append(adv)
        self._adv_rng = np.random.RandomState(self.seed * 7 + 3)        # (the reference uses numpy's global generator here)
        self._rap_out = {}

    # ------------------------------------------------------------------------------------------
    def _blob_of(self, agent):
        '''-> (packed rollout blob, dims) of one agent, like PPO._policy_blob / RARL._adversary_blob'''
        from safe_control_gym_b200.engine import pack_policy
        return agent.fused.rollout_blob() if agent.fused is not None else pack_policy(agent.ac, self.device)

    def collect_rollouts(self, T=None):
        '''rap.py:348-470.  Returns (protagonist rollouts, [[adversary index, its rollouts], ...], episode statistics of the batch).'''
        T = self.rollout_steps if T is None else T
        N = self.rollout_batch_size
        eng = self.env.unwrapped.engine
        _, groups, starts, ends = draw_adversaries(self._adv_rng, self.num_adversaries, N)
        self.env.reset()                                                 # rap.py:360: every collect starts from fresh episodes
        snap = eng.get_state()
        blob, dims = self._policy_blob()
        want = ('ret', 'adv', 'cviol', 'mse', 'last', 'episode')
        rollouts, rollouts_adv, stats = {}, {}, {}
        for g, s, e in zip(groups, starts, ends):
            s, e = int(s), int(e)
            adv_blob, adv_dims = self._blob_of(self.adversaries[int(g)])     # (packed blob, (obs_dim, adversary act_dim, hidden))
            for adv_record in (False, True):
                eng.set_state(snap)
                key = (T, adv_record)
                out = eng.rollout(blob, dims, T, activation=self.activation, gamma=self.gamma, gae_lambda=self.gae_lambda, use_gae=self.use_gae,
                                  out=self._rap_out.get(key), want=want, adversary=(adv_blob, adv_dims), adv_record=adv_record)
                self._rap_out[key] = out
                dst = rollouts_adv if adv_record else rollouts
                for k in PLAYER_KEYS + (() if adv_record else SHARED_KEYS):
                    if k not in dst:
                        dst[k] = torch.empty_like(out[k])
                    dst[k][:, s:e] = out[k][:, s:e]
                if not adv_record:
                    for k in EPISODE_KEYS:
                        if k not in stats:
                            stats[k] = torch.empty_like(out[k])
                        stats[k][s:e] = out[k][s:e]
        for k in SHARED_KEYS:
            rollouts_adv[k] = rollouts[k]
        self.total_steps += N * T                                        # rap.py:420
        rollouts['adv'] = normalise_advantages(rollouts['adv'])
        rollouts_adv['adv'] = normalise_advantages(rollouts_adv['adv'])  # over the whole batch, THEN split (rap.py:459-470)
        return rollouts, split_by_adversary(rollouts_adv, groups, starts, ends), stats

    def train_step(self):
        '''rap.py:257-281'''
        self.agent.train()
        for adv in self.adversaries:
            adv.train()
        self.obs_normalizer.unset_read_only()
        start = time.time()
        rollouts, splits, stats = self.collect_rollouts()
        results = dict(self.agent.update(rollouts, self.device, generator=self._perm_gen))
        for idx, adv_rollouts in splits:
            res = self.adversaries[idx].update(adv_rollouts, self.device, generator=self._perm_gen)
            results.update({k + f'_adv{idx}': v for k, v in res.items()})
        self.env.record_rollout(stats)
        results.update({'step': self.total_steps, 'elapsed_time': time.time() - start, 'adv_indices': [idx for idx, _ in splits]})
        return results

    # ------------------------------------------------------------------------------------------
    def save(self, path):
        super().save(path)
        if self.rank == 0:                                               # rap.py:130: 'adversary' is the list of the population's state dicts
            state = torch.load(path, weights_only=False)
            state['adversary'] = [adv.state_dict() for adv in self.adversaries]
            torch.save(state, path)

    def load(self, path):
        super(RARL, self).load(path)                                     # the protagonist, normalisers, counters, env state (PPO.load)
        state = torch.load(path, weights_only=False, map_location=self.device)
        for adv, sd in zip(self.adversaries, state['adversary']):        # rap.py:150-151
            adv.load_state_dict(sd)

    def run(self, env=None, render=False, n_episodes=10, verbose=False, use_adv=False, **kwargs):
        '''rap.py:197-255.  The reference draws an adversary of the population at every STEP of the evaluation episode; the fused evaluation launch
        takes one adversary for its whole horizon, so one is drawn per call.'''
        if use_adv:
            self.adversary = self.adversaries[int(self._adv_rng.randint(self.num_adversaries))]
        try:
            return super().run(env=env, render=render, n_episodes=n_episodes, verbose=verbose, use_adv=use_adv, **kwargs)
        finally:
            self.adversary = self.adversaries[0]

    def log_step(self, results):
        step = results['step']
        self.logger.add_scalars({'step': step, 'time': results['elapsed_time'], 'progress': step / self.max_env_steps}, step, prefix='time')
        self.logger.add_scalars({k: results[k] for k in ['policy_loss', 'value_loss', 'entropy_loss'] if k in results}, step, prefix='loss')
        for idx in results['adv_indices']:                               # rap.py:305-312
            self.logger.add_scalars({k: results[k + f'_adv{idx}'] for k in ['policy_loss', 'value_loss', 'entropy_loss']}, step, prefix=f'loss_adv{idx}')
        if len(self.env.return_queue) > 0:
            r, l = np.asarray(self.env.return_queue), np.asarray(self.env.length_queue)
            self.logger.add_scalars({'ep_length': l.mean(), 'ep_return': r.mean(), 'ep_reward': (r / l).mean()}, step, prefix='stat')
        self.logger.dump_scalars()
