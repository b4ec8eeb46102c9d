'''PPO with the fused B200 rollout collector.

Same controller API as safe_control_gym/controllers/ppo/ppo.py (`make('ppo', env_func, **config)`, `reset`, `learn`,
`train_step`, `run`, `select_action`, `save`, `load`, `close`).  What changes is `train_step()`: the reference's
host loop of `rollout_steps` x (`ac.step` -> `env.step` -> buffer push) (ppo.py:266-285) plus
`compute_returns_and_advantages` (ppo.py:289-299) is ONE `scg_rollout` launch that leaves the PPOBuffer tensors
(obs, act, rew, mask, v, logp, terminal_v, ret, adv) in HBM; the clipped-surrogate update then reads them in place.

Multi-GPU (one process per GPU, torch.distributed/NCCL): every rank rolls out its own shard of the env batch
(env index offset = rank * shard, so seeds stay `seed + global index`), the shards are exchanged with a single
all-gather of the rollout buffer, and every rank applies the identical update (same permutation seed), which keeps
the replicas in lock-step without a gradient all-reduce.
'''

import os
import time

import numpy as np
import torch

from safe_control_gym_b200.controllers.ppo.ppo_utils import PPOAgent
from safe_control_gym_b200.engine import pack_policy
from safe_control_gym_b200.envs.env_wrappers.record_episode_statistics import VecRecordEpisodeStatistics
from safe_control_gym_b200.envs.env_wrappers.vectorized_env import make_vec_envs
from safe_control_gym_b200.math_and_models.normalization import BaseNormalizer, MeanStdNormalizer, RewardStdNormalizer

# defaults of safe_control_gym/controllers/ppo/ppo.yaml:1-39
DEFAULT_CONFIG = dict(
    hidden_dim=64, activation='tanh', norm_obs=False, norm_reward=False, clip_obs=10, clip_reward=10,
    gamma=0.99, use_gae=False, gae_lambda=0.95, use_clipped_value=False, clip_param=0.2, target_kl=0.01, entropy_coef=0.01,
    opt_epochs=10, mini_batch_size=64, actor_lr=0.0003, critic_lr=0.001, max_grad_norm=0.5,
    max_env_steps=1000000, num_workers=1, rollout_batch_size=4, rollout_steps=100, deque_size=10, eval_batch_size=10,
    log_interval=0, save_interval=0, num_checkpoints=0, eval_interval=0, eval_save_best=False, tensorboard=False,
)


class _Logger:
    '''Stand-in for ExperimentLogger (utils/logging.py is observability only, out of scope): keeps the last scalars.'''

    def __init__(self, output_dir):
        self.output_dir, self.scalars = output_dir, {}

    def info(self, msg):
        print(msg, flush=True)

    def add_scalars(self, data, step, prefix=''):
        for k, v in data.items():
            self.scalars[f'{prefix}/{k}'] = (step, float(v))

    def dump_scalars(self):
        pass

    def load(self, step):
        pass

    def close(self):
        pass


def all_gather_rollout(out):
    '''THE collective of the path: one NCCL all-gather of the flat rollout record (engine.rollout allocates every PPOBuffer key
    inside one buffer).  Returns per-key views shaped [G][T*n][width] -- zero copies on either side.'''
    import torch.distributed as dist
    world = dist.get_world_size()
    flat = out['_flat']
    gathered = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(gathered, flat)
    gathered = gathered.view(world, flat.numel())
    T, n = out['_TN']
    res, o = {}, 0
    for k, w in out['_layout']:
        res[k] = gathered[:, o:o + T * n * w].view(world, T * n, w)
        o += T * n * w
    return res


def rollout_flat_numel(obs_dim, act_dim, T, n, with_ret_adv=True):
    '''float32 elements of one rank's flat rollout record (the layout engine.rollout builds).'''
    return (obs_dim + act_dim + 5 + (2 if with_ret_adv else 0)) * T * n


class PeerGatheredRollout:
    '''Fused rollout + all-gather: the gathered PPOBuffer record [G][flat] lives in peer-mapped memory (torch symmetric
    memory over NVLink / NVSwitch); rank r's collector kernel writes every record entry into slot r of EVERY rank's copy while it
    runs, so the only collective left is a barrier.  Falls back to the NCCL all-gather (all_gather_rollout) when symmetric memory
    cannot be set up (no P2P access, old torch).

        pg = PeerGatheredRollout(flat_numel, device)         # collective: every rank calls it
        out = eng.rollout(..., **pg.rollout_kwargs())         # first call allocates the views on pg's slot
        rec = pg.finish(out)                                  # barrier; per-key views [G][T*n][width] like all_gather_rollout
    '''

    def __init__(self, flat_numel, device, group=None):
        import torch.distributed as dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.numel = int(flat_numel)
        self.slot = ((self.numel * 4 + 127) // 128) * 32          # floats per slot, 128-byte aligned
        self.fused, self.why, self.buf, self.hdl, self.offsets, self.mc_offset = False, '', None, None, [], None
        symm_mem = None

        def agree(ok):                                            # all ranks or none, BEFORE the next collective is entered
            t = torch.tensor([1 if ok else 0], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return int(t) == 1
        # step 1 (local, cannot hang): import + allocation.  A rank that fails here must not leave the others inside rendezvous().
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(self.world * self.slot, dtype=torch.float32, device=device)
            local_ok = True
        except Exception as e:   # noqa: BLE001 -- any failure means "use the NCCL all-gather"
            self.why, local_ok = f'{type(e).__name__}: {e}', False
        if agree(local_ok):
            # step 2 (collective): every rank enters the rendezvous
            try:
                self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in self.hdl.buffer_ptrs]
                mine = ptrs[self.rank]
                self.offsets = [ptrs[g] - mine for g in range(self.world) if g != self.rank]
                mc = int(getattr(self.hdl, 'multicast_ptr', 0) or 0)
                self.mc_offset = (mc - mine) if mc else None      # NVSwitch multicast mapping of the same buffer (multimem.st)
                rdv_ok = True
            except Exception as e:   # noqa: BLE001
                self.why, rdv_ok = f'{type(e).__name__}: {e}', False
            self.fused = agree(rdv_ok)
            if not self.fused and not self.why:
                self.why = 'a peer could not map symmetric memory'
        elif not self.why:
            self.why = 'a peer could not allocate symmetric memory'
        if not self.fused:
            self.hdl, self.offsets, self.mc_offset = None, [], None
            if self.buf is None:
                self.buf = torch.empty(self.world * self.slot, dtype=torch.float32, device=device)
        # multicast needs every rank to have the mapping
        if self.fused and not agree(self.mc_offset is not None):
            self.mc_offset = None
        self.local = self.buf[self.rank * self.slot: self.rank * self.slot + self.numel]
        self._began = False

    def rollout_kwargs(self):
        kw = {'flat': self.local, 'peer_byte_offsets': self.offsets if self.fused else None}
        if self.fused and self.mc_offset is not None and os.environ.get('SCG_NO_MULTICAST') != '1':
            kw['multicast_byte_offset'] = self.mc_offset
        return kw

    def begin(self):
        '''Call BEFORE launching a rollout into the gathered buffer: a stream-ordered barrier after which no rank is still reading
        the previous record (its PPO update has finished on every rank), so the peers' stores of the next rollout cannot race with
        a slower rank's update (write-after-read across GPUs).'''
        if self.fused:
            self.hdl.barrier(channel=1)
        self._began = True

    def finish(self, out):
        import torch.distributed as dist
        if self.fused:
            self.hdl.barrier(channel=0)                           # stream-ordered: every rank's kernel (and its peer stores) is done
        else:
            dist.all_gather_into_tensor(self.buf, self.buf[self.rank * self.slot:(self.rank + 1) * self.slot].clone())
        g = self.buf.view(self.world, self.slot)
        T, n = out['_TN']
        res, o = {}, 0
        for k, w in out['_layout']:
            res[k] = g[:, o:o + T * n * w].view(self.world, T * n, w)
            o += T * n * w
        return res


class PPO:
    '''Proximal policy optimisation on the fused collector.'''

    def __init__(self, env_func, training=True, checkpoint_path='model_latest.pt', output_dir='temp', use_gpu=True, seed=0,
                 rng_mode='philox', device=None, precision=None, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(kwargs)
        for k, v in cfg.items():
            self.__dict__[k] = v
        self.env_func, self.training = env_func, training
        self.checkpoint_path, self.output_dir, self.seed = checkpoint_path, output_dir, seed
        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', 0)) if self.world > 1 else 0
        self.device = torch.device('cuda', int(device))
        if self.rollout_batch_size % self.world != 0:
            raise ValueError('rollout_batch_size must be divisible by the number of ranks')
        self.local_batch = self.rollout_batch_size // self.world
        self.env = make_vec_envs(env_func, None, self.local_batch, self.num_workers, seed, rng_mode=rng_mode, device=int(device),
                                 precision=precision, env_index_offset=self.rank * self.local_batch)
        self.env = VecRecordEpisodeStatistics(self.env, self.deque_size)
        self.eval_env = None
        self.agent = PPOAgent(self.env.observation_space, self.env.action_space, hidden_dim=self.hidden_dim,
                              use_clipped_value=self.use_clipped_value, clip_param=self.clip_param, target_kl=self.target_kl,
                              entropy_coef=self.entropy_coef, actor_lr=self.actor_lr, critic_lr=self.critic_lr,
                              opt_epochs=self.opt_epochs, mini_batch_size=self.mini_batch_size, activation=self.activation,
                              fused_update=cfg.get('fused_update'))
        self.agent.to(self.device)
        self._sync_replica(self.agent)
        # pre-/post-processing (ppo.py:70-75).  Inside the fused collector the statistics are frozen for one rollout and updated after it
        # from the raw streams the kernel records (math_and_models/normalization.py)
        self.obs_normalizer, self.reward_normalizer = BaseNormalizer(), BaseNormalizer()
        if self.norm_obs:
            self.obs_normalizer = MeanStdNormalizer(shape=self.env.observation_space.shape, clip=self.clip_obs, epsilon=1e-8, device=self.device)
        if self.norm_reward:
            self.reward_normalizer = RewardStdNormalizer(gamma=self.gamma, clip=self.clip_reward, epsilon=1e-8, device=self.device)
        self.logger = _Logger(output_dir)
        self._perm_gen = torch.Generator(device=self.device)
        self._perm_gen.manual_seed(int(seed) + 12345)      # identical on every rank
        self.total_steps = 0
        self._rollout_out, self._peers, self._peer = {}, {}, None     # cached record buffers / gathered buffers, keyed on the horizon
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self.precision = precision

    def _sync_replica(self, agent):
        '''Lock-step replicas start from rank 0's parameters: the random initialisation comes from torch's global generator, whose default
        seed is not the same in every process (nothing on this path seeds it -- the reference leaves that to the experiment script).'''
        if self.world > 1:
            import torch.distributed as dist
            with torch.no_grad():
                if agent.fused is not None:
                    dist.broadcast(agent.fused.params, src=0)          # the module's parameters are views into this flat vector
                else:
                    for p_ in agent.ac.parameters():
                        dist.broadcast(p_.data, src=0)

    # ------------------------------------------------------------------------------------------
    def reset(self):
        if self.training:
            self.env.add_tracker('constraint_violation', 0)
            self.env.add_tracker('constraint_violation', 0, mode='queue')
            self.total_steps = 0
            obs, _ = self.env.reset()
            self.obs = self.obs_normalizer(obs).to(torch.float32) if self.norm_obs else obs      # ppo.py:96-97

    def close(self):
        self.env.close()
        if self.eval_env is not None:
            self.eval_env.close()
        self.logger.close()

    def save(self, path):
        '''Checkpoint in the reference's format (ppo.py:112-131).  Multi-rank runs: every rank's env shard (engine state blob + its
        observations) is gathered to rank 0, which writes the file; `load` hands each rank its own slice back.'''
        import random
        state = {'agent': self.agent.state_dict(), 'obs_normalizer': self.obs_normalizer.state_dict(),
                 'reward_normalizer': self.reward_normalizer.state_dict()}
        if self.training:
            shard = {'obs': self.obs.detach().cpu().numpy(), 'env_random_state': self.env.get_env_random_state()}
            shards = [shard]
            if self.world > 1:
                import torch.distributed as dist
                shards = [None] * self.world if self.rank == 0 else None
                dist.gather_object(shard, shards, dst=0)
            if self.rank == 0:
                state.update({'total_steps': self.total_steps,
                              'obs': np.concatenate([sh['obs'] for sh in shards], 0),
                              # the reference's get_random_state() keys (utils/utils.py:109-123) + the minibatch permutation generator
                              'random_state': {'random': random.getstate(), 'numpy': np.random.get_state(), 'torch': torch.get_rng_state(),
                                               'perm': self._perm_gen.get_state()},
                              'env_random_state': [st for sh in shards for st in sh['env_random_state']]})
        if self.rank == 0:
            os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
            torch.save(state, path)

    def load(self, path):
        import random
        state = torch.load(path, weights_only=False, map_location=self.device)
        self.agent.load_state_dict(state['agent'])
        self.obs_normalizer.load_state_dict(state.get('obs_normalizer', {}))
        self.reward_normalizer.load_state_dict(state.get('reward_normalizer', {}))
        if self.training and 'env_random_state' in state and isinstance(state['env_random_state'][0], dict):
            self.total_steps = state['total_steps']
            ers = state['env_random_state']
            if len(ers) != self.world:
                raise ValueError(f'checkpoint holds {len(ers)} env shards, this run has {self.world} ranks')
            n = self.local_batch
            obs = np.asarray(state['obs'])
            if obs.shape[0] != n * self.world:
                raise ValueError('checkpoint batch size does not match rollout_batch_size')
            self.obs = torch.as_tensor(obs[self.rank * n:(self.rank + 1) * n], dtype=torch.float32, device=self.device)
            rs = state['random_state']
            if 'random' in rs:
                random.setstate(rs['random'])
            if 'numpy' in rs:
                np.random.set_state(rs['numpy'])
            if 'torch' in rs:
                torch.set_rng_state(rs['torch'].cpu())
            if 'perm' in rs:
                self._perm_gen.set_state(rs['perm'].cpu())
            self.env.set_env_random_state([ers[self.rank]])
            self.logger.load(self.total_steps)

    def learn(self, env=None, **kwargs):
        '''Training loop with the reference's checkpoint / evaluation / logging cadence (ppo.py:150-191).'''
        if self.num_checkpoints > 0:
            step_interval = np.linspace(0, self.max_env_steps, self.num_checkpoints)
            interval_save = np.zeros_like(step_interval, dtype=bool)
        while self.total_steps < self.max_env_steps:
            results = self.train_step()
            if self.total_steps >= self.max_env_steps or (self.save_interval and self.total_steps % self.save_interval == 0):
                self.save(self.checkpoint_path)                       # latest / final checkpoint (collective when world > 1)
                self.save(os.path.join(self.output_dir, 'checkpoints', f'model_{self.total_steps}.pt'))
                if self.rank == 0:
                    self.logger.info(f'Checkpoint | {self.checkpoint_path}')
            if self.num_checkpoints > 0:
                interval_id = int(np.argmin(np.abs(np.array(step_interval) - self.total_steps)))
                if not interval_save[interval_id]:                    # intermediate checkpoint
                    self.save(os.path.join(self.output_dir, 'checkpoints', f'model_{self.total_steps}.pt'))
                    interval_save[interval_id] = True
            if self.eval_interval and self.total_steps % self.eval_interval == 0:
                ev = self.run(n_episodes=self.eval_batch_size)
                results['eval'] = ev
                if self.rank == 0:
                    self.logger.info('Eval | ep_lengths {:.2f} +/- {:.2f} | ep_return {:.3f} +/- {:.3f}'.format(
                        ev['ep_lengths'].mean(), ev['ep_lengths'].std(), ev['ep_returns'].mean(), ev['ep_returns'].std()))
                score = float(ev['ep_returns'].mean())
                if self.eval_save_best and getattr(self, 'eval_best_score', -np.inf) < score:
                    self.eval_best_score = score
                    self.save(os.path.join(self.output_dir, 'model_best.pt'))
            if self.log_interval and self.total_steps % self.log_interval == 0:
                self.log_step(results)

    def select_action(self, obs, info=None):
        with torch.inference_mode():
            obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
            return self.agent.ac.act(obs).cpu().numpy()

    # ------------------------------------------------------------------------------------------
    def collect(self, T=None, deterministic=False, eps=None):
        '''The fused collector: one kernel launch for the whole horizon (replaces ppo.py:266-299).'''
        T = self.rollout_steps if T is None else T
        blob, dims = self._policy_blob()
        eng = self.env.unwrapped.engine
        kw = self._collector_kwargs()
        cached = self._rollout_out.get(T) if eps is None else None
        if self.world > 1 and eps is None:
            if T not in self._peers:
                # multi-GPU training: the record is written straight into every rank's gathered buffer (fused rollout + all-gather)
                self._peers[T] = PeerGatheredRollout(rollout_flat_numel(dims[0], dims[1], T, eng.N), self.device)
            self._peer = self._peers[T]
            if cached is None:
                kw.update(self._peer.rollout_kwargs())
            self._peer.begin()      # every rank's previous update has consumed the gathered buffer before anyone overwrites it
        want = ('ret', 'adv', 'cviol', 'mse', 'last', 'episode')
        if self.norm_obs or self.norm_reward:
            want += ('raw',)
            kw.update(norm=self._norm_blob(), obs0=self.obs.contiguous() if self.norm_obs else None)
        out = eng.rollout(blob, dims, T, eps=eps, activation=self.activation, gamma=self.gamma, gae_lambda=self.gae_lambda,
                          use_gae=self.use_gae, deterministic=deterministic, out=cached, want=want, **kw)
        if eps is None:
            self._rollout_out[T] = out
        return out

    def _collector_kwargs(self):
        '''Extra `engine.rollout` arguments of a sibling collector (SafeExplorerPPO: the safety layer).'''
        return {}

    def _norm_blob(self):
        '''Frozen normaliser statistics in the layout of scg_rollout_args.d_norm.'''
        od = self.env.observation_space.shape[0]
        b = torch.zeros(2 * od + 5, dtype=torch.float64, device=self.device)
        if self.norm_obs:
            rms = self.obs_normalizer.rms
            b[:od] = rms.mean
            b[od:2 * od] = 1.0 / torch.sqrt(rms.var + self.obs_normalizer.epsilon)
            b[2 * od], b[2 * od + 1] = float(self.obs_normalizer.clip), 1.0
        if self.norm_reward:
            b[2 * od + 2], b[2 * od + 3], b[2 * od + 4] = self.reward_normalizer.scale(), float(self.reward_normalizer.clip), 1.0
        return b

    def _update_normalizers(self, out):
        '''Once per rollout (the reference: once per step, ppo.py:270-271): merge the raw streams of the rollout into the statistics.'''
        if self.norm_obs and not self.obs_normalizer.read_only:
            raw = out['obs_raw'].reshape(-1, out['obs_raw'].shape[-1])
            if self.world > 1:
                import torch.distributed as dist
                n = torch.tensor([raw.shape[0]], dtype=torch.float64, device=self.device)
                s1, s2 = raw.double().sum(0), raw.double().square().sum(0)
                for t in (n, s1, s2):
                    dist.all_reduce(t)
                mean = s1 / n
                self.obs_normalizer.rms.update_from_moments(mean, s2 / n - mean.square(), float(n))
            else:
                self.obs_normalizer.rms.update(raw)
        if self.norm_reward and not self.reward_normalizer.read_only:
            if self.world > 1:
                raise NotImplementedError('norm_reward with several ranks: the running-return statistics are per-env and not reduced yet')
            self.reward_normalizer.update_from_rollout(out['rew_raw'][..., 0], 1.0 - out['mask'][..., 0])

    def _policy_blob(self):
        '''Packed rollout blob: one pack kernel from the flat training vector (fused update) or the torch packer.'''
        if self.agent.fused is not None:
            return self.agent.fused.rollout_blob()
        return pack_policy(self.agent.ac, self.device)

    def train_step(self):
        self.agent.train()
        start = time.time()
        out = self.collect()
        self.obs = out['last_obs']
        self.total_steps += self.rollout_batch_size * self.rollout_steps
        if self.norm_obs or self.norm_reward:
            self.obs_normalizer.unset_read_only()
            self._update_normalizers(out)
        if self.world > 1:
            batch = self._peer.finish(out) if self._peer is not None else all_gather_rollout(out)
        else:
            batch = {k: v for k, v in out.items() if not k.startswith('_')}
        adv = batch['adv']
        rollouts = dict(batch)
        self._ev[0].record()
        rollouts['adv'] = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-6)          # ppo.py:300 (numpy std, ddof=0)
        results = self.agent.update(rollouts, self.device, generator=self._perm_gen, defer_stats=True)
        self._ev[1].record()
        # host-side readbacks only after the update has been launched: the episode statistics of the rollout and the loss sums
        self.env.record_rollout(out)
        results = self.agent.finish_update(results)
        results.update({'step': self.total_steps, 'elapsed_time': time.time() - start, 'update_ms': self._ev[0].elapsed_time(self._ev[1])})
        return results

    def run(self, env=None, render=False, n_episodes=10, verbose=False):
        '''Deterministic evaluation (ppo.py:210-257) as one fused rollout over `n_episodes` fresh envs.'''
        self.agent.eval()
        if self.eval_env is None or self.eval_env.num_envs != n_episodes:
            if self.eval_env is not None:
                self.eval_env.close()
            self.eval_env = make_vec_envs(self.env_func, None, n_episodes, 1, self.seed * 111, rng_mode=self.env.unwrapped.rng_mode,
                                          device=self.device.index, precision=self.precision)
        obs, _ = self.eval_env.reset()
        blob, dims = self._policy_blob()
        T = int(self.eval_env.proto.CTRL_STEPS)
        kw = self._collector_kwargs()
        if self.norm_obs:              # ppo.py:218,229: frozen statistics during evaluation
            self.obs_normalizer.set_read_only()
            kw.update(norm=self._norm_blob(), obs0=self.obs_normalizer.normalize(obs).to(torch.float32).contiguous())
            kw['norm'][2 * dims[0] + 4] = 0.0       # evaluation reports raw rewards
        out = self.eval_env.engine.rollout(blob, dims, T, activation=self.activation, deterministic=True, want=('cviol', 'mse', 'episode'), **kw)
        mask = out['mask'][..., 0]
        first_done = torch.where((mask == 0).any(0), (mask == 0).float().argmax(0), torch.full((n_episodes,), T - 1, device=self.device))
        steps = torch.arange(T, device=self.device)[:, None]
        live = (steps <= first_done[None, :]).float()
        ep_returns = (out['rew'][..., 0] * live).sum(0)
        ep_lengths = first_done + 1
        res = {'ep_returns': ep_returns.cpu().numpy(), 'ep_lengths': ep_lengths.cpu().numpy(),
               'constraint_violation': (out['cviol'][..., 0] * live).sum(0).cpu().numpy(),
               'mse': ((out['mse'][..., 0] * live).sum(0) / ep_lengths).cpu().numpy()}
        return res

    def log_step(self, results):
        step = results['step']
        self.logger.add_scalars({'step': step, 'step_time': results['elapsed_time'], 'progress': step / self.max_env_steps}, step, prefix='time')
        self.logger.add_scalars({k: results[k] for k in ['policy_loss', 'value_loss', 'entropy_loss', 'approx_kl']}, step, prefix='loss')
        if len(self.env.return_queue) > 0:
            r, l = np.asarray(self.env.return_queue), np.asarray(self.env.length_queue)
            self.logger.add_scalars({'ep_length': l.mean(), 'ep_return': r.mean(), 'ep_reward': (r / l).mean()}, step, prefix='stat')
        self.logger.dump_scalars()
