'''Soft actor-critic on the batched CUDA env (safe_control_gym/controllers/sac/sac.py) -- the off-policy sibling of the PPO collector
("replay-buffer writes instead of the on-policy buffer").

The reference's train_step (sac.py:267-333) steps `rollout_batch_size` CPU envs once, walks the info dicts on the host to find time-limit
truncations, pushes numpy rows into a host replay buffer and copies every sampled minibatch to the device.  Here one train_step is one
`scg_step` launch for the whole batch; the "true next observation / true mask" correction (sac.py:288-305) is two `torch.where` on the
batched `truncated` / `terminal_obs` tensors the step returns; the transition goes into a replay buffer that lives in HBM
(`DeviceReplayBuffer`), minibatches are gathered on the device, and no loss value is read back before log time.  The hidden width (256,
relu) is outside the fused on-policy collector's range, and an off-policy learner needs a gradient step per env step anyway, so the
step API is the right granularity here.
'''

import os
import time
from collections import defaultdict

import numpy as np
import torch

from safe_control_gym_b200.controllers.ppo.ppo import _Logger
from safe_control_gym_b200.controllers.sac.sac_utils import DeviceReplayBuffer, SACAgent
from safe_control_gym_b200.envs.env_wrappers.record_episode_statistics import VecRecordEpisodeStatistics
from safe_control_gym_b200.envs.env_wrappers.vectorized_env import make_vec_envs
from safe_control_gym_b200.math_and_models.normalization import BaseNormalizer, MeanStdNormalizer, RewardStdNormalizer

# defaults of safe_control_gym/controllers/sac/sac.yaml
DEFAULT_CONFIG = dict(
    hidden_dim=256, activation='relu', norm_obs=False, norm_reward=False, clip_obs=10., clip_reward=10.,
    gamma=0.99, tau=0.005, init_temperature=0.2, use_entropy_tuning=False, target_entropy=None,
    train_interval=100, train_batch_size=64, actor_lr=0.001, critic_lr=0.001, entropy_lr=0.001,
    max_env_steps=1000000, warm_up_steps=1000, rollout_batch_size=4, num_workers=1, max_buffer_size=1000000, deque_size=10, eval_batch_size=10,
    log_interval=0, save_interval=0, num_checkpoints=0, eval_interval=0, eval_save_best=False, tensorboard=False,
)


class OffPolicyController:
    '''What SAC and DDPG share: the batched env, the device replay buffer, the transition bookkeeping, checkpoints, evaluation.'''

    DEFAULTS = DEFAULT_CONFIG

    def __init__(self, env_func, training=True, checkpoint_path='model_latest.pt', output_dir='temp', use_gpu=True, seed=0, rng_mode='philox',
                 device=0, precision=None, **kwargs):
        cfg = dict(self.DEFAULTS)
        cfg.update(kwargs)
        for k, v in cfg.items():
            self.__dict__[k] = v
        self.env_func, self.training, self.checkpoint_path, self.output_dir, self.seed = env_func, training, checkpoint_path, output_dir, seed
        self.device, self.precision, self.rng_mode = torch.device('cuda', int(device)), precision, rng_mode
        self.env = make_vec_envs(env_func, None, self.rollout_batch_size, self.num_workers, seed, rng_mode=rng_mode, device=int(device), precision=precision)
        self.env = VecRecordEpisodeStatistics(self.env, self.deque_size)
        self.eval_env = None
        self.agent = self._make_agent()
        self.agent.to(self.device)
        self.obs_normalizer, self.reward_normalizer = BaseNormalizer(), BaseNormalizer()
        if self.norm_obs:
            self.obs_normalizer = MeanStdNormalizer(shape=self.env.observation_space.shape, clip=self.clip_obs, epsilon=1e-8, device=self.device)
        if self.norm_reward:
            self.reward_normalizer = RewardStdNormalizer(gamma=self.gamma, clip=self.clip_reward, epsilon=1e-8, device=self.device)
        self.logger = _Logger(output_dir)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed) + 4242)
        low, high = self.env.action_space.low, self.env.action_space.high
        self._act_low = torch.as_tensor(low, dtype=torch.float32, device=self.device)
        self._act_high = torch.as_tensor(high, dtype=torch.float32, device=self.device)
        self.total_steps = 0

    def _make_agent(self):
        raise NotImplementedError

    def _explore(self, obs):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------------
    def reset(self):
        if self.training:
            self.env.add_tracker('constraint_violation', 0)
            self.env.add_tracker('constraint_violation', 0, mode='queue')
            self.total_steps = 0
            obs, _ = self.env.reset()
            self.obs = self.obs_normalizer(obs).to(torch.float32) if self.norm_obs else obs
            self.buffer = DeviceReplayBuffer(self.env.observation_space, self.env.action_space, self.max_buffer_size, self.train_batch_size, self.device)

    def close(self):
        self.env.close()
        if self.eval_env is not None:
            self.eval_env.close()
        self.logger.close()

    def save(self, path, save_buffer=True):
        '''sac.py:121-147 (keys of the reference checkpoint).'''
        import random
        state = {'agent': self.agent.state_dict(), 'obs_normalizer': self.obs_normalizer.state_dict(), 'reward_normalizer': self.reward_normalizer.state_dict()}
        if self.training:
            state.update({'total_steps': self.total_steps, 'obs': self.obs.detach().cpu().numpy(),
                          'random_state': {'random': random.getstate(), 'numpy': np.random.get_state(), 'torch': torch.get_rng_state(), 'device_gen': self._gen.get_state()},
                          'env_random_state': self.env.get_env_random_state()})
            if save_buffer:
                state['buffer'] = self.buffer.state_dict()
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        torch.save(state, path)

    def load(self, path):
        import random
        state = torch.load(path, weights_only=False, map_location=self.device)
        self.agent.load_state_dict(state['agent'])
        self.obs_normalizer.load_state_dict(state.get('obs_normalizer', {}))
        self.reward_normalizer.load_state_dict(state.get('reward_normalizer', {}))
        if self.training and 'total_steps' in state:
            self.total_steps = state['total_steps']
            self.obs = torch.as_tensor(state['obs'], dtype=torch.float32, device=self.device)
            rs = state['random_state']
            random.setstate(rs['random']); np.random.set_state(rs['numpy']); torch.set_rng_state(rs['torch'].cpu())
            if 'device_gen' in rs:
                self._gen.set_state(rs['device_gen'].cpu())
            self.env.set_env_random_state(state['env_random_state'])
            if 'buffer' in state:
                self.buffer.load_state_dict(state['buffer'])
            self.logger.load(self.total_steps)

    def learn(self, env=None, **kwargs):
        '''sac.py:149-190.'''
        while self.total_steps < self.max_env_steps:
            results = self.train_step()
            if self.total_steps >= self.max_env_steps or (self.save_interval and self.total_steps % self.save_interval == 0):
                self.save(self.checkpoint_path)
                self.logger.info(f'Checkpoint | {self.checkpoint_path}')
            if self.num_checkpoints and self.total_steps % max(1, self.max_env_steps // self.num_checkpoints) == 0:
                self.save(os.path.join(self.output_dir, 'checkpoints', f'model_{self.total_steps}.pt'), save_buffer=False)
            if self.eval_interval and self.total_steps % self.eval_interval == 0:
                ev = self.run(n_episodes=self.eval_batch_size)
                results['eval'] = ev
                score = float(ev['ep_returns'].mean())
                if self.eval_save_best and getattr(self, 'eval_best_score', -np.inf) < score:
                    self.eval_best_score = score
                    self.save(os.path.join(self.output_dir, 'model_best.pt'), save_buffer=False)
            if self.log_interval and self.total_steps % self.log_interval == 0:
                self.log_step(results)

    def select_action(self, obs, info=None):
        with torch.inference_mode():
            obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
            return self.agent.ac.act(obs, deterministic=True).cpu().numpy()

    # ------------------------------------------------------------------------------------------
    def train_step(self, **kwargs):
        '''One batched env step into the replay buffer (+ `train_interval` gradient steps every `train_interval` env steps, sac.py:316-328).'''
        self.agent.train()
        self.obs_normalizer.unset_read_only()
        obs = self.obs
        start = time.time()
        if self.total_steps < self.warm_up_steps:                    # action_space.sample() per env (sac.py:274-275): uniform in the box
            u = torch.rand(self.rollout_batch_size, self._act_low.numel(), device=self.device, generator=self._gen)
            act = self._act_low + u * (self._act_high - self._act_low)
        else:
            with torch.inference_mode():
                act = self._explore(obs)
        next_obs, rew, done, info = self.env.step(act)
        batch = info['batch']
        next_obs = self.obs_normalizer(next_obs).to(torch.float32) if self.norm_obs else next_obs
        rew = self.reward_normalizer(rew, done).to(torch.float32) if self.norm_reward else rew
        # time truncation is not true termination (sac.py:288-305): the stored next observation is the terminal one and the mask stays 1
        trunc = batch['truncated'].bool() & done
        term_obs = batch['terminal_obs']
        if self.norm_obs:
            self.obs_normalizer.set_read_only()
            term_obs = self.obs_normalizer(term_obs).to(torch.float32)
            self.obs_normalizer.unset_read_only()
        true_next = torch.where(trunc[:, None], term_obs, next_obs)
        true_mask = torch.where(trunc, torch.ones_like(rew), 1.0 - done.float())
        self.buffer.push({'obs': obs, 'act': act, 'rew': rew[:, None], 'next_obs': true_next, 'mask': true_mask[:, None]})
        self.obs = next_obs
        self.total_steps += self.rollout_batch_size
        results = defaultdict(list)
        if self.total_steps > self.warm_up_steps and not self.total_steps % self.train_interval:
            acc = None
            for _ in range(self.train_interval):                      # the ratio of env steps to gradient steps is locked to 1
                res = self.agent.update(self.buffer.sample(self.train_batch_size, generator=self._gen))
                acc = res if acc is None else {k: acc[k] + v for k, v in res.items()}
            results = {k: float(v) / self.train_interval for k, v in acc.items()}     # ONE readback per train_interval updates
        results = dict(results)
        results.update({'step': self.total_steps, 'elapsed_time': time.time() - start})
        return results

    def run(self, env=None, render=False, n_episodes=10, verbose=False, **kwargs):
        '''Deterministic evaluation (sac.py:218-265) on `n_episodes` fresh envs stepped together; the first finished episode of each env counts.'''
        self.agent.eval()
        self.obs_normalizer.set_read_only()
        if self.eval_env is None or self.eval_env.num_envs != n_episodes:
            if self.eval_env is not None:
                self.eval_env.close()
            self.eval_env = make_vec_envs(self.env_func, None, n_episodes, 1, self.seed * 111, rng_mode=self.rng_mode, device=self.device.index, precision=self.precision)
        env = self.eval_env
        obs, _ = env.reset()
        ret = torch.zeros(n_episodes, device=self.device, dtype=torch.float64)
        length, cv, mse = torch.zeros_like(ret), torch.zeros_like(ret), torch.zeros_like(ret)
        alive = torch.ones(n_episodes, dtype=torch.bool, device=self.device)
        for _ in range(int(env.proto.CTRL_STEPS)):
            obs_n = self.obs_normalizer(obs).to(torch.float32) if self.norm_obs else obs
            with torch.inference_mode():
                act = self.agent.ac.act(obs_n, deterministic=True)
            obs, rew, done, info = env.step(act)
            b = info['batch']
            w = alive.double()
            ret += w * rew.double(); length += w; cv += w * b['cviol'].double(); mse += w * b['mse']
            alive &= ~done
        return {'ep_returns': ret.cpu().numpy(), 'ep_lengths': length.cpu().numpy(), 'constraint_violation': cv.cpu().numpy(), 'mse': (mse / length).cpu().numpy()}

    def log_step(self, results):
        step = results['step']
        self.logger.add_scalars({'step': step, 'time': results['elapsed_time'], 'progress': step / self.max_env_steps}, step, prefix='time')
        self.logger.add_scalars({k: results[k] for k in ('policy_loss', 'critic_loss', 'entropy_loss') if k in results}, step, prefix='loss')
        if len(self.env.return_queue) > 0:
            r, l = np.asarray(self.env.return_queue), np.asarray(self.env.length_queue)
            self.logger.add_scalars({'ep_length': l.mean(), 'ep_return': r.mean(), 'ep_reward': (r / l).mean()}, step, prefix='stat')
        self.logger.dump_scalars()


class SAC(OffPolicyController):
    def _make_agent(self):
        return SACAgent(self.env.observation_space, self.env.action_space, hidden_dim=self.hidden_dim, gamma=self.gamma, tau=self.tau,
                        init_temperature=self.init_temperature, use_entropy_tuning=self.use_entropy_tuning, target_entropy=self.target_entropy,
                        actor_lr=self.actor_lr, critic_lr=self.critic_lr, entropy_lr=self.entropy_lr, activation=self.activation)

    def _explore(self, obs):
        return self.agent.ac.act(obs, deterministic=False)           # sac.py:277-278
