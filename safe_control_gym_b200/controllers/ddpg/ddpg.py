'''Deep deterministic policy gradient on the batched CUDA env (safe_control_gym/controllers/ddpg/ddpg.py): the same device-resident
collector as SAC (controllers/sac/sac.py: one `scg_step` launch per train_step, replay buffer in HBM, truncation-aware next observation),
with the deterministic actor + exploration noise of ddpg.py:272-281.'''

import torch

from safe_control_gym_b200.controllers.ddpg.ddpg_utils import DDPGAgent, make_action_noise_process
from safe_control_gym_b200.controllers.sac.sac import DEFAULT_CONFIG as SAC_DEFAULTS
from safe_control_gym_b200.controllers.sac.sac import OffPolicyController

# defaults of safe_control_gym/controllers/ddpg/ddpg.yaml
DEFAULT_CONFIG = {k: v for k, v in SAC_DEFAULTS.items() if k not in ('init_temperature', 'use_entropy_tuning', 'target_entropy', 'entropy_lr')}
DEFAULT_CONFIG.update(warm_up_steps=10000, random_process=dict(func='OrnsteinUhlenbeckProcess', std=dict(func='LinearSchedule', args=0.2)))


class DDPG(OffPolicyController):
    DEFAULTS = DEFAULT_CONFIG

    def _make_agent(self):
        return DDPGAgent(self.env.observation_space, self.env.action_space, hidden_dim=self.hidden_dim, gamma=self.gamma, tau=self.tau,
                         actor_lr=self.actor_lr, critic_lr=self.critic_lr, activation=self.activation)

    def reset(self):
        super().reset()
        self.noise_process = None
        if self.training and self.random_process:
            self.noise_process = make_action_noise_process(self.random_process, self.rollout_batch_size, self.env.action_space, self.device, self._gen)

    def _explore(self, obs):
        act = self.agent.ac.act(obs)
        if self.noise_process is not None:                           # ddpg.py:278-280 (the env clips to its physical bounds)
            act = act + self.noise_process.sample()
        return act
