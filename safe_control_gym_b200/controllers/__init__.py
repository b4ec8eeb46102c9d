'''Register controllers (safe_control_gym/controllers/__init__.py:29-47): `ppo` (the hot path) and its sibling collectors `safe_explorer_ppo`, `rarl`, `rap`, `sac` and `ddpg`.'''

from safe_control_gym_b200.utils.registration import register

register(idx='ppo',
         entry_point='safe_control_gym_b200.controllers.ppo.ppo:PPO',
         config_entry_point='safe_control_gym_b200.controllers.ppo.ppo:DEFAULT_CONFIG')

register(idx='safe_explorer_ppo',
         entry_point='safe_control_gym_b200.controllers.safe_explorer.safe_ppo:SafeExplorerPPO',
         config_entry_point='safe_control_gym_b200.controllers.safe_explorer.safe_ppo:DEFAULT_CONFIG')

register(idx='rarl',
         entry_point='safe_control_gym_b200.controllers.rarl.rarl:RARL',
         config_entry_point='safe_control_gym_b200.controllers.rarl.rarl:DEFAULT_CONFIG')

register(idx='rap',
         entry_point='safe_control_gym_b200.controllers.rarl.rap:RAP',
         config_entry_point='safe_control_gym_b200.controllers.rarl.rap:DEFAULT_CONFIG')

register(idx='sac',
         entry_point='safe_control_gym_b200.controllers.sac.sac:SAC',
         config_entry_point='safe_control_gym_b200.controllers.sac.sac:DEFAULT_CONFIG')

register(idx='ddpg',
         entry_point='safe_control_gym_b200.controllers.ddpg.ddpg:DDPG',
         config_entry_point='safe_control_gym_b200.controllers.ddpg.ddpg:DEFAULT_CONFIG')
