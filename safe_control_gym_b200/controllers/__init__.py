'''Register controllers (safe_control_gym/controllers/__init__.py:29-31): only `ppo` is on the hot path.'''

from safe_control_gym_b200.utils.registration import register

register(idx='ppo',
         entry_point='safe_control_gym_b200.controllers.ppo.ppo:PPO',
         config_entry_point='safe_control_gym_b200.controllers.ppo.ppo:DEFAULT_CONFIG')
