'''Register environments (safe_control_gym/envs/__init__.py:5-11): same ids, engine-backed classes.'''

from safe_control_gym_b200.utils.registration import register

register(idx='cartpole',
         entry_point='safe_control_gym_b200.envs.gym_control.cartpole:CartPole',
         config_entry_point='safe_control_gym_b200.envs.gym_control.cartpole:DEFAULT_CONFIG')

register(idx='quadrotor',
         entry_point='safe_control_gym_b200.envs.gym_pybullet_drones.quadrotor:Quadrotor',
         config_entry_point='safe_control_gym_b200.envs.gym_pybullet_drones.quadrotor:DEFAULT_CONFIG')
