'''safe_control_gym_b200: B200-native batched rollout engine behind safe-control-gym's env / PPO API.

Importing the package registers the environments and the controller, like
safe_control_gym/__init__.py:1-5 does.  (The directory is `safe_control_gym_b200`, with an underscore:
a hyphen cannot appear in an importable package name.)
'''

from safe_control_gym_b200 import envs  # noqa: F401  (registration side effect)
from safe_control_gym_b200 import controllers  # noqa: F401

__all__ = ['envs', 'controllers']
