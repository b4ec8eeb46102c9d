'''normalize_angle + the identity normaliser (safe_control_gym/math_and_models/normalization.py:8-10, 33-60).

The running mean/std normalisers couple all envs every step (normalization.py:106-113,148-159) and are therefore
not available inside the fused collector (ppo.yaml defaults have them off).
'''

import numpy as np


def normalize_angle(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


class BaseNormalizer:
    def __init__(self, read_only=False):
        self.read_only = read_only

    def set_read_only(self):
        self.read_only = True

    def unset_read_only(self):
        self.read_only = False

    def __call__(self, x, *args, **kwargs):
        return x

    def state_dict(self):
        return {}

    def load_state_dict(self, _):
        pass
