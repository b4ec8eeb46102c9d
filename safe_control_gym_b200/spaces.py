'''Minimal `Box` space (gymnasium is optional; the reference imports `gymnasium.spaces.Box`).

Only what the hot path's callers touch is provided: low/high/shape/dtype, seed(), sample(), contains().
'''

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium is installed
    from gymnasium.spaces import Box as _GymBox
except Exception:  # gymnasium absent (this image)
    _GymBox = None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low) if np.ndim(low) > 0 else np.shape(high)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._np_random = np.random.default_rng()

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._np_random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f'Box({self.low}, {self.high}, {self.shape}, {self.dtype})'


Box = _GymBox if _GymBox is not None else _Box
