"""gym 0.21 `spaces.Box` stand-in (test infrastructure only, see gym/__init__.py)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.asarray(low).shape if not np.isscalar(low) else np.asarray(high).shape
        self.shape = tuple(shape)
        low = np.full(self.shape, low) if np.isscalar(low) else np.asarray(low)
        high = np.full(self.shape, high) if np.isscalar(high) else np.asarray(high)
        self.low = low.astype(self.dtype)
        self.high = high.astype(self.dtype)
        self.np_random = np.random.RandomState(seed)

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        # gym 0.21 box.py: uniform in float64 over [low, high], then cast to the space dtype
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype)
            and x.shape == self.shape
            and np.all(x >= self.low)
            and np.all(x <= self.high)
        )
