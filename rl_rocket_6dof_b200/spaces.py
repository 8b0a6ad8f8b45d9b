"""Minimal `Box` space used when neither gym nor gymnasium is importable (they are not in this
image).  Same attributes SB3 reads from `gym.spaces.Box` (low, high, shape, dtype, sample, contains)."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low) if np.ndim(low) else np.shape(high)
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def make_box(low, high, shape, dtype=np.float32):
    """gymnasium / gym Box when available (so SB3 type checks pass), else the local stand-in."""
    for mod in ("gymnasium", "gym"):
        try:
            spaces = __import__(mod + ".spaces", fromlist=["Box"])
            return spaces.Box(low=low, high=high, shape=shape, dtype=dtype)
        except Exception:
            continue
    return Box(low, high, shape, dtype)
