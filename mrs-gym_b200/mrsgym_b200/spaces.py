"""`Box` for observation_space / action_space (/root/reference/mrsgym/MRS.py:51-52, MRSWrapper.py:37-38).

The reference takes it from `gym.spaces`.  gym / gymnasium are used when importable; otherwise this minimal
stand-in (low, high, shape, dtype, sample, contains) keeps RLlib-style consumers working without the dependency."""
from __future__ import annotations

import numpy as np

try:                                              # pragma: no cover - depends on the environment
    from gym.spaces import Box                    # noqa: F401
except Exception:                                 # noqa: BLE001
    try:
        from gymnasium.spaces import Box          # noqa: F401
    except Exception:                             # noqa: BLE001
        class Box:
            """Minimal gym.spaces.Box: an axis-aligned box in R^n."""

            def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
                self.shape = tuple(shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
                self._rng = np.random.default_rng(seed)

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                x = self._rng.uniform(lo, hi)
                unb = ~np.isfinite(self.low) & ~np.isfinite(self.high)
                x = np.where(unb, self._rng.standard_normal(self.shape), x)
                return x.astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

            __contains__ = contains

            def __repr__(self):
                return 'Box(%s, %s, %s, %s)' % (self.low.min(), self.high.max(), self.shape, self.dtype)

            def __eq__(self, other):
                return (isinstance(other, Box) and self.shape == other.shape and np.array_equal(self.low, other.low)
                        and np.array_equal(self.high, other.high))
