"""TEST INFRASTRUCTURE -- gym.spaces stub (see gym/__init__.py)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high = np.asarray(low), np.asarray(high)
        self.shape = self.low.shape if shape is None else shape
        self.dtype = dtype
