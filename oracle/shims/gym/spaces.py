import numpy as np


class Box:
    """Only what rl_env/WRSN.py:31-32,299 touches: .low/.high arrays of `shape`."""

    def __init__(self, low, high, shape=None, dtype=np.float64):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
