"""gym.spaces stand-in: Discrete and Box as used by envs/env_wrapper.py:94-104,146-159."""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()
        # test hook: when set, sample() returns sample_hook(self)
        self.sample_hook = None

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]


class Discrete(Space):
    def __init__(self, n):
        assert n >= 0
        self.n = n
        super(Discrete, self).__init__((), np.int64)

    def sample(self):
        if self.sample_hook is not None:
            return int(self.sample_hook(self))
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        if isinstance(x, int):
            as_int = x
        elif isinstance(x, (np.generic, np.ndarray)) and (x.dtype.char in np.typecodes["AllInteger"] and x.shape == ()):
            as_int = int(x)
        else:
            return False
        return 0 <= as_int < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            shape = low.shape
        else:
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        super(Box, self).__init__(shape, dtype)

    def sample(self):
        if self.sample_hook is not None:
            return self.sample_hook(self)
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def __repr__(self):
        return "Box" + str(self.shape)
