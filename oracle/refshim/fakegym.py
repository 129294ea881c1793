"""Minimal gymnasium stand-in (spaces + Env) for running the reference under stubs. TEST INFRASTRUCTURE."""
import types

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        raise NotImplementedError


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(tuple(shape), np.dtype(dtype))
        self.low = np.full(self.shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype=dtype)
        self.high = np.full(self.shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype=dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def sample(self):
        rng = getattr(self, "_rng", np.random)
        return rng.uniform(self.low, self.high).astype(self.dtype)


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def sample(self):
        return int(getattr(self, "_rng", np.random).randint(self.n))


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))

    def sample(self):
        return (getattr(self, "_rng", np.random).random_sample(self.nvec.shape) * self.nvec).astype(np.int64)


class Dict(Space):
    def __init__(self, spaces_=None):
        super().__init__(None, None)
        self.spaces = dict(spaces_ or {})

    def contains(self, x):
        return isinstance(x, dict) and all(k in self.spaces and self.spaces[k].contains(v) for k, v in x.items())

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


class Env:
    metadata = {}

    def reset(self, *a, **k):
        raise NotImplementedError

    def step(self, *a, **k):
        raise NotImplementedError

    def close(self):
        pass


spaces = types.ModuleType("spaces")
spaces.Space = Space
spaces.Box = Box
spaces.Discrete = Discrete
spaces.MultiDiscrete = MultiDiscrete
spaces.Dict = Dict
_space_mod = types.ModuleType("space")
_space_mod.Space = Space
spaces.space = _space_mod
