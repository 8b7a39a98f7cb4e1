"""Minimal observation/action space objects (Box / Tuple) with the attribute surface the
reference exposes through ``gym.spaces`` (``low``, ``high``, ``shape``, ``dtype``, ``sample``,
``contains``, ``seed``).  ``gym`` itself is not a dependency of this package."""

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._rng = None

    @property
    def np_random(self):
        if self._rng is None:
            self.seed()
        return self._rng

    def seed(self, seed=None):
        self._rng = np.random.RandomState(None if seed is None else int(seed) % (2 ** 32))
        return [seed]

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float64):
        low = np.asarray(low, dtype=dtype)
        high = np.asarray(high, dtype=dtype)
        if shape is not None:
            low = np.broadcast_to(low, shape).copy()
            high = np.broadcast_to(high, shape).copy()
        super().__init__(low.shape, dtype)
        self.low, self.high = low, high
        self.bounded_below = np.isfinite(low)
        self.bounded_above = np.isfinite(high)

    def sample(self):
        low = np.where(self.bounded_below, self.low, -1.0)
        high = np.where(self.bounded_above, self.high, 1.0)
        return self.np_random.uniform(low, high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x, dtype=self.dtype)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f'Box({self.low.min() if self.low.size else 0.0}, {self.high.max() if self.high.size else 0.0}, {self.shape}, {self.dtype})'

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.array_equal(self.low, other.low) \
            and np.array_equal(self.high, other.high)


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        try:
            value = int(x)
        except (TypeError, ValueError):
            return False
        return 0 <= value < self.n and value == x

    def __repr__(self):
        return f'Discrete({self.n})'

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n


class Tuple(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = tuple(spaces)

    def seed(self, seed=None):
        return [s for space in self.spaces for s in space.seed(seed)]

    def sample(self):
        return tuple(space.sample() for space in self.spaces)

    def contains(self, x):
        return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and all(
            space.contains(part) for space, part in zip(self.spaces, x))

    def __getitem__(self, index):
        return self.spaces[index]

    def __len__(self):
        return len(self.spaces)

    def __repr__(self):
        return 'Tuple(' + ', '.join(map(repr, self.spaces)) + ')'
