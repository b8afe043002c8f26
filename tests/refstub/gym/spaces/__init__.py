"""gym.spaces subset used by the reference envs (see gym/__init__.py header)."""
import numpy as np

from . import utils  # noqa: F401  (gym.spaces.utils.flatten / flatten_space)

_tuple = tuple  # the name ``tuple`` is rebound to a fake submodule below


class Space(object):
    pass


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and 0 <= xi < self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = _tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape)

    def contains(self, x):
        x = np.asarray(x)
        return (
            x.shape == self.shape
            and np.can_cast(x.dtype, self.dtype)
            and bool(np.all(x >= self.low))
            and bool(np.all(x <= self.high))
        )

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = _tuple(spaces)

    def sample(self):
        return _tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(
            s.contains(v) for s, v in zip(self.spaces, x)
        )

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, key):
        return self.spaces[key]

    def contains(self, x):
        return set(x.keys()) == set(self.spaces.keys()) and all(
            self.spaces[k].contains(v) for k, v in x.items()
        )

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


# the TF models test ``isinstance(space, gym.spaces.tuple.Tuple)``
class _Mod(object):
    pass


tuple = _Mod()  # noqa: A001
tuple.Tuple = Tuple
