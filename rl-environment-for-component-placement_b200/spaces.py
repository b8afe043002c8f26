"""The gym spaces the reference environments and wrappers publish.

``action_space`` / ``observation_space`` of every ``DummyPlacementEnv``
(dummy_env_square.py:53-65, dummy_env_rectangular.py:193-230,
dummy_env_rectangular_pin.py:482-542, dummy_env_rectangular_pin_spatial.py:487-551)
and of the flatten wrappers (utils/environment/env_wrappers.py:28-33, 76-78;
environment/wrapper_square.py) with the reference's shapes, bounds and dtypes.

The space classes come from ``gym`` (or ``gymnasium``) when one of them is importable, so
RLlib sees genuine gym spaces; otherwise the minimal classes below provide the surface the
reference's callers touch (``.n``, ``.spaces``, ``.shape``, ``.dtype``, ``.low``, ``.high``,
``contains``, ``sample``, indexing) with gym-0.22 semantics.
"""
import numpy as np


class _Space:
    shape = None
    dtype = None

    def __contains__(self, x):
        return self.contains(x)


class _Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape, self.dtype = (), np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        if isinstance(x, (np.generic, np.ndarray)):
            if not (np.issubdtype(np.asarray(x).dtype, np.integer) and np.asarray(x).shape == ()):
                return False
            x = int(x)
        elif not isinstance(x, int):
            return False
        return 0 <= x < self.n

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return hasattr(other, "n") and int(other.n) == self.n


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def contains(self, x):  # gym 0.22 Box.contains
        if not isinstance(x, np.ndarray):
            x = np.asarray(x)
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        if np.issubdtype(self.dtype, np.integer):
            return np.random.randint(self.low, self.high + 1).astype(self.dtype)
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class _Tuple(_Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        if isinstance(x, (list, np.ndarray)):
            x = tuple(x)
        return (isinstance(x, tuple) and len(x) == len(self.spaces)
                and all(s.contains(v) for s, v in zip(self.spaces, x)))

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __repr__(self):
        return "Tuple(" + ", ".join(map(repr, self.spaces)) + ")"


class _Dict(_Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, key):
        return self.spaces[key]

    def __iter__(self):
        return iter(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def contains(self, x):
        return (isinstance(x, dict) and set(x.keys()) == set(self.spaces.keys())
                and all(self.spaces[k].contains(v) for k, v in x.items()))

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k}: {v!r}" for k, v in self.spaces.items()) + ")"


def _backend():
    """(Discrete, Box, Tuple, Dict, name) -- gym, gymnasium or the classes above."""
    for mod in ("gym", "gymnasium"):
        try:
            sp = __import__(mod).spaces
            # a partial stand-in (e.g. a test stub) is not taken for the real library
            if all(hasattr(sp, n) for n in ("Discrete", "Box", "Tuple", "Dict")) and hasattr(
                    __import__(mod), "__version__"):
                return sp.Discrete, sp.Box, sp.Tuple, sp.Dict, mod
        except ImportError:
            continue
    return _Discrete, _Box, _Tuple, _Dict, "pcbgpu.spaces"


Discrete, Box, Tuple, Dict, BACKEND = _backend()


def action_space(spec):
    """Tuple(Discrete(H), Discrete(W)) for the square env (square:53-55), otherwise
    Tuple(Discrete(O), Discrete(H), Discrete(W)) with O = 2 (rect:193-199) or 4 (pin:482-488)."""
    sizes = ([spec.height, spec.width] if spec.kind == "square"
             else [spec.n_orient, spec.height, spec.width])
    return Tuple([Discrete(n) for n in sizes])


def observation_space(spec):
    """The reference's Dict of Boxes: same keys, shapes, bounds and dtypes, including the
    spatial env's ``all_pins_cat_feature`` declared int32 although the env emits float64
    (spatial:537-545)."""
    H, W, f64 = spec.height, spec.width, np.float64
    if spec.kind == "square":  # square:56-65
        return Dict({"grid": Box(0, 1, (H, W), np.float32),
                     "action_mask": Box(0, 1, (H, W), np.float32)})
    C = spec.max_num_components
    if spec.kind == "rectangle":  # rect:202-230
        return Dict({
            "grid": Box(0, 1, (H, W), f64),
            "action_mask": Box(0, 1, (2, H, W), f64),
            "all_components_feature": Box(-1, max(H, W), (C, 5), f64),
            "component_mask": Box(0, 1, (C,), f64),
            "placement_mask": Box(0, 1, (C,), f64),
        })
    P, nets = spec.max_num_pins_per_component, spec.max_num_nets
    if spec.kind == "rectangle_pin":  # pin:504-542
        return Dict({
            "grid": Box(0, 1, (H, W), f64),
            "action_mask": Box(0, 1, (4, H, W), f64),
            "all_components_feature": Box(-1, max(H, W), (C, 5), f64),
            "all_pins_num_feature": Box(-1, max(H, W), (C, P, 4), f64),
            "all_pins_cat_feature": Box(-1, nets, (C, P, 1), f64),
            "placement_mask": Box(0, 3, (C,), f64),
        })
    ppn = int(spec.cfg["max_num_pins_per_net"])  # spatial:487-551
    return Dict({
        "grid": Box(0, 1, (H, W), f64),
        "pin_grid": Box(0, 100, (H, W, nets + 1), f64),
        "component_grid": Box(0, 1, (C, spec.c.max_component_h, spec.c.max_component_w, nets + 1),
                              f64),
        "action_mask": Box(0, 1, (4, H, W), f64),
        "all_components_feature": Box(-1, max(ppn * nets, H * W), (C, 5 + P), f64),
        "all_pins_num_feature": Box(-1, max(H, W), (C * P + 1, 4), f64),
        "all_pins_cat_feature": Box(-1, max(C, nets), (C * P + 1, 2), np.int32),
        "placement_mask": Box(0, 3, (C,), f64),
    })


def flatten_box(box):
    """gym.spaces.utils.flatten_space of a Box: same bounds and dtype, one dimension."""
    n = int(np.prod(box.shape))
    return Box(np.asarray(box.low).reshape(-1)[0] if n else 0,
               np.asarray(box.high).reshape(-1)[0] if n else 0, (n,), box.dtype)


def flattened_mask_observation_space(space):
    """env_wrappers.py:28-33 / wrapper_square.py: ``action_mask`` flattened, other keys kept."""
    return Dict({k: flatten_box(v) if k == "action_mask" else v for k, v in space.spaces.items()})


def flat_action_space(space):
    """env_wrappers.py:76-78: Discrete(prod of the tuple's sizes)."""
    return Discrete(int(np.prod([s.n for s in space.spaces])))
