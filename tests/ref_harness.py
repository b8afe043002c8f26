"""Import the UNMODIFIED reference (test infrastructure, build container only).

The reference lives read-only at /root/reference (or $PCB_REF) and needs
``gym``; tests/refstub provides the few gym names it touches.  On the GPU box
the reference is absent: ``load_reference()`` returns None and tests marked
``ref`` skip; the golden fixtures under tests/golden/ stand in for it there.
"""
import importlib
import os
import random
import sys

import numpy as np

REF_ROOT = os.environ.get("PCB_REF", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstub")

from pcbgpu.configs import CONFIGS as _CANONICAL, SHIPPED_CONFIGS  # noqa: E402,F401

# canonical configs A-E, C_both/C_beam, D, D_both, REF_*: pcbgpu/configs.py (the package owns them;
# bench.py and examples/ read them from there too)
CONFIGS = {k: dict(v) for k, v in _CANONICAL.items()}
CONFIGS.update({
    # small / ragged configs for differential testing
    "S_sq": dict(type="square", height=11, width=10, component_n=3),
    "S_rect": dict(type="rectangle", height=12, width=14, min_component_w=1, max_component_w=5,
                   min_component_h=2, max_component_h=4, max_num_components=9,
                   min_num_components=3),
    "S_rect_full": dict(type="rectangle", height=10, width=10, min_component_w=2,
                        max_component_w=4, min_component_h=2, max_component_h=4,
                        max_num_components=20, min_num_components=20),
    "S_pin": dict(type="rectangle_pin", height=12, width=12, net_distribution=5, pin_spread=4,
                  min_component_w=2, max_component_w=4, min_component_h=2, max_component_h=3,
                  max_num_components=6, min_num_components=3, min_num_nets=2, max_num_nets=4,
                  max_num_pins_per_net=6, min_num_pins_per_net=2, reward_type="both",
                  reward_beam_width=2, weight_wirelength=0.5, weight_num_intersections=0.5),
    "S_pin_beam3": dict(type="rectangle_pin", height=16, width=16, net_distribution=9,
                        pin_spread=9, min_component_w=2, max_component_w=4, min_component_h=2,
                        max_component_h=4, max_num_components=8, min_num_components=4,
                        min_num_nets=3, max_num_nets=6, max_num_pins_per_net=8,
                        min_num_pins_per_net=2, reward_type="beam", reward_beam_width=3,
                        weight_wirelength=0.7, weight_num_intersections=0.5),
    "S_pin_tight": dict(type="rectangle_pin", height=8, width=9, net_distribution=2,
                        pin_spread=0, min_component_w=2, max_component_w=4, min_component_h=2,
                        max_component_h=4, max_num_components=8, min_num_components=6,
                        min_num_nets=2, max_num_nets=3, max_num_pins_per_net=5,
                        min_num_pins_per_net=2, reward_type="centroid", reward_beam_width=2,
                        weight_wirelength=0.5, weight_num_intersections=0.5),
})
CONFIGS["S_wide_pin"] = dict(  # 64-bit rows, runtime-shaped kernels, non-square grid
    type="rectangle_pin", height=40, width=50, net_distribution=3, pin_spread=6,
    min_component_w=2, max_component_w=6, min_component_h=1, max_component_h=5,
    max_num_components=12, min_num_components=5, min_num_nets=2, max_num_nets=7,
    max_num_pins_per_net=7, min_num_pins_per_net=2, reward_type="both", reward_beam_width=3,
    weight_wirelength=0.3, weight_num_intersections=0.9)
CONFIGS["S_wide_spa"] = dict(CONFIGS["S_wide_pin"], type="rectangle_spatial_pin", height=64, width=33,
                             reward_type="centroid", reward_beam_width=2)
CONFIGS["S_wide_sq"] = dict(type="square", height=64, width=64, component_n=5)
CONFIGS["S_wide_rect"] = dict(type="rectangle", height=33, width=64, min_component_w=3,
                              max_component_w=9, min_component_h=2, max_component_h=8,
                              max_num_components=14, min_num_components=8)
CONFIGS["S_spa"] = dict(CONFIGS["S_pin"], type="rectangle_spatial_pin")
CONFIGS["S_spa_tight"] = dict(CONFIGS["S_pin_tight"], type="rectangle_spatial_pin")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "environment"))


def load_reference():
    """Returns {'square','rectangle','rectangle_pin','rectangle_spatial_pin',
    'wrappers','wrapper_square'} -> module, or None when the reference is absent."""
    if not reference_available():
        return None
    for p in (REF_ROOT, _STUB):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import gym  # noqa: F401
    except ImportError:
        return None
    names = {
        "square": "environment.dummy_env_square",
        "rectangle": "environment.dummy_env_rectangular",
        "rectangle_pin": "environment.dummy_env_rectangular_pin",
        "rectangle_spatial_pin": "environment.dummy_env_rectangular_pin_spatial",
        "wrappers": "utils.environment.env_wrappers",
        "wrapper_square": "environment.wrapper_square",
    }
    return {k: importlib.import_module(v) for k, v in names.items()}


_PIN_ARGS = ("height", "width", "net_distribution", "pin_spread", "min_component_w",
             "max_component_w", "min_component_h", "max_component_h", "max_num_components",
             "min_num_components", "min_num_nets", "max_num_nets", "max_num_pins_per_net",
             "min_num_pins_per_net", "reward_type", "reward_beam_width", "weight_wirelength",
             "weight_num_intersections")


def make_reference_env(mods, cfg):
    """The reference constructor for an env_config dict (all 18 pin args are
    passed, i.e. weight_num_intersections IS forwarded here)."""
    kind = cfg["type"]
    if kind == "square":
        return mods["square"].DummyPlacementEnv(cfg["height"], cfg["width"], cfg["component_n"])
    if kind == "rectangle":
        return mods["rectangle"].DummyPlacementEnv(*[cfg[k] for k in (
            "height", "width", "min_component_w", "max_component_w", "min_component_h",
            "max_component_h", "max_num_components", "min_num_components")])
    return mods[kind].DummyPlacementEnv(*[cfg[k] for k in _PIN_ARGS])


def make_oracle(cfg):
    import oracle

    kind = cfg["type"]
    if kind == "square":
        return oracle.SquareEnv(cfg["height"], cfg["width"], cfg["component_n"])
    if kind == "rectangle":
        return oracle.RectEnv(*[cfg[k] for k in (
            "height", "width", "min_component_w", "max_component_w", "min_component_h",
            "max_component_h", "max_num_components", "min_num_components")])
    return oracle.PinEnv(*[cfg[k] for k in _PIN_ARGS], spatial=(kind == "rectangle_spatial_pin"))


def seed_all(s):
    np.random.seed(s & 0xFFFFFFFF)
    random.seed(s)


def extract_instance(env, kind):
    """Reference env -> oracle.Instance (SURVEY.md section 8c exchange format)."""
    from oracle import Instance

    h = np.array([c.h for c in env.components], np.int64)
    w = np.array([c.w for c in env.components], np.int64)
    if kind in ("square", "rectangle"):
        return Instance(h, w)
    pins = env.pins
    return Instance(
        h, w,
        np.array([p.component_id for p in pins], np.int64),
        np.array([p.relative_x for p in pins], np.int64),
        np.array([p.relative_y for p in pins], np.int64),
        np.array([p.net_id for p in pins], np.int64),
        np.array([p.pin_id for p in pins], np.int64),
        int(env.num_nets),
    )


def pick_action(rng, mask, p_invalid=0.0):
    """Uniform valid action from a reference-layout mask, or (with probability
    p_invalid) a uniformly random in-range action that may be invalid."""
    mask = np.asarray(mask)
    if rng.random() < p_invalid:
        return tuple(int(rng.integers(0, n)) for n in mask.shape)
    valid = np.argwhere(mask == 1)
    if len(valid) == 0:
        return tuple(0 for _ in mask.shape)
    return tuple(int(v) for v in valid[rng.integers(0, len(valid))])

# the reference's other shipped configurations (agent/config/*.json), all 10x10
# many small components on a small board: more component rows than lanes per env in the
# multi-env-per-warp rollout kernel (8 lanes per env, 20 rows)
CONFIGS["S_rect_many"] = dict(type="rectangle", height=8, width=8, min_component_w=1,
                              max_component_w=2, min_component_h=1, max_component_h=2,
                              max_num_components=20, min_num_components=17)

# random configurations accepted by the reference (tests/golden/make_fuzz_configs.py)
from fuzz_configs import FUZZ  # noqa: E402

CONFIGS.update(FUZZ)
# centroid-reward twins of ragged pin configurations: the multi-step pin kernel
# (pcb_pin_small.cuh) with runtime shapes and 8 / 16 / 32 lanes per env
for _k in ("S_pin", "S_spa", "FZ_rect_pin_0", "FZ_rect_pin_3", "FZ_rect_pin_4", "FZ_rect_pin_6",
           "FZ_rect_spatial_pin_4"):
    CONFIGS[_k + "_cen"] = dict(CONFIGS[_k], reward_type="centroid")
