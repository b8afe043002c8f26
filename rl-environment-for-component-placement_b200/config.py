"""env_config dict (the keys of agent/config/*.json "env_config") -> EnvSpec.

Mirrors the constructor validation of the four reference environments and
evaluates the float64 constants of find_reward on the host with the reference's
own expression order (dummy_env_rectangular_pin.py:761-830, 879-909; spatial
variant: dummy_env_rectangular_pin_spatial.py:746, 785-791), so the kernels
receive them as inputs.
"""
from dataclasses import dataclass, field

import numpy as np

from ._lib import ABI_VERSION, PcbConfig

KINDS = {"square": 0, "rectangle": 1, "rectangle_pin": 2, "rectangle_spatial_pin": 3}
# model types of utils/agent/utils.py:326-390 that share an environment class
_ALIASES = {
    "rectangle_factorized": "rectangle",
    "rectangle_factorized_pin": "rectangle_pin",
    "rectangle_pin_attn_component": "rectangle_pin",
    "rectangle_pin_attn_all": "rectangle_pin",
    "rectangle_pin_attn_all_no_grid": "rectangle_pin",
    "rectangle_pin_all_attn_factorized": "rectangle_pin",
}
_REWARD = {"centroid": 0, "beam": 1, "both": 2}


@dataclass
class EnvSpec:
    kind: str
    cfg: dict
    height: int
    width: int
    n_orient: int
    max_num_components: int = 0
    max_num_pins_per_component: int = 0
    max_num_nets: int = 0
    max_pins: int = 0
    max_wirelength: float = 0.0
    max_num_intersections: float = 0
    c: PcbConfig = field(default=None, repr=False)

    @property
    def mask_shape(self):
        return (self.height, self.width) if self.kind == "square" else (
            self.n_orient, self.height, self.width)

    @property
    def action_width(self):
        return 2 if self.kind == "square" else 3

    def generator_ranges(self):
        g = self.cfg
        keys = ("min_component_w", "max_component_w", "min_component_h", "max_component_h",
                "min_num_components", "max_num_components", "min_num_nets", "max_num_nets",
                "min_num_pins_per_net", "max_num_pins_per_net", "net_distribution", "pin_spread")
        return [int(g.get(k, 0)) for k in keys]


def _validate(kind, g):
    """The ValueErrors of the reference constructors (square:67-72, rect:239-251,
    pin:598-641, spatial:572-601), same messages."""
    H, W = g["height"], g["width"]
    if H < 0 or W < 0:
        raise ValueError("Grid size must be greater than 0.")
    if kind == "square":
        if g["component_n"] > H or g["component_n"] > W:
            raise ValueError("Component size must be less than or equal to the grid size.")
        return
    if kind == "rectangle_pin":
        too_big = g["max_component_w"] > W or g["max_component_h"] > H
    else:  # rect:241 and spatial:582 compare w to the height (sic)
        too_big = g["max_component_w"] > H or g["max_component_h"] > W
    if too_big:
        raise ValueError("Component size must be less than or equal to the grid size.")
    if g["min_component_w"] < 1 or g["min_component_h"] < 1:
        raise ValueError("Component size must be greater than 0.")
    if g["max_num_components"] < 1 or g["max_num_components"] > H * W:
        raise ValueError("Number of components must be greater than 0"
                         "and less than or equal to the grid area.")
    if kind == "rectangle":
        return
    bw = g["reward_beam_width"]
    if kind == "rectangle_spatial_pin":
        if g["reward_type"] not in _REWARD:
            raise ValueError("Invalid type of routing for the reward")
        if bw < 2 or bw > g["max_num_pins_per_net"]:
            raise ValueError("The beam width must be an integer greater than 2 and at most "
                             "the maximum number of pins per net.")
        if not isinstance(bw, int):
            raise ValueError("The beam width must be an integer.")
        if not isinstance(g["weight_wirelength"], float):
            raise ValueError("The weight of wirelength must be a float.")
        if g["weight_wirelength"] < 0:
            raise ValueError("The weight of wirelength must be greater than 0.")
        return
    if g["min_num_pins_per_net"] > g["max_num_pins_per_net"]:
        raise ValueError("min_num_pins_per_net must not be greater than max num pins per net")
    if g["min_num_pins_per_net"] < 2:
        raise ValueError("min num pins per net must be less than or equal to 2.")
    if (g["min_num_pins_per_net"] * g["min_num_nets"]
            > g["min_component_w"] * g["min_component_h"] * g["min_num_components"]):
        raise ValueError("min_num_pins_per_net * min_num_nets must be less than or equal to "
                         "the total minimum area covered by the components")
    if bw < 1 or not isinstance(bw, int):
        raise ValueError("Beam width must be a positive integer.")
    if g["reward_type"] not in _REWARD:
        raise ValueError("Reward type must be either 'beam', 'centroid', or 'both'.")


def make_spec(env_config, autoreset=True, obs_ring=1, norm_uses_fma=True):
    g = dict(env_config)
    kind = _ALIASES.get(g["type"], g["type"])
    if kind not in KINDS:
        raise ValueError(f"unknown environment type {g['type']!r}")
    g.setdefault("min_num_pins_per_net", 2)
    g.setdefault("reward_type", "both")
    g.setdefault("reward_beam_width", 2)
    g.setdefault("weight_wirelength", 0.5)
    g.setdefault("weight_num_intersections", 0.5)
    _validate(kind, g)
    H, W = int(g["height"]), int(g["width"])
    c = PcbConfig()
    c.abi_version = ABI_VERSION
    c.kind = KINDS[kind]
    c.height, c.width = H, W
    c.autoreset = 1 if autoreset else 0
    c.obs_ring = int(obs_ring)
    c.norm_uses_fma = 1 if norm_uses_fma else 0
    spec = EnvSpec(kind=kind, cfg=g, height=H, width=W,
                   n_orient={"square": 1, "rectangle": 2}.get(kind, 4), c=c)
    if kind == "square":
        c.component_n = int(g["component_n"])
        return spec
    c.max_num_components = spec.max_num_components = int(g["max_num_components"])
    c.max_component_h, c.max_component_w = int(g["max_component_h"]), int(g["max_component_w"])
    spec.max_num_pins_per_component = c.max_component_h * c.max_component_w  # pin:481
    if kind == "rectangle":
        return spec
    nets, ppn = int(g["max_num_nets"]), int(g["max_num_pins_per_net"])
    c.max_num_nets = spec.max_num_nets = nets
    spec.max_pins = c.max_pins = max(1, min(nets * ppn,
                                            spec.max_num_components * spec.max_num_pins_per_component))
    c.reward_type = _REWARD[g["reward_type"]]
    c.reward_beam_width = int(g["reward_beam_width"])
    ww, wi = g["weight_wirelength"], g["weight_num_intersections"]
    c.weight_wirelength, c.weight_num_intersections = float(ww), float(wi)
    # calculate_upper_bound_wirelength / _intersections (pin:761-830)
    diag = np.linalg.norm(np.array((0, 0)) - np.array((H, W)))
    max_wl = 0.5 * diag * (nets * ppn)
    bound = 0.5 * (ppn ** 2) * nets * (nets - 1)
    if kind == "rectangle_spatial_pin":
        max_wl = max_wl / (H + W)  # spatial:746 (normalised twice, reproduced)
        max_int = bound            # spatial:785-791 keeps the float
    else:
        max_int = int(bound)
    wl_norm = H + W
    by_area = (np.mean([g["min_component_h"], g["max_component_h"]])
               * np.mean([g["min_component_w"], g["max_component_w"]])
               * np.mean([g["min_num_components"], g["max_num_components"]]))
    by_net = (np.mean([g["min_num_pins_per_net"], ppn]) * np.mean([g["min_num_nets"], nets]))
    ni_norm = np.min([by_area, by_net])
    penalty = -ww * (max_wl / wl_norm) - wi * (max_int / ni_norm)  # pin:899-905
    c.wirelength_norm, c.intersections_norm = float(wl_norm), float(ni_norm)
    c.penalty_reward = float(penalty)
    c.penalty_wirelength, c.penalty_intersections = float(max_wl), float(max_int)
    spec.max_wirelength, spec.max_num_intersections = max_wl, max_int
    return spec
