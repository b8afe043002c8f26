"""Canonical ``env_config`` blocks.

``CONFIGS``          A-E of BASELINE.json ``configs`` / SURVEY.md section 8 (concrete
                     constructor arguments fixed there), the reward-type variants of C and
                     the spatial twin D.
``SHIPPED_CONFIGS``  the ``env_config`` block of every agent/config/*.json the reference
                     ships (file name -> block), verbatim: what ``create_env`` receives from
                     the reference's PPO entry point (utils/agent/utils.py:204, 275).
``REF_*``            short names for the four distinct shipped environments (benchmarks).

bench.py, examples/ and the tests read the configurations from here.
"""

_PIN_C = dict(type="rectangle_pin", height=32, width=32, net_distribution=9, pin_spread=9,
              min_component_w=2, max_component_w=4, min_component_h=2, max_component_h=4,
              max_num_components=10, min_num_components=10, min_num_nets=5, max_num_nets=5,
              max_num_pins_per_net=6, min_num_pins_per_net=2, reward_type="centroid",
              reward_beam_width=2, weight_wirelength=0.5, weight_num_intersections=0.5)

CONFIGS = {
    "A": dict(type="square", height=8, width=8, component_n=3),
    "A2": dict(type="square", height=8, width=8, component_n=2),
    "B": dict(type="rectangle", height=16, width=16, min_component_w=2, max_component_w=4,
              min_component_h=2, max_component_h=4, max_num_components=6, min_num_components=6),
    "C": dict(_PIN_C),
    "C_both": dict(_PIN_C, reward_type="both"),
    "C_beam": dict(_PIN_C, reward_type="beam"),
    "D": dict(_PIN_C, type="rectangle_spatial_pin"),
    "D_both": dict(_PIN_C, type="rectangle_spatial_pin", reward_type="both"),
    "E": dict(_PIN_C, height=64, width=64, max_num_components=32, min_num_components=32,
              min_num_nets=16, max_num_nets=16),
}

_SHIP_RECT = dict(height=10, width=10, min_component_w=2, min_component_h=2, max_component_w=4,
                  max_component_h=4, max_num_components=20, min_num_components=20)
_SHIP_PIN = dict(height=10, width=10, net_distribution=9, pin_spread=9, min_component_w=2,
                 max_component_w=2, min_component_h=2, max_component_h=2, max_num_components=5,
                 min_num_components=5, min_num_nets=3, max_num_nets=3, max_num_pins_per_net=6,
                 min_num_pins_per_net=6, reward_type="centroid", reward_beam_width=2,
                 weight_wirelength=1.0, weight_num_intersections=0.0)

SHIPPED_CONFIGS = {
    "square_model": dict(type="square", height=10, width=10, component_n=2),
    "rectangle_model": dict(_SHIP_RECT, type="rectangle"),
    "rectangle_model_factorized": dict(_SHIP_RECT, type="rectangle_factorized"),
    "rectangle_pin_model": dict(_SHIP_PIN, type="rectangle_pin", weight_wirelength=0.5,
                                weight_num_intersections=0.5),
    "rectangle_pin_factorized_model": dict(
        _SHIP_PIN, type="rectangle_factorized_pin", max_num_components=2, min_num_components=2,
        min_num_nets=2, max_num_nets=2, max_num_pins_per_net=2, min_num_pins_per_net=2),
    "rectangle_pin_attn_component_model": dict(_SHIP_PIN, type="rectangle_pin_attn_component"),
    "rectangle_pin_attn_component_pin_model": dict(_SHIP_PIN, type="rectangle_pin_attn_all"),
    "rectangle_pin_attn_all_no_grid_model": dict(_SHIP_PIN, type="rectangle_pin_attn_all_no_grid"),
    "rectangle_pin_all_attn_factorized_model": dict(_SHIP_PIN,
                                                    type="rectangle_pin_all_attn_factorized"),
    "rectangle_pin_spatial_model": dict(_SHIP_PIN, type="rectangle_spatial_pin",
                                        weight_wirelength=0.75, weight_num_intersections=0.25),
}

CONFIGS["REF_square"] = dict(SHIPPED_CONFIGS["square_model"])
CONFIGS["REF_rect"] = dict(SHIPPED_CONFIGS["rectangle_model"])
CONFIGS["REF_pin"] = dict(SHIPPED_CONFIGS["rectangle_pin_model"])
CONFIGS["REF_spatial"] = dict(SHIPPED_CONFIGS["rectangle_pin_model"], type="rectangle_spatial_pin")
