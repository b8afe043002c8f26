"""Flatten wrappers restated as pure functions (oracle; test infra).

environment/wrapper_square.py:31-80 and utils/environment/env_wrappers.py:
54-110, 158-211 (flat action -> tuple) and :9-51, 113-155 (mask -> 1-D).
"""
import numpy as np


def flat_to_tuple(a, shape):
    """shape = (H, W) square or (O, H, W) rect/pin.  env_wrappers.py:80-98."""
    if len(shape) == 2:
        return list(divmod(int(a), shape[1]))
    o, rem = divmod(int(a), shape[1] * shape[2])
    x, y = divmod(rem, shape[2])
    return [o, x, y]


def tuple_to_flat(t, shape):
    if len(shape) == 2:
        return int(t[0]) * shape[1] + int(t[1])
    return (int(t[0]) * shape[1] + int(t[1])) * shape[2] + int(t[2])


def flatten_mask(mask):
    """C-order ravel, env_wrappers.py:35-51."""
    return np.asarray(mask).ravel()
