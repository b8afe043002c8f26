"""gym.spaces.utils.flatten / flatten_space for Box spaces only."""
import numpy as np


def flatten_space(space):
    from . import Box

    return Box(
        low=np.asarray(space.low).ravel(),
        high=np.asarray(space.high).ravel(),
        shape=(int(np.prod(space.shape)),),
        dtype=space.dtype,
    )


def flatten(space, x):
    return np.asarray(x, dtype=space.dtype).ravel()
