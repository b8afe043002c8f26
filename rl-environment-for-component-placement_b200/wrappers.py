"""The reference's flatten wrappers for the gym-compatible (N = 1) environment.

Same class names and behaviour as utils/environment/env_wrappers.py and
environment/wrapper_square.py: the observation wrapper ravels ``action_mask``
C-order, the action wrapper turns ``Discrete(prod)`` into the tuple action.
They wrap anything with ``reset``/``step``/``action_mask`` (duck-typed; gym
itself is not required).  For the batched env use ``PcbVecEnv.action_mask_flat``
and flat actions (decoded inside the step kernel) instead.
"""
import numpy as np

from . import spaces as _spaces


class _Wrapper:
    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, *a, **k):
        return self.env.reset(*a, **k)

    def step(self, action):
        return self.env.step(action)


class FlatteningActionMaskObservationWrapperRect(_Wrapper):
    """env_wrappers.py:9-51 (Rect) / 113-155 (Square) / wrapper_square.py:9-28."""

    def __init__(self, env):
        super().__init__(env)
        if self.observation_space is not None:  # env_wrappers.py:28-33
            self.observation_space = _spaces.flattened_mask_observation_space(self.observation_space)

    def observation(self, observation):
        return {k: (np.asarray(v).ravel() if k == "action_mask" else v)
                for k, v in observation.items()}

    def reset(self, *a, **k):
        return self.observation(self.env.reset(*a, **k))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info


class FlatteningActionWrapperRect(_Wrapper):
    """env_wrappers.py:54-110: flat -> (orientation, x, y)."""

    def __init__(self, env):
        super().__init__(env)
        if self.action_space is not None:  # env_wrappers.py:74-78
            assert hasattr(self.action_space, "spaces"), "the action space must be a Tuple"
            self.factor_sizes = [int(s.n) for s in self.action_space.spaces]
            self.action_space = _spaces.flat_action_space(self.action_space)
        else:
            self.factor_sizes = list(env.action_sizes)
        self.n = int(np.prod(self.factor_sizes))

    def action(self, action):
        o, rem = divmod(int(action), self.factor_sizes[1] * self.factor_sizes[2])
        x, y = divmod(rem, self.factor_sizes[2])
        return [o, x, y]

    def step(self, action):
        return self.env.step(self.action(action))

    def validate_action(self, action):
        return self.env.validate_action(*self.action(action))


class FlatteningActionWrapperSquare(FlatteningActionWrapperRect):
    """env_wrappers.py:158-211 / wrapper_square.py:31-80: flat -> (x, y)."""

    def action(self, action):
        return list(divmod(int(action), self.factor_sizes[1]))


FlatteningActionMaskObservationWrapperSquare = FlatteningActionMaskObservationWrapperRect
FlatteningActionMaskObservationWrapper = FlatteningActionMaskObservationWrapperRect  # wrapper_square.py
FlatteningActionWrapper = FlatteningActionWrapperSquare  # wrapper_square.py
