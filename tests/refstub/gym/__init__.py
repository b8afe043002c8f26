"""Minimal stand-in for gym==0.22 -- TEST INFRASTRUCTURE ONLY.

The reference environments (PBozmarov/RL-Environment-for-Component-Placement)
import ``gym`` for the ``Env`` base class, the ``spaces`` containers and the
``ObservationWrapper``/``ActionWrapper`` bases.  Real gym
is not installed in the build image, so tests that import the *unmodified*
reference from /root/reference put this directory on ``sys.path`` first.
Nothing in the product package imports it.
"""
from . import spaces  # noqa: F401


class Env(object):
    action_space = None
    observation_space = None

    def reset(self, *a, **k):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, *a, **k):
        return self.env.reset(*a, **k)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    def reset(self, *a, **k):
        return self.observation(self.env.reset(*a, **k))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info

    def observation(self, observation):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError
