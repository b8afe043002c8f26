"""GymCompatEnv -- one environment with the reference's exact gym surface.

``create_env(env_config)`` is the creator the reference registers with RLlib
(utils/agent/utils.py:394-418); here it returns a GymCompatEnv behind the same
flatten wrappers.  N = 1, numpy in / numpy out, the reference's dtypes (float64;
float32 for the square env) and attribute names (``grid``, ``action_mask``,
``components``, ``current_component``, ``validate_action``), so the reference's
unit tests and random policy run against it unchanged.  All dynamics still run
in the CUDA kernels; the host only casts the compact device tensors to the
reference dtypes.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import spaces as _spaces
from .vec_env import OBS_KEYS, PcbVecEnv
from .wrappers import (FlatteningActionMaskObservationWrapperRect,
                       FlatteningActionMaskObservationWrapperSquare,
                       FlatteningActionWrapperRect, FlatteningActionWrapperSquare)


class GymCompatEnv:
    def __init__(self, env_config, device=0, seed=0, norm_uses_fma=True):
        self._vec = PcbVecEnv(env_config, 1, device=device, autoreset=False, obs_ring=1,
                              bank_size=1, seed=seed, norm_uses_fma=norm_uses_fma)
        self.spec = self._vec.spec
        self.kind = self.spec.kind
        self.height, self.width = self.spec.height, self.spec.width
        self.action_sizes = ([self.height, self.width] if self.kind == "square"
                             else [self.spec.n_orient, self.height, self.width])
        # the reference's gym spaces (square:53-65, rect:193-230, pin:482-542, spatial:487-551)
        self.action_space = _spaces.action_space(self.spec)
        self.observation_space = _spaces.observation_space(self.spec)
        self.max_wirelength = self.spec.max_wirelength
        self.max_num_intersections = self.spec.max_num_intersections
        self.reward_wirelength, self.reward_intersection = -1.0, -1
        self._seed, self._episodes = seed, 0
        self._dtype = np.float32 if self.kind == "square" else np.float64
        self._inst = None

    # -- conversion of the compact device observation to the reference dtypes --
    def _convert(self, obs):
        out = {}
        for k in OBS_KEYS[self.kind]:
            v = obs[k][0].cpu().numpy()
            if k == "all_components_feature" and self._inst is not None:
                f = v.astype(np.float64)
                h = np.asarray(self._inst["comp_h"], np.float64)
                w = np.asarray(self._inst["comp_w"], np.float64)
                f[: len(h), 4] = (h * w) / (self.height * self.width)  # exact float64 ratio
                v = f
            elif k == "component_grid" and self._inst is not None:
                v = v[: len(self._inst["comp_h"])]  # spatial:1680-1685: (len(components), ...)
            out[k] = v.astype(self._dtype)
        self.grid, self.action_mask = out["grid"], out["action_mask"]
        for k, v in out.items():
            setattr(self, k, v)
        return {k: v.copy() for k, v in out.items()}

    def reset(self, *a, instance=None, **k):
        """reset(): a fresh instance from the library generator, or ``instance``
        (dict / object with comp_h, comp_w, pin_* arrays) injected like the
        reference's tests inject ``env.components``."""
        if self.kind != "square":
            if instance is not None:
                self._vec.set_instances([instance])
                self._inst = instance if isinstance(instance, dict) else instance.to_dict()
            else:
                self._vec.generate_instances(1, self._seed + self._episodes)
                self._inst = None
        self._episodes += 1
        obs = self._vec.reset()
        if self.kind != "square" and self._inst is None:
            f = obs["all_components_feature"][0].cpu().numpy()
            n = int((obs["placement_mask"][0] > 0).sum()) if self.kind != "rectangle" else int(
                obs["component_mask"][0].sum())
            self._inst = {"comp_h": f[:n, 0].astype(np.int64), "comp_w": f[:n, 1].astype(np.int64)}
        self._cur = 0
        return self._convert(obs)

    @property
    def components(self):
        if self._inst is None:
            return []
        return [SimpleNamespace(h=int(h), w=int(w), comp_id=i)
                for i, (h, w) in enumerate(zip(self._inst["comp_h"], self._inst["comp_w"]))]

    @property
    def current_component(self):
        comps = self.components
        if self._cur < len(comps):
            return comps[self._cur]
        return SimpleNamespace(h=-1, w=-1, comp_id=-1)

    def validate_action(self, *action):
        try:
            return bool(self.action_mask[tuple(int(v) for v in action)] == 1)
        except IndexError:
            return False

    def step(self, action, verbose=False):
        a = torch.tensor([[int(v) for v in action]], dtype=torch.int32)
        valid = self.validate_action(*action) and all(int(v) >= 0 for v in action)
        obs, reward, done, info = self._vec.step(a)
        torch.cuda.synchronize(self._vec.device)
        if valid and self.kind != "square":
            self._cur += 1
        out = self._convert(obs)
        reward, done = float(reward[0]), bool(done[0])
        info_out = {}
        if self.kind in ("rectangle_pin", "rectangle_spatial_pin") and done:
            self.reward_wirelength = float(info["wirelength"][0])
            self.reward_intersection = float(info["num_intersections"][0])
            info_out = {"wirelength": self.reward_wirelength,
                        "num_intersections": self.reward_intersection}
        return out, reward, done, info_out

    def close(self):
        self._vec.close()


_PIN_TYPES = ("rectangle_pin", "rectangle_factorized_pin", "rectangle_pin_attn_component",
              "rectangle_pin_attn_all", "rectangle_pin_attn_all_no_grid",
              "rectangle_pin_all_attn_factorized", "rectangle_spatial_pin")


def init_env(env_config, forward_weight_num_intersections=False, **kw):
    """utils/agent/utils.py:317-391.

    The reference's ``init_env`` passes 17 of the 18 constructor arguments of the pin
    environments: ``env_config["weight_num_intersections"]`` is never forwarded (:352-390), so
    its environments always run with the constructor default 0.5 whatever the JSON says.
    The default here does the same, so that ``create_env`` returns the reference's rewards on
    the reference's own config files; ``forward_weight_num_intersections=True`` honours the key.
    """
    return GymCompatEnv(init_env_config(env_config, forward_weight_num_intersections), **kw)


def init_env_config(env_config, forward_weight_num_intersections=False):
    """The constructor arguments the reference's ``init_env`` really uses, as an env_config."""
    cfg = dict(env_config)
    if cfg.get("type") in _PIN_TYPES and not forward_weight_num_intersections:
        cfg["weight_num_intersections"] = 0.5  # pin:415 / spatial:415 constructor default
    return cfg


def create_env(env_config, **kw):
    """utils/agent/utils.py:394-418: env + flatten wrappers (factorized model
    types keep the tuple action space, as in the reference)."""
    env = init_env(env_config, **kw)
    kind = env_config["type"]
    if kind == "square":
        return FlatteningActionWrapperSquare(FlatteningActionMaskObservationWrapperSquare(env))
    if kind in ("rectangle", "rectangle_pin", "rectangle_pin_attn_component",
                "rectangle_pin_attn_all", "rectangle_pin_attn_all_no_grid",
                "rectangle_spatial_pin"):
        return FlatteningActionWrapperRect(FlatteningActionMaskObservationWrapperRect(env))
    return env
