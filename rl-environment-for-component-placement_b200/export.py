"""Rollout export in the reference's on-disk format.

``utils/visualization/csv_utils.py:11-25`` pickles ``components`` (a list, per sampled
episode, of the env's ``Component`` objects after the episode) and ``actions`` (per
episode, the list of (orientation, x, y) actions) into ``components.pkl`` /
``actions.pkl``; ``web_app/visualization_grid.py`` loads them to draw placements.  The
objects must unpickle as ``environment.dummy_env_rectangular_pin.Component`` / ``Pin``,
so the stand-in classes below are pickled under that module path (when the reference
package is importable its own classes are used instead).
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

_MOD = "environment.dummy_env_rectangular_pin"


def _classes():
    try:
        mod = __import__(_MOD, fromlist=["Component", "Pin"])
        return mod.Component, mod.Pin, None
    except Exception:  # noqa: BLE001 -- reference (or gym) not installed: stand-ins
        Pin = type("Pin", (), {"__module__": _MOD})
        Component = type("Component", (), {"__module__": _MOD})
        fake = {"environment": types.ModuleType("environment"), _MOD: types.ModuleType(_MOD)}
        fake[_MOD].Component, fake[_MOD].Pin = Component, Pin
        fake["environment"].dummy_env_rectangular_pin = fake[_MOD]
        return Component, Pin, fake


def _objects(decoded, classes):
    """decode_state() entry -> list of Component objects (attributes as in
    dummy_env_rectangular_pin.py:135-154 and :30-55)."""
    Component, Pin = classes
    comps = []
    for cid, c in enumerate(decoded["components"]):
        obj = Component.__new__(Component)
        obj.__dict__.update(h=int(c["h"]), w=int(c["w"]), area=int(c["h"] * c["w"]), comp_id=cid,
                            placed=bool(c["x"] >= 0), position=(int(c["x"]), int(c["y"])), pins=[])
        comps.append(obj)
    for p in decoded["pins"]:
        pin = Pin.__new__(Pin)
        pin.__dict__.update({k: int(v) for k, v in p.items()})
        comps[p["component_id"]].pins.append(pin)
    return comps


def export_rollouts(env_config, num_samples, path, policy=None, device=0, seed=0):
    """Plays ``num_samples`` episodes (one env each, random-valid policy unless ``policy``
    maps an observation dict to int32 actions) and writes components.pkl / actions.pkl."""
    from .vec_env import PcbVecEnv

    env = PcbVecEnv(env_config, num_samples, device=device, autoreset=False,
                    bank_size=num_samples, seed=seed)
    obs = env.reset()
    actions = [[] for _ in range(num_samples)]
    finished = np.zeros(num_samples, bool)
    for t in range(4 * max(env.spec.max_num_components, env.spec.height * env.spec.width)):
        a = env.sample_random_actions(seed=seed) if policy is None else policy(obs)
        a_host = a.cpu().numpy()
        obs, _, done, _ = env.step(a)
        torch.cuda.synchronize(env.device)
        d = done.cpu().numpy().astype(bool)
        for i in np.flatnonzero(~finished):
            actions[i].append([int(v) for v in a_host[i]])
        finished |= d
        if finished.all():
            break
    decoded = env.decode_state()
    env.close()
    Component, Pin, fake = _classes()
    comps = [_objects(dec, (Component, Pin)) for dec in decoded]
    saved = {}
    if fake:  # make the stand-in module path importable while pickling
        for k, v in fake.items():
            saved[k] = sys.modules.get(k)
            sys.modules[k] = v
    try:
        os.makedirs(path, exist_ok=True)
        for name, data in (("components", comps), ("actions", actions)):
            with open(os.path.join(path, name + ".pkl"), "wb") as f:
                pickle.dump(data, f)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return comps, actions
