"""Generates tests/golden/shipped/*.npz: trajectories of the reference's OWN ``create_env``
on the ``env_config`` block of every agent/config/*.json it ships (build container only).

    python tests/golden/make_shipped_golden.py

utils/agent/utils.py cannot be imported here (it imports ray and tensorflow at module
level), so ``init_env`` and ``create_env`` are cut out of that file with ``ast`` and executed
unmodified against the reference's env classes and flatten wrappers (imported behind the gym
stub).  What is recorded is therefore exactly what ``tune.register_env(model_type,
create_env)`` hands to RLlib: wrapped envs take flat ``Discrete`` actions and return the
flattened mask; the factorized model types keep the tuple action.  Every array is stored in
the dtype the reference returned it in.  Note that ``init_env`` never forwards
``weight_num_intersections`` (utils.py:352-390) -- these fixtures pin that behaviour.
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ref_harness import REF_ROOT, extract_instance, load_reference, seed_all  # noqa: E402

OUT = os.path.join(HERE, "shipped")
INST_FIELDS = ("comp_h", "comp_w", "pin_comp", "pin_rx", "pin_ry", "pin_net", "pin_id")
EPISODES = 4
P_INVALID = 0.08


def reference_create_env(mods):
    """init_env / create_env of utils/agent/utils.py, source taken from the file."""
    path = os.path.join(REF_ROOT, "utils", "agent", "utils.py")
    src = open(path).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("init_env", "create_env")]
    assert len(fns) == 2
    for f in fns:  # the annotations name ray's PPOConfig
        for a in f.args.args:
            a.annotation = None
    ns = {
        "DummyPlacementEnvSquare": mods["square"].DummyPlacementEnv,
        "DummyPlacementEnvRect": mods["rectangle"].DummyPlacementEnv,
        "DummyPlacementEnvRectPin": mods["rectangle_pin"].DummyPlacementEnv,
        "DummyPlacementEnvRectPinSpatial": mods["rectangle_spatial_pin"].DummyPlacementEnv,
    }
    for n in ("FlatteningActionWrapperSquare", "FlatteningActionMaskObservationWrapperSquare",
              "FlatteningActionWrapperRect", "FlatteningActionMaskObservationWrapperRect"):
        ns[n] = getattr(mods["wrappers"], n)
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return ns["create_env"]


def base_env(env):
    while hasattr(env, "env"):
        env = env.env
    return env


def kind_of(cfg):
    t = cfg["type"]
    if t == "square":
        return "square"
    if t in ("rectangle", "rectangle_factorized"):
        return "rectangle"
    return "rectangle_spatial_pin" if t == "rectangle_spatial_pin" else "rectangle_pin"


def main():
    mods = load_reference()
    assert mods is not None, "reference not available"
    create_env = reference_create_env(mods)
    os.makedirs(OUT, exist_ok=True)
    for fname in sorted(os.listdir(os.path.join(REF_ROOT, "agent", "config"))):
        name = fname[:-5]
        cfg = json.load(open(os.path.join(REF_ROOT, "agent", "config", fname)))["env_config"]
        kind = kind_of(cfg)
        env = create_env(cfg)
        base = base_env(env)
        flat = env is not base  # wrapped: flat actions, flattened mask
        out = {"config_json": np.array(json.dumps(cfg)), "flat": np.int64(flat)}
        rows, rewards, dones, wls, nis, actions, ep_start = {}, [], [], [], [], [], []

        def push(obs):
            for k, v in obs.items():
                rows.setdefault(k, []).append(np.asarray(v))

        for e in range(EPISODES):
            seed_all(1000 + e)
            obs = env.reset()
            ep_start.append(len(rewards))
            if kind != "square":
                inst = extract_instance(base, kind)
                for f in INST_FIELDS:
                    out[f"inst{e}_{f}"] = getattr(inst, f)
                out[f"inst{e}_n_nets"] = np.int64(inst.n_nets)
            push(obs)
            rewards.append(0.0); dones.append(0); wls.append(np.nan); nis.append(np.nan)
            actions.append([-1, -1, -1])
            rng = np.random.default_rng(e)
            done = False
            while not done:
                mask = np.asarray(base.action_mask)
                if rng.random() < P_INVALID or not mask.any():
                    t = tuple(int(rng.integers(0, n)) for n in mask.shape)
                else:
                    valid = np.argwhere(mask == 1)
                    t = tuple(int(v) for v in valid[rng.integers(0, len(valid))])
                if flat:
                    a = int(np.ravel_multi_index(t, mask.shape))
                    rec = [a, -1, -1]
                else:
                    a = t
                    rec = list(t)
                obs, r, done, info = env.step(a)
                push(obs)
                rewards.append(float(r)); dones.append(int(done))
                wls.append(float(info.get("wirelength", np.nan)))
                nis.append(float(info.get("num_intersections", np.nan)))
                actions.append(rec)
        ep_start.append(len(rewards))
        out.update(ep_start=np.array(ep_start, np.int64), reward=np.array(rewards),
                   done=np.array(dones, np.uint8), wirelength=np.array(wls),
                   num_intersections=np.array(nis), action=np.array(actions, np.int64),
                   penalty_weight_num_intersections=np.float64(
                       getattr(base, "weight_num_intersections", np.nan)))
        for k, rr in rows.items():
            if k == "component_grid":  # (n_comp, ...) per episode -> pad to C rows
                C = cfg["max_num_components"]
                rr = [np.concatenate([r, np.zeros((C - r.shape[0],) + r.shape[1:], r.dtype)], 0)
                      for r in rr]
            out[f"obs_{k}"] = np.stack(rr)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: type {cfg['type']}, flat={flat}, {len(rewards)} rows, "
              f"wi used by the reference = {out['penalty_weight_num_intersections']}, "
              f"{os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
