"""Generates tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

For each configuration: seeded episodes of the reference environment (instance drawn
by the reference's own reset() after np.random.seed(s); random.seed(s)), a random
valid-action sequence with a few deliberately invalid actions, and after every
reset/step the complete observation in the compact dtypes plus reward, done and info.
The fixtures travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from parity import COMPACT_DTYPE  # noqa: E402
from ref_harness import (CONFIGS, extract_instance, load_reference, make_reference_env,  # noqa: E402
                         pick_action, seed_all)

CASES = {  # name -> (seeds, probability of an invalid action)
    "A": (range(4), 0.0), "S_sq": (range(4), 0.15),
    "B": (range(8), 0.0), "S_rect": (range(8), 0.1),
    "C": (range(8), 0.0), "C_both": (range(8), 0.0), "C_beam": (range(6), 0.0),
    "S_pin": (range(10), 0.1), "S_pin_beam3": (range(8), 0.0), "S_pin_tight": (range(10), 0.05),
    "D": (range(6), 0.0), "S_spa": (range(8), 0.1), "S_spa_tight": (range(8), 0.05),
    "REF_pin": (range(6), 0.0),
}
INST_FIELDS = ("comp_h", "comp_w", "pin_comp", "pin_rx", "pin_ry", "pin_net", "pin_id")


def main():
    mods = load_reference()
    assert mods is not None, "reference not available"
    for name, (seeds, p_bad) in CASES.items():
        cfg = CONFIGS[name]
        env = make_reference_env(mods, cfg)
        out = {"seeds": np.array(list(seeds), np.int64)}
        obs_rows, rewards, dones, wls, nis, actions, ep_start = {}, [], [], [], [], [], []
        width = 2 if cfg["type"] == "square" else 3

        def push(obs):
            for k, v in obs.items():
                obs_rows.setdefault(k, []).append(np.asarray(v).astype(COMPACT_DTYPE[k]))

        for e, s in enumerate(seeds):
            seed_all(s)
            obs = env.reset()
            ep_start.append(len(rewards))
            if cfg["type"] != "square":
                inst = extract_instance(env, cfg["type"])
                for f in INST_FIELDS:
                    out[f"inst{e}_{f}"] = getattr(inst, f)
                out[f"inst{e}_n_nets"] = np.int64(inst.n_nets)
            push(obs)  # row of the reset observation
            rewards.append(0.0); dones.append(0); wls.append(np.nan); nis.append(np.nan)
            actions.append([-1] * width)
            rng = np.random.default_rng(s)
            done = False
            while not done:
                a = pick_action(rng, env.action_mask, p_bad)
                obs, r, done, info = env.step(a)
                push(obs)
                rewards.append(float(r)); dones.append(int(done))
                wls.append(float(info.get("wirelength", np.nan)))
                nis.append(float(info.get("num_intersections", np.nan)))
                actions.append(list(a))
        ep_start.append(len(rewards))
        out.update(ep_start=np.array(ep_start, np.int64), reward=np.array(rewards),
                   done=np.array(dones, np.uint8), wirelength=np.array(wls),
                   num_intersections=np.array(nis), action=np.array(actions, np.int32))
        for k, rows in obs_rows.items():
            if k == "component_grid":  # (n_comp, ...) varies per episode: pad to C rows
                C = cfg["max_num_components"]
                rows = [np.concatenate([r, np.zeros((C - r.shape[0],) + r.shape[1:], r.dtype)], 0)
                        for r in rows]
            out[f"obs_{k}"] = np.stack(rows)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(rewards)} rows, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
