"""Build-container check: env-steps/s of the unmodified reference vs the oracle port on
the same loop (random valid policy, reset included), one process, per config."""
import sys, time, random
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from ref_harness import CONFIGS, load_reference, make_reference_env, make_oracle

def rate(env, budget=4.0, seed=0):
    np.random.seed(seed); random.seed(seed)
    rng = np.random.default_rng(seed)
    steps = 0; t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget:
        env.reset(); done = False
        while not done:
            valid = np.flatnonzero(env.action_mask)
            a = np.unravel_index(valid[rng.integers(0, len(valid))], env.action_mask.shape)
            _, _, done, _ = env.step(tuple(int(v) for v in a)); steps += 1
    return steps / (time.perf_counter() - t0)

mods = load_reference()
for name in sys.argv[1:] or ["A", "B", "C", "C_both", "D", "E"]:
    r = rate(make_reference_env(mods, CONFIGS[name])); o = rate(make_oracle(CONFIGS[name]))
    print(f"{name}: reference {r:8.0f} steps/s   oracle port {o:8.0f} steps/s   ratio {o / r:.2f}")
