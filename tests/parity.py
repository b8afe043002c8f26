"""GPU-vs-oracle parity harness (used by tests/test_gpu_*.py and smoke()).

Both sides get the same instances (drawn by the oracle's generators from seeded
Mersenne-Twister streams, i.e. the reference's own draw) and the same actions
(picked from the oracle's mask).  Integer-valued outputs must be bit-exact;
float64 rewards / infos are compared exactly first and within REL_TOL otherwise.
"""
import numpy as np
import torch

import oracle
from ref_harness import CONFIGS, make_oracle, pick_action, seed_all

REL_TOL = 1e-6  # BASELINE.json north_star: "within 1e-6 relative for any floating-point reward term"

COMPACT_DTYPE = {
    "grid": np.uint8, "action_mask": np.uint8, "all_components_feature": np.float32,
    "component_mask": np.uint8, "placement_mask": np.uint8, "all_pins_num_feature": np.int16,
    "all_pins_cat_feature": np.int16, "pin_grid": np.uint8, "component_grid": np.uint8,
}


def draw_instance(cfg, seed):
    seed_all(seed)
    if cfg["type"] == "rectangle":
        return oracle.generate_rect(cfg)
    return oracle.generate_pin(cfg, spatial=(cfg["type"] == "rectangle_spatial_pin"))


def compact(key, ref_value, n_rows=None):
    """Reference-dtype observation -> the compact contract."""
    v = np.asarray(ref_value)
    out = v.astype(COMPACT_DTYPE[key])
    if key == "component_grid" and n_rows is not None and out.shape[0] < n_rows:
        pad = np.zeros((n_rows - out.shape[0],) + out.shape[1:], out.dtype)
        out = np.concatenate([out, pad], 0)
    return out


def compare_obs(gpu_obs, i, ref_obs, C, ctx):
    for k, v in ref_obs.items():
        want = compact(k, v, C)
        got = gpu_obs[k][i].cpu().numpy()
        assert got.shape == want.shape, (ctx, k, got.shape, want.shape)
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)
            raise AssertionError((ctx, k, "first mismatches", bad[:5].tolist(),
                                  got[tuple(bad[0])], want[tuple(bad[0])]))


def float_close(a, b):
    if a == b:
        return True
    return abs(a - b) <= REL_TOL * max(abs(a), abs(b))


def run_parity(name, seeds, autoreset=False, p_invalid=0.0, episodes=1, extra_steps=1,
               stats=None, **env_kw):
    """Steps len(seeds) GPU envs and len(seeds) oracle envs in lock-step.
    autoreset=True runs `episodes` episodes per env (instances seed + k*N)."""
    import pcbgpu

    cfg = CONFIGS[name]
    seeds = list(seeds)
    N = len(seeds)
    square = cfg["type"] == "square"
    insts = None
    if not square:
        insts = [draw_instance(cfg, s + k * 100003) for k in range(episodes) for s in seeds]
    env = pcbgpu.PcbVecEnv(cfg, N, autoreset=autoreset, instances=insts, **env_kw)
    C = env.spec.max_num_components
    orcs = [make_oracle(cfg) for _ in seeds]
    rngs = [np.random.default_rng(s) for s in seeds]
    gobs = env.reset()
    torch.cuda.synchronize()
    epi = [0] * N
    for i, o in enumerate(orcs):
        ro = o.reset() if square else o.reset(instance=insts[i])
        compare_obs(gobs, i, ro, C, (name, seeds[i], "reset"))
    alive = [True] * N  # oracle env still being compared
    after = [0] * N
    n_steps = exact = inexact = 0
    width = 2 if square else 3
    while any(alive):
        acts = np.zeros((N, width), np.int32)
        for i, o in enumerate(orcs):
            acts[i] = pick_action(rngs[i], o.action_mask, p_invalid)
        gobs, grew, gdone, ginfo = env.step(torch.from_numpy(acts).cuda())
        torch.cuda.synchronize()
        grew, gdone = grew.cpu().numpy(), gdone.cpu().numpy()
        gwl, gni = ginfo["wirelength"].cpu().numpy(), ginfo["num_intersections"].cpu().numpy()
        for i, o in enumerate(orcs):
            if not alive[i]:
                continue
            ctx = (name, seeds[i], epi[i], tuple(acts[i]))
            ro, rr, rd, rinfo = o.step(tuple(int(v) for v in acts[i]))
            n_steps += 1
            assert bool(gdone[i]) == rd, (ctx, "done", gdone[i], rd)
            assert float_close(float(grew[i]), float(rr)), (ctx, "reward", grew[i], rr)
            if float(grew[i]) == float(rr):
                exact += 1
            else:
                inexact += 1
            if rinfo:
                assert float_close(float(gwl[i]), float(rinfo["wirelength"])), (ctx, gwl[i], rinfo)
                assert float_close(float(gni[i]), float(rinfo["num_intersections"])), (ctx, gni[i], rinfo)
            if rd and autoreset:
                epi[i] += 1
                if epi[i] >= episodes:
                    alive[i] = False
                    continue
                ro = o.reset() if square else o.reset(instance=insts[epi[i] * N + i])
                compare_obs(gobs, i, ro, C, ctx + ("autoreset",))
                continue
            compare_obs(gobs, i, ro, C, ctx)
            if rd:
                after[i] += 1
                if after[i] >= extra_steps:
                    alive[i] = False
    if stats is not None:
        stats["steps"] = stats.get("steps", 0) + n_steps
        stats["reward_exact"] = stats.get("reward_exact", 0) + exact
        stats["reward_inexact"] = stats.get("reward_inexact", 0) + inexact
    env.close()
    return n_steps


def run_rollout_parity(name, N=24, T=30, episodes=12, seed=5, expect_done=True):
    """pcb_rollout_random (one launch of rollout_small_kernel for the pin-less kinds) against
    the oracle: the logged actions are replayed through oracle envs and every observation of
    every step (ring of T+1 slices), the per-step rewards / dones and the episode count must
    agree.  Returns the number of env-steps compared."""
    import pcbgpu

    cfg = CONFIGS[name]
    square = cfg["type"] == "square"
    insts = None if square else [draw_instance(cfg, s + k * 100003) for k in range(episodes)
                                 for s in range(N)]
    env = pcbgpu.PcbVecEnv(cfg, N, autoreset=True, instances=insts, obs_ring=T + 1)
    env.reset()
    log = torch.zeros((T, N, env.spec.action_width), dtype=torch.int32, device="cuda")
    rlog = torch.zeros((T, N), dtype=torch.float64, device="cuda")
    dlog = torch.zeros((T, N), dtype=torch.uint8, device="cuda")
    env.rollout_random(T, seed=seed, actions_log=log, reward_log=rlog, done_log=dlog)
    torch.cuda.synchronize()
    acts, rews, dones = log.cpu().numpy(), rlog.cpu().numpy(), dlog.cpu().numpy()
    ring = env.obs_ring_tensors
    C = env.spec.max_num_components
    n_done = 0
    for i in range(N):
        o = make_oracle(cfg)
        o.reset() if square else o.reset(instance=insts[i])
        epi = 0
        for t in range(T):
            ro, rr, rd, _ = o.step(tuple(int(v) for v in acts[t, i]))
            assert float_close(float(rews[t, i]), float(rr)), (name, i, t, rews[t, i], rr)
            assert bool(dones[t, i]) == rd, (name, i, t, "done")
            if rd:
                epi += 1
                n_done += 1
                ro = o.reset() if square else o.reset(instance=insts[(epi * N + i) % len(insts)])
            compare_obs({k: v[t + 1] for k, v in ring.items()}, i, ro, C, (name, i, t))
    st = env.episode_stats()
    assert st["episodes"] == n_done and (n_done > 0 or not expect_done), (st, n_done)
    env.close()
    return N * T
