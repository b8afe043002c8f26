"""Parity at BASELINE.json's full sizes (4 096 / 65 536 / 131 072 envs per GPU), where a
per-env oracle comparison would take hours: size-independent properties of the domain,
checked on EVERY env with an independent torch implementation, plus the oracle itself
replayed on a subset of envs spread over the whole id range (first, last, around every
CTA / 2^15 boundary).

  * the policy's pick had its bit set in the previous action mask,
  * occupied cells == sum of the placed components' areas (feature rows),
  * action_mask == "window free and inside the board", recomputed from the observation's
    grid with a ones-kernel convolution (what the reference does with scipy:
    dummy_env_rectangular_pin.py:1808-1851), both orientations, orientations 2/3 == 0/1,
  * a finished env shows the first observation of a new episode,
  * rewards: 1.0 per valid step (square, rect); 0.0 until the end and
    -(ww * wirelength + wi * intersections) at the end, float64-exact (pin kinds),
  * episode statistics == number / lengths of the finished episodes.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from parity import compare_obs, float_close
from ref_harness import CONFIGS, make_oracle

pytestmark = pytest.mark.gpu

FULL = [("A", 65536, 14), ("B", 4096, 20), ("B", 65536, 14), ("C", 65536, 23), ("D", 65536, 13),
        ("E", 131072, 35),
        # 2^20 envs on one GPU: 32-bit index arithmetic of the multi-env-per-warp kernel
        ("A", 1048576, 6), ("B", 1048576, 8)]


def window_free(grid, ph, pw):
    """[n, H, W] 0/1 grid -> [n, H, W] uint8: 1 where a ph x pw window is free and in bounds."""
    n, H, W = grid.shape
    out = torch.zeros((n, H, W), dtype=torch.uint8, device=grid.device)
    if ph > H or pw > W or n == 0:
        return out
    s = F.conv2d(grid.float().unsqueeze(1), torch.ones((1, 1, ph, pw), device=grid.device))
    out[:, :H - ph + 1, :W - pw + 1] = (s[:, 0] == 0).to(torch.uint8)
    return out


def expected_masks(kind, cfg, obs):
    """action_mask recomputed from grid + features of the same observation."""
    grid = obs["grid"]
    N, H, W = grid.shape
    if kind == "square":
        n = cfg["component_n"]
        return window_free(grid, n, n)
    feat = obs["all_components_feature"]
    pm = obs["placement_mask"]
    if kind == "rectangle":
        cur = pm.long().sum(1)  # placed components come first
        n_comp = obs["component_mask"].long().sum(1)
    else:
        cur = (pm == 3).long().argmax(1)
        n_comp = (pm != 0).long().sum(1)
        cur = torch.where((pm == 3).any(1), cur, n_comp)
    have = cur < n_comp
    idx = cur.clamp(max=feat.shape[1] - 1)
    hw = feat[torch.arange(N, device=grid.device), idx][:, :2].long()
    O = obs["action_mask"].shape[1]
    want = torch.zeros((N, O, H, W), dtype=torch.uint8, device=grid.device)
    for h in hw[:, 0].unique().tolist():
        for w in hw[:, 1].unique().tolist():
            sel = ((hw[:, 0] == h) & (hw[:, 1] == w) & have).nonzero().squeeze(1)
            if sel.numel() == 0:
                continue
            g = grid[sel]
            m0, m1 = window_free(g, h, w), window_free(g, w, h)
            want[sel, 0], want[sel, 1] = m0, m1
            if O == 4:
                want[sel, 2], want[sel, 3] = m0, m1
    return want


def bank_instance(bank, i, pins):
    n = int(bank.n_comp[i])
    h, w = bank.comp_h[i, :n].astype(np.int64), bank.comp_w[i, :n].astype(np.int64)
    if not pins:
        return oracle.Instance(h, w)
    m = int(bank.n_pins[i])
    f = lambda a: a[i, :m].astype(np.int64)  # noqa: E731
    return oracle.Instance(h, w, f(bank.pin_comp), f(bank.pin_rx), f(bank.pin_ry), f(bank.pin_net),
                           f(bank.pin_id), int(bank.n_nets[i]))


@pytest.mark.parametrize("name,N,T", FULL)
def test_full_size_properties_and_oracle_subset(name, N, T):
    import pcbgpu

    cfg = CONFIGS[name]
    kind = cfg["type"]
    square, pins = kind == "square", kind in ("rectangle_pin", "rectangle_spatial_pin")
    spec = pcbgpu.make_spec(cfg)
    bank = None if square else pcbgpu.generate_bank(spec, 2 * N, seed=77)
    env = pcbgpu.PcbVecEnv(cfg, N, autoreset=True, obs_ring=2, instances=bank)
    dev = env.device
    C = env.spec.max_num_components
    ww, wi = cfg.get("weight_wirelength", 0.0), cfg.get("weight_num_intersections", 0.0)

    # oracle subset: both ends, warp / CTA / 2^15 boundaries, a few random ids
    rng = np.random.default_rng(5)
    ids = sorted(set([0, 1, 3, 4, 7, 8, 31, 32, N // 2 - 1, N // 2, N - 2, N - 1]
                     + [int(v) for v in rng.integers(0, N, 12)]))
    ids_t = torch.tensor(ids, device=dev)
    orcs = [make_oracle(cfg) for _ in ids]
    epi = [0] * len(ids)
    obs = env.reset()
    for j, i in enumerate(ids):
        ro = orcs[j].reset() if square else orcs[j].reset(instance=bank_instance(bank, i, pins))
        compare_obs({k: v[ids_t] for k, v in obs.items()}, j, ro, C, (name, i, "reset"))

    ar = torch.arange(N, device=dev)
    w = env.spec.action_width
    log = torch.zeros((1, N, w), dtype=torch.int32, device=dev)
    n_done = 0
    sum_len = 0
    prev_cells = obs["grid"].long().sum((1, 2))
    for t in range(T):
        prev_mask = obs["action_mask"]  # ring of 2: stays intact while the next slice is written
        env.rollout_random(1, seed=11, actions_log=log)
        a = log[0].long()
        picked = prev_mask[ar, a[:, 0], a[:, 1]] if w == 2 else prev_mask[ar, a[:, 0], a[:, 1], a[:, 2]]
        assert bool((picked == 1).all()), (name, t, "policy picked a masked-out action")
        obs, res = env._obs(), env.results()
        done = res["done"].bool()
        reward = res["reward"]

        # occupancy == placed area
        cells = obs["grid"].long().sum((1, 2))
        if square:
            n = cfg["component_n"]
            assert bool(((cells - prev_cells == n * n) | done).all())
        else:
            feat = obs["all_components_feature"]
            placed = feat[:, :, 2] >= 0
            area = (feat[:, :, 0] * feat[:, :, 1] * placed).sum(1).long()
            assert torch.equal(cells, area), (name, t, "grid cells != placed area")
            if bool(placed.any()):
                ratio_err = feat[:, :, 4][placed] * (cfg["height"] * cfg["width"]) - (feat[:, :, 0] * feat[:, :, 1])[placed]
                assert float(ratio_err.abs().max()) < 1e-3
        # a finished env restarted: empty board
        assert bool((cells[done] == 0).all()), (name, t, "finished env did not restart")
        prev_cells = cells

        # masks recomputed from the grid
        want = expected_masks(kind, cfg, obs)
        got = obs["action_mask"]
        assert torch.equal(got, want if got.dim() == want.dim() else want.reshape(got.shape)), (name, t, "mask")

        # rewards
        if pins:
            assert bool((reward[~done] == 0).all())
            # (a blocked episode reports the penalty with un-normalised bounds instead: pin:899-909)
            routed = done & (res["episode_length"] == cfg["max_num_components"])
            wl, ni = res["wirelength"][routed], res["num_intersections"][routed]
            assert torch.equal(reward[routed], -1.0 * (ww * wl + wi * ni)), (name, t, "reward != -(ww*wl+wi*ni)")
            assert bool((wl >= 0).all()) and bool((ni >= 0).all())
            # pin features: absolute = component position + relative position, or -1
            if kind == "rectangle_pin":
                pn = obs["all_pins_num_feature"].int()       # [N, C, P, 4]
                pos = obs["all_components_feature"][:, :, 2:4].int().unsqueeze(2)  # [N, C, 1, 2]
                used = (pn != 0).any(-1)
                placed_c = (pos[..., 0] >= 0).expand_as(used)
                ok = torch.where(placed_c.unsqueeze(-1), pn[..., 2:4] == pn[..., 0:2] + pos,
                                 pn[..., 2:4] == -1)
                assert bool(ok.all(-1)[used].all()), (name, t, "pin absolute position")
        else:
            assert bool((reward == 1.0).all())
        n_done += int(done.sum())
        sum_len += int(res["episode_length"][done].sum())

        # the oracle on the subset
        acts = a[ids_t].cpu().numpy()
        sub = {k: v[ids_t].cpu() for k, v in obs.items()}
        r_s, d_s = reward[ids_t].cpu().numpy(), done[ids_t].cpu().numpy()
        wl_s, ni_s = res["wirelength"][ids_t].cpu().numpy(), res["num_intersections"][ids_t].cpu().numpy()
        for j, i in enumerate(ids):
            ro, rr, rd, rinfo = orcs[j].step(tuple(int(v) for v in acts[j]))
            ctx = (name, i, t, tuple(acts[j]))
            assert bool(d_s[j]) == rd, ctx
            assert float(r_s[j]) == float(rr), (ctx, r_s[j], rr)
            if rinfo:
                assert float_close(float(wl_s[j]), float(rinfo["wirelength"])), ctx
                assert float_close(float(ni_s[j]), float(rinfo["num_intersections"])), ctx
            if rd:
                epi[j] += 1
                slot = (i + epi[j] * N) % (2 * N)
                ro = orcs[j].reset() if square else orcs[j].reset(instance=bank_instance(bank, slot, pins))
            compare_obs(sub, j, ro, C, ctx)

    st = env.episode_stats()
    assert st["episodes"] == n_done > 0 and st["sum_length"] == sum_len
    assert st["steps"] == float(N) * T
    env.close()
