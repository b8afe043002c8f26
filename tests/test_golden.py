"""Golden trajectories recorded from the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py) replayed on the oracle (CPU) and on the CUDA
environments through the C ABI (gpu).  Integer outputs bit-exact; float64 reward /
info exact on the oracle, within 1e-6 relative (north_star tolerance) on the GPU --
the exact-match count is asserted separately in tests/test_gpu_parity.py."""
import glob
import os

import numpy as np
import pytest

from oracle import Instance
from parity import REL_TOL, compact, float_close
from ref_harness import CONFIGS, make_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
FIELDS = ("comp_h", "comp_w", "pin_comp", "pin_rx", "pin_ry", "pin_net", "pin_id")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = CONFIGS[name]
    n_ep = len(z["seeds"])
    insts = None
    if cfg["type"] != "square":
        insts = [Instance(*[z[f"inst{e}_{f}"] for f in FIELDS], int(z[f"inst{e}_n_nets"]))
                 for e in range(n_ep)]
    obs_keys = [k[4:] for k in z.files if k.startswith("obs_")]
    return z, cfg, n_ep, insts, obs_keys


def test_fixtures_present():
    assert len(NAMES) >= 12


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference(name):
    z, cfg, n_ep, insts, obs_keys = load(name)
    env = make_oracle(cfg)
    C = cfg.get("max_num_components", 0)
    for e in range(n_ep):
        lo, hi = int(z["ep_start"][e]), int(z["ep_start"][e + 1])
        obs = env.reset() if insts is None else env.reset(instance=insts[e])
        for row in range(lo, hi):
            if row > lo:
                obs, r, done, info = env.step(tuple(int(v) for v in z["action"][row]))
                assert r == z["reward"][row] and int(done) == z["done"][row], (name, e, row)
                if info:
                    assert info["wirelength"] == z["wirelength"][row]
                    assert info["num_intersections"] == z["num_intersections"][row]
                else:
                    assert np.isnan(z["wirelength"][row])
            for k in obs_keys:
                assert np.array_equal(compact(k, obs[k], C), z["obs_" + k][row]), (name, e, row, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_reference(name):
    import torch

    import pcbgpu

    z, cfg, n_ep, insts, obs_keys = load(name)
    env = pcbgpu.PcbVecEnv(cfg, n_ep, autoreset=False, instances=insts)
    starts = z["ep_start"]
    lens = [int(starts[e + 1] - starts[e]) for e in range(n_ep)]
    obs = env.reset()
    width = z["action"].shape[1]
    for t in range(max(lens)):
        if t > 0:
            acts = np.zeros((n_ep, width), np.int32)
            for e in range(n_ep):
                if t < lens[e]:
                    acts[e] = z["action"][starts[e] + t]
            obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
            torch.cuda.synchronize()
            rew, done = rew.cpu().numpy(), done.cpu().numpy()
            wl, ni = info["wirelength"].cpu().numpy(), info["num_intersections"].cpu().numpy()
        for e in range(n_ep):
            if t >= lens[e]:
                continue
            row = int(starts[e] + t)
            if t > 0:
                assert int(done[e]) == z["done"][row], (name, e, t)
                assert float_close(float(rew[e]), float(z["reward"][row])), (name, e, t, rew[e])
                if not np.isnan(z["wirelength"][row]):
                    assert float_close(float(wl[e]), float(z["wirelength"][row]))
                    assert float_close(float(ni[e]), float(z["num_intersections"][row]))
            for k in obs_keys:
                got = obs[k][e].cpu().numpy()
                assert np.array_equal(got, z["obs_" + k][row]), (name, e, t, k)
    env.close()
    assert REL_TOL == 1e-6
