"""GPU parity on random configurations (tests/fuzz_configs.py: accepted by the unmodified
reference, oracle == reference checked in test_oracle_vs_reference.py): the runtime-shaped
kernels on non-square, wide (64-bit rows), narrow and tiny boards, every reward type, beam
widths 2..4; stepping with given actions (some invalid), auto-reset, and -- for the pin-less
kinds -- the one-launch rollout kernel against the oracle and against per-step launches."""
import pytest
import torch

from fuzz_configs import FUZZ
from parity import run_parity, run_rollout_parity
from ref_harness import CONFIGS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(FUZZ))
def test_random_configuration(name):
    st = {}
    run_parity(name, range(10), extra_steps=2, p_invalid=0.1, stats=st)
    run_parity(name, range(10, 18), autoreset=True, episodes=3, stats=st)
    assert st["reward_inexact"] == 0, st
    if CONFIGS[name]["type"] in ("square", "rectangle"):
        import pcbgpu

        run_rollout_parity(name, N=12, T=24, episodes=8, expect_done=False)  # big boards: no episode ends
        a = pcbgpu.PcbVecEnv(CONFIGS[name], 100, obs_ring=3, seed=1)
        b = pcbgpu.PcbVecEnv(CONFIGS[name], 100, obs_ring=3, seed=1)
        fused = a.fused_rollout(True)
        assert fused == (max(CONFIGS[name]["height"], CONFIGS[name]["width"]) <= 32)
        b.fused_rollout(False)
        a.reset(); b.reset()
        a.rollout_random(17, seed=2); b.rollout_random(17, seed=2)
        for k in a.obs_ring_tensors:
            assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, k)
        a.close(); b.close()
