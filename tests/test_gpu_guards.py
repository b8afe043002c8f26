"""Out-of-bounds stores: compute-sanitizer is not available on the test pool, so the library
can allocate every device buffer between two guard bands (PCB_GUARD=1) and verify them.
Every kernel family is driven with an env count that is not a multiple of the CTA size and
the guards must stay intact; one deliberate stray store shows that the check does detect."""
import ctypes
import os

import pytest
import torch

from ref_harness import CONFIGS

pytestmark = pytest.mark.gpu

NAMES = ["A", "A2", "B", "C", "C_both", "D", "E", "S_sq", "S_rect", "S_rect_many", "S_pin_tight",
         "S_spa", "S_wide_pin", "S_wide_spa", "S_wide_sq", "S_wide_rect", "REF_pin", "REF_spatial",
         "REF_rect", "FZ_rect_pin_0", "FZ_rect_spatial_pin_4", "FZ_rect_4", "FZ_square_0"]


@pytest.fixture
def guarded():
    os.environ["PCB_GUARD"] = "1"
    yield
    os.environ.pop("PCB_GUARD", None)


@pytest.mark.parametrize("name", NAMES)
def test_no_store_leaves_its_buffer(guarded, name):
    import pcbgpu

    n = 333 if name != "E" else 77
    for autoreset in (True, False):
        for generic in (False, True):
            env = pcbgpu.PcbVecEnv(CONFIGS[name], n, autoreset=autoreset, obs_ring=3, seed=2,
                                   bank_size=2 * n, force_generic=generic)
            env.reset()
            env.rollout_random(23, seed=1)
            if env.kind in ("rectangle_pin", "rectangle_spatial_pin"):   # both rollout paths of the pin envs
                env.fused_rollout(False)
                env.rollout_random(9, seed=6)
                env.fused_rollout(True)
            log = torch.zeros((5, n, env.spec.action_width), dtype=torch.int32, device="cuda")
            rl = torch.zeros((5, n), dtype=torch.float64, device="cuda")
            dl = torch.zeros((5, n), dtype=torch.uint8, device="cuda")
            env.rollout_random(5, seed=2, actions_log=log, reward_log=rl, done_log=dl)
            for t in range(4):
                env.step(env.sample_random_actions(seed=3))
            env.reset(mask=(torch.arange(n, device="cuda") % 3 == 0))
            if env.kind != "square":
                env.generate_instances(2 * n, seed=5, on_device=True)
                env.rollout_random(7, seed=4)
            assert env.check_guards() == 0, (name, autoreset, generic)
            env.close()


def test_a_stray_store_is_detected(guarded):
    import pcbgpu

    env = pcbgpu.PcbVecEnv(CONFIGS["B"], 64, seed=1)
    env.reset()
    assert env.check_guards() == 0
    grid = env.obs_ring_tensors["grid"]
    cudart = ctypes.CDLL("libcudart.so.12")
    end = grid.data_ptr() + grid.numel() * grid.element_size()
    assert cudart.cudaMemset(ctypes.c_void_p(end + 5), 0, ctypes.c_size_t(3)) == 0
    assert env.check_guards() == 3
    env.close()


def test_guards_are_off_by_default():
    import pcbgpu

    os.environ.pop("PCB_GUARD", None)
    env = pcbgpu.PcbVecEnv(CONFIGS["A"], 8)
    with pytest.raises(pcbgpu._lib.PcbError):
        env.check_guards()
    env.close()
