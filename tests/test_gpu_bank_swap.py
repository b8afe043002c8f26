"""The instance bank may be replaced while episodes are running (pcb_upload_instances,
pcb_generate_instances, pcb_generate_instances_device "between steps").  A running episode
must not notice: everything it needs was copied into the per-env state at reset (the spatial
env's component_grid, drawn from the UN-rotated pins, is reconstructed from the state and the
placement orientations).  The next episode of env i takes slot (slot_i + N) % new_count of
the NEW bank -- also when the new bank is smaller than the slot the env is on.  Runs under
PCB_GUARD so an out-of-range bank read that lands in a guard band would be reported."""
import os

import numpy as np
import pytest
import torch

from parity import compare_obs, draw_instance, float_close
from ref_harness import CONFIGS, make_oracle, pick_action

pytestmark = pytest.mark.gpu


@pytest.fixture
def guarded():
    os.environ["PCB_GUARD"] = "1"
    yield
    os.environ.pop("PCB_GUARD", None)


@pytest.mark.parametrize("name,generic", [("S_rect", False), ("B", False), ("S_pin", False),
                                          ("S_spa", False), ("D", False), ("D", True),
                                          ("REF_spatial", False), ("S_wide_spa", False)])
@pytest.mark.parametrize("new_count", [4, 11])
def test_bank_replaced_mid_episode(guarded, name, generic, new_count):
    import pcbgpu

    cfg = CONFIGS[name]
    N, swap_at, total_steps = 6, 2, 40
    bank0 = [draw_instance(cfg, 100 + s) for s in range(3 * N)]
    bank1 = [draw_instance(cfg, 9000 + s) for s in range(new_count)]
    env = pcbgpu.PcbVecEnv(cfg, N, autoreset=True, instances=bank0, force_generic=generic)
    C = env.spec.max_num_components
    orcs = [make_oracle(cfg) for _ in range(N)]
    rngs = [np.random.default_rng(i) for i in range(N)]
    gobs = env.reset()
    torch.cuda.synchronize()
    bank, slot = bank0, list(range(N))
    for i, o in enumerate(orcs):
        compare_obs(gobs, i, o.reset(instance=bank[slot[i]]), C, (name, i, "reset"))
    episodes = 0
    for t in range(total_steps):
        if t == swap_at:  # every env is in the middle of its first episode
            env.set_instances(bank1)
            bank = bank1
        acts = np.zeros((N, 3), np.int32)
        for i, o in enumerate(orcs):
            acts[i] = pick_action(rngs[i], o.action_mask, 0.05)
        gobs, grew, gdone, _ = env.step(torch.from_numpy(acts).cuda())
        torch.cuda.synchronize()
        grew, gdone = grew.cpu().numpy(), gdone.cpu().numpy()
        for i, o in enumerate(orcs):
            ro, rr, rd, _ = o.step(tuple(int(v) for v in acts[i]))
            assert bool(gdone[i]) == rd, (name, i, t)
            assert float_close(float(grew[i]), float(rr)), (name, i, t, grew[i], rr)
            if rd:
                episodes += 1
                slot[i] = (slot[i] + N) % len(bank)
                ro = o.reset(instance=bank[slot[i]])
            compare_obs(gobs, i, ro, C, (name, i, t, "swap" if t >= swap_at else "before"))
    assert episodes >= N
    assert env.check_guards() == 0
    env.close()


@pytest.mark.parametrize("name", ["B", "C", "D"])
def test_device_refresh_mid_episode_keeps_running_episodes_consistent(guarded, name):
    """pcb_generate_instances_device in place (same count) and with a smaller count, mid-rollout:
    the spatial env's component_grid must stay the one drawn at reset, masks must stay the
    convolution of the grid, and nothing may be read outside the new bank."""
    import pcbgpu

    n = 999
    env = pcbgpu.PcbVecEnv(CONFIGS[name], n, autoreset=True, bank_size=4 * n, seed=5)
    first = env.reset()
    cg0 = first["component_grid"].clone() if name == "D" else None
    env.rollout_random(3, seed=1)
    env.generate_instances(4 * n, seed=77, on_device=True)      # in place
    obs, _, done, _ = env.step(env.sample_random_actions(seed=2))
    torch.cuda.synchronize()
    if cg0 is not None:  # nobody finished an episode in 4 steps: still the reset-time drawing
        assert int(done.sum()) == 0
        assert torch.equal(obs["component_grid"], cg0)
    env.generate_instances(n // 3, seed=78, on_device=True)     # smaller bank: slots >= count exist
    env.rollout_random(40, seed=3)
    obs = env.step(env.sample_random_actions(seed=4))[0]
    torch.cuda.synchronize()
    grid = obs["grid"].float()
    assert float(grid.sum()) > 0
    st = env.episode_stats()
    assert st["episodes"] > n
    assert env.check_guards() == 0
    env.close()
