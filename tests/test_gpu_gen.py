"""K8, the device instance generator: same instances as the host generator (same
per-instance SplitMix64 stream; libm vs libdevice may differ in the last ulp of
exp/log/cos, so a tiny fraction may differ), valid, and usable for stepping."""
import numpy as np
import pytest
import torch

from ref_harness import CONFIGS

pytestmark = pytest.mark.gpu


def play(env, steps):
    """first observations + a short random rollout; returns tensors to compare"""
    obs = {k: v.clone() for k, v in env.reset().items()}
    env.rollout_random(steps, seed=4)
    torch.cuda.synchronize()
    return obs, env.episode_stats()


@pytest.mark.parametrize("name", ["C", "D", "B", "S_pin", "E"])
def test_device_generator_matches_host_generator(name):
    import pcbgpu

    n = 2048 if name != "E" else 512
    host = pcbgpu.PcbVecEnv(CONFIGS[name], n, bank_size=2 * n, seed=77)
    dev = pcbgpu.PcbVecEnv(CONFIGS[name], n, bank_size=n, seed=1)
    dev.generate_instances(2 * n, seed=77, on_device=True)
    steps = 70 if name == "E" else 25  # at least two full episodes
    oh, sh = play(host, steps)
    od, sd = play(dev, steps)
    same = torch.ones(n, dtype=torch.bool, device="cuda")
    for k in oh:
        same &= (oh[k].reshape(n, -1) == od[k].reshape(n, -1)).all(1)
    frac = float(same.float().mean())
    assert frac > 0.995, (name, frac)  # identical instances -> identical first observations
    if frac == 1.0:
        assert sh["episodes"] == sd["episodes"] and np.isclose(sh["sum_return"], sd["sum_return"])
    # every generated env must be playable: all episodes run to the end with valid actions
    assert sd["episodes"] > 0 and sd["sum_length"] / sd["episodes"] >= 1
    host.close(); dev.close()


def test_device_generator_refreshes_in_place():
    import pcbgpu

    env = pcbgpu.PcbVecEnv(CONFIGS["C"], 1024, bank_size=2048, seed=3)
    a = {k: v.clone() for k, v in env.reset().items()}
    env.generate_instances(2048, seed=999, on_device=True)  # new instances, same bank size
    b = env.reset()
    assert not torch.equal(a["all_pins_num_feature"], b["all_pins_num_feature"])
    env.rollout_random(30, seed=1)
    assert env.episode_stats()["episodes"] == 1024 * 3
    env.close()
