"""create_env on the env_config block of every agent/config/*.json the reference ships.

tests/golden/shipped/*.npz were recorded from the reference's own ``create_env``
(tests/golden/make_shipped_golden.py: init_env + flatten wrappers, flat Discrete actions
where the reference wraps, tuple actions for the factorized model types).  The reference's
init_env never forwards ``weight_num_intersections`` (utils/agent/utils.py:352-390), so its
envs run with 0.5 whatever the JSON says; ``pcbgpu.create_env`` must do the same.
CPU: the oracle's factory + flatten functions reproduce the fixtures, and the host-side
config handling drops the key.  GPU: ``pcbgpu.create_env`` (CUDA kernels behind the C ABI)
reproduces them array for array in the reference's dtypes.
"""
import glob
import json
import os

import numpy as np
import pytest

import oracle
from parity import REL_TOL, float_close

SHIPPED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shipped")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(SHIPPED, "*.npz")))
FIELDS = ("comp_h", "comp_w", "pin_comp", "pin_rx", "pin_ry", "pin_net", "pin_id")


def load(name):
    z = np.load(os.path.join(SHIPPED, name + ".npz"))
    cfg = json.loads(str(z["config_json"]))
    n_ep = len(z["ep_start"]) - 1
    insts = None
    if cfg["type"] != "square":
        insts = [oracle.Instance(*[z[f"inst{e}_{f}"] for f in FIELDS], int(z[f"inst{e}_n_nets"]))
                 for e in range(n_ep)]
    keys = [k[4:] for k in z.files if k.startswith("obs_")]
    return z, cfg, n_ep, insts, keys


def test_all_ten_shipped_configs_have_fixtures():
    import pcbgpu

    assert set(NAMES) == set(pcbgpu.SHIPPED_CONFIGS) and len(NAMES) == 10
    for name in NAMES:
        assert load(name)[1] == pcbgpu.SHIPPED_CONFIGS[name], name


@pytest.mark.parametrize("name", NAMES)
def test_init_env_drops_weight_num_intersections(name):
    """host side of create_env: the spec the kernels get carries 0.5 unless forwarding is asked for"""
    import pcbgpu

    cfg = pcbgpu.SHIPPED_CONFIGS[name]
    if "weight_num_intersections" not in cfg:
        assert pcbgpu.init_env_config(cfg) == cfg
        return
    spec = pcbgpu.make_spec(pcbgpu.init_env_config(cfg))
    assert spec.c.weight_num_intersections == 0.5
    assert spec.c.weight_wirelength == cfg["weight_wirelength"]
    fwd = pcbgpu.make_spec(pcbgpu.init_env_config(cfg, forward_weight_num_intersections=True))
    assert fwd.c.weight_num_intersections == cfg["weight_num_intersections"]
    z = load(name)[0]
    assert float(z["penalty_weight_num_intersections"]) == 0.5  # what the reference env really had


def replay(name, make_env, flat_obs, check_spaces=False):
    z, cfg, n_ep, insts, keys = load(name)
    flat = bool(z["flat"])
    env = make_env(cfg)
    C = cfg.get("max_num_components", 0)
    exact = total = 0
    for e in range(n_ep):
        lo, hi = int(z["ep_start"][e]), int(z["ep_start"][e + 1])
        obs = env.reset() if insts is None else env.reset(instance=insts[e])
        for row in range(lo, hi):
            if row > lo:
                a = z["action"][row]
                act = int(a[0]) if flat else tuple(int(v) for v in a[: (2 if cfg["type"] == "square" else 3)])
                obs, r, done, info = env.step(act)
                assert int(done) == z["done"][row], (name, e, row)
                assert float_close(float(r), float(z["reward"][row])), (name, e, row, r, z["reward"][row])
                exact += float(r) == float(z["reward"][row])
                total += 1
                if np.isnan(z["wirelength"][row]):
                    assert not info, (name, e, row, info)
                else:
                    assert float_close(float(info["wirelength"]), float(z["wirelength"][row]))
                    assert float_close(float(info["num_intersections"]), float(z["num_intersections"][row]))
            obs = flat_obs(obs) if flat else obs
            assert set(obs) == set(keys), (name, set(obs) ^ set(keys))
            for k in keys:
                want, got = z["obs_" + k][row], np.asarray(obs[k])
                if k == "component_grid" and got.shape[0] < C:
                    got = np.concatenate([got, np.zeros((C - got.shape[0],) + got.shape[1:], got.dtype)], 0)
                assert got.dtype == want.dtype, (name, k, got.dtype, want.dtype)
                assert np.array_equal(got, want), (name, e, row, k)
            if check_spaces:
                for k, sp in env.observation_space.spaces.items():
                    if k == "all_pins_cat_feature" and cfg["type"] == "rectangle_spatial_pin":
                        # declared int32, emitted float64 -- in the reference too (spatial:537-545)
                        assert not sp.contains(obs[k]) and sp.contains(np.asarray(obs[k]).astype(np.int32))
                    elif k == "component_grid":
                        assert sp.shape[1:] == np.asarray(obs[k]).shape[1:]
                    else:
                        assert sp.contains(np.asarray(obs[k])), (name, k)
    assert exact == total, (name, "rewards not bit-identical", exact, total)
    return env


@pytest.mark.parametrize("name", NAMES)
def test_oracle_factory_reproduces_reference_create_env(name):
    class FlatOracle:  # env_wrappers.py on the oracle: flat action in, tuple to the env
        def __init__(self, cfg):
            self.env = oracle.make_oracle_env(cfg)

        def reset(self, **k):
            return self.env.reset(**k)

        def step(self, a):
            if isinstance(a, int):
                a = oracle.flat_to_tuple(a, self.env.action_mask.shape)
            return self.env.step(tuple(a))

    replay(name, FlatOracle,
           lambda o: {k: oracle.flatten_mask(v) if k == "action_mask" else v for k, v in o.items()})


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_create_env_reproduces_reference_create_env(name):
    """The plugin API itself: pcbgpu.create_env(env_config) against the reference's create_env."""
    import pcbgpu

    env = replay(name, pcbgpu.create_env, lambda o: o, check_spaces=True)
    cfg = pcbgpu.SHIPPED_CONFIGS[name]
    z = load(name)[0]
    H, W = cfg["height"], cfg["width"]
    O = 1 if cfg["type"] == "square" else (2 if cfg["type"].startswith("rectangle_model") or cfg["type"] in (
        "rectangle", "rectangle_factorized") else 4)
    if bool(z["flat"]):
        assert env.action_space.n == O * H * W
        assert env.observation_space["action_mask"].shape == (O * H * W,)
    else:
        assert [s.n for s in env.action_space.spaces] == [O, H, W]
    env.close()
    assert REL_TOL == 1e-6
