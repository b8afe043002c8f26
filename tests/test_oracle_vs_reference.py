"""Pins the oracle against the UNMODIFIED reference (build container only).

Same seeds -> same instances (the oracle's generators issue the reference's
RNG calls in the reference's order) -> same observations, rewards, dones and
infos along random action sequences, bit for bit (float64 equality).
"""
import numpy as np
import pytest

from ref_harness import (CONFIGS, extract_instance, load_reference, make_oracle,
                         make_reference_env, pick_action, seed_all)

MODS = load_reference()
pytestmark = [pytest.mark.ref,
              pytest.mark.skipif(MODS is None, reason="reference not present")]


def assert_obs_equal(a, b, ctx):
    assert set(a.keys()) == set(b.keys()), ctx
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, (ctx, k, x.shape, y.shape)
        assert x.dtype == y.dtype, (ctx, k, x.dtype, y.dtype)
        assert np.array_equal(x, y), (ctx, k)


def run_pair(name, seeds, p_invalid=0.0, extra_steps=1):
    cfg = CONFIGS[name]
    ref = make_reference_env(MODS, cfg)
    orc = make_oracle(cfg)
    n_steps = 0
    for s in seeds:
        seed_all(s)
        o_ref = ref.reset()
        seed_all(s)
        o_orc = orc.reset()
        assert_obs_equal(o_ref, o_orc, (name, s, "reset"))
        if cfg["type"] != "square":
            inst = extract_instance(ref, cfg["type"])
            for f in ("comp_h", "comp_w", "pin_comp", "pin_rx", "pin_ry", "pin_net", "pin_id"):
                assert np.array_equal(getattr(inst, f), getattr(orc.inst, f)), (name, s, f)
        rng = np.random.default_rng(s)
        done, after = False, 0
        while after < extra_steps:  # keep stepping past done to cover the frozen case
            a = pick_action(rng, ref.action_mask, p_invalid)
            r1 = ref.step(a)
            r2 = orc.step(a)
            assert_obs_equal(r1[0], r2[0], (name, s, a))
            assert r1[1] == r2[1] and type(r1[1]) is type(r2[1]) or r1[1] == r2[1], (name, s, a)
            assert r1[2] == r2[2], (name, s, a)
            assert set(r1[3].keys()) == set(r2[3].keys())
            for k in r1[3]:
                assert r1[3][k] == r2[3][k], (name, s, a, k)
            n_steps += 1
            done = done or r1[2]
            if done:
                after += 1
    return n_steps


@pytest.mark.parametrize("name", ["A", "A2", "S_sq"])
def test_square(name):
    assert run_pair(name, range(20)) > 0
    run_pair(name, range(20, 30), p_invalid=0.3)


@pytest.mark.parametrize("name", ["B", "S_rect", "S_rect_full", "S_rect_many"])
def test_rect(name):
    run_pair(name, range(40))
    run_pair(name, range(40, 60), p_invalid=0.2)


@pytest.mark.parametrize("name", ["C", "C_both", "C_beam", "S_pin", "S_pin_beam3",
                                  "S_pin_tight", "REF_pin"])
def test_pin(name):
    run_pair(name, range(40), extra_steps=2)
    run_pair(name, range(40, 50), p_invalid=0.1)


@pytest.mark.parametrize("name", ["D", "D_both", "S_spa", "S_spa_tight"])
def test_spatial(name):
    run_pair(name, range(30), extra_steps=2)
    run_pair(name, range(30, 40), p_invalid=0.1)


def test_large_E():
    run_pair("E", range(3))


@pytest.mark.parametrize("name", ["S_wide_pin", "S_wide_spa", "S_wide_rect"])
def test_wide_grids(name):
    run_pair(name, range(12), extra_steps=2)
    run_pair(name, range(12, 18), p_invalid=0.1)


def test_wide_square():
    run_pair("S_wide_sq", range(2))


@pytest.mark.parametrize("name", ["REF_spatial", "REF_rect", "REF_square"])
def test_reference_shipped_configs(name):
    run_pair(name, range(20), extra_steps=2)


def _fuzz_names():
    from fuzz_configs import FUZZ

    return sorted(FUZZ)


@pytest.mark.parametrize("name", _fuzz_names())
def test_random_configurations(name):
    """Configurations drawn at random (non-square, wide, narrow, tiny boards; every reward
    type; beam widths 2..4): the oracle must still equal the reference bit for bit."""
    assert run_pair(name, range(4), p_invalid=0.1, extra_steps=2) > 0
