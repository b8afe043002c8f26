"""pcbgpu.spaces against the spaces the unmodified reference constructs (CPU, build container)
and against gym-0.22 semantics of contains / sample the reference's tests rely on
(tests/rectangular_environment/test_env.py:9-29, tests/pin_environment/test_env.py:594-603)."""
import numpy as np
import pytest

from ref_harness import CONFIGS, load_reference, make_reference_env


def spec_of(cfg):
    import pcbgpu

    return pcbgpu.make_spec(cfg)


NAMES = ["A", "B", "C", "D", "E", "S_sq", "S_rect", "S_pin", "S_spa", "S_wide_pin", "S_wide_spa",
         "REF_pin", "REF_spatial", "REF_rect", "REF_square"]


@pytest.mark.parametrize("name", NAMES)
def test_spaces_equal_the_reference(name):
    from pcbgpu import spaces

    mods = load_reference()
    if mods is None:
        pytest.skip("reference not present")
    cfg = CONFIGS[name]
    ref = make_reference_env(mods, cfg)
    spec = spec_of(cfg)
    a, o = spaces.action_space(spec), spaces.observation_space(spec)
    assert [s.n for s in a.spaces] == [s.n for s in ref.action_space.spaces]
    assert set(o.spaces) == set(ref.observation_space.spaces)
    for k, want in ref.observation_space.spaces.items():
        got = o.spaces[k]
        assert tuple(got.shape) == tuple(want.shape), (name, k)
        assert np.dtype(got.dtype) == np.dtype(want.dtype), (name, k)
        assert float(np.min(got.low)) == float(np.min(want.low)), (name, k)
        assert float(np.max(got.high)) == float(np.max(want.high)), (name, k)
    # the wrappers: flattened mask space and Discrete(prod) (env_wrappers.py:28-33, 76-78)
    wr = mods["wrappers"]
    if cfg["type"] == "square":
        w = wr.FlatteningActionWrapperSquare(wr.FlatteningActionMaskObservationWrapperSquare(ref))
    else:
        w = wr.FlatteningActionWrapperRect(wr.FlatteningActionMaskObservationWrapperRect(ref))
    assert spaces.flat_action_space(a).n == w.action_space.n
    fo = spaces.flattened_mask_observation_space(o)
    assert tuple(fo["action_mask"].shape) == tuple(w.observation_space["action_mask"].shape)
    assert np.dtype(fo["action_mask"].dtype) == np.dtype(w.observation_space["action_mask"].dtype)
    for k in o.spaces:
        if k != "action_mask":
            assert fo[k] is o[k]


def test_contains_and_sample_semantics():
    from pcbgpu import spaces

    spec = spec_of(CONFIGS["S_rect"])
    a = spaces.action_space(spec)
    H, W = spec.height, spec.width
    assert a.contains((0, 0, 0)) and a.contains((1, H - 1, W - 1))
    assert not a.contains((0, H, 0)) and not a.contains((2, 0, 0)) and not a.contains((0, 0, W))
    assert not a.contains((0, 0)) and not a.contains((0, -1, 0))
    for _ in range(50):
        assert a.contains(a.sample())
    o = spaces.observation_space(spec)
    s = o.sample()
    assert o.contains(s)
    bad = dict(s, grid=s["grid"] + 2.0)
    assert not o.contains(bad)
    assert not o.contains({k: v for k, v in s.items() if k != "grid"})
    d = spaces.flat_action_space(a)
    assert d.n == 2 * H * W and d.contains(d.n - 1) and not d.contains(d.n) and not d.contains(-1)
    assert d.contains(np.int64(3)) and not d.contains(2.5)


def test_wrappers_publish_spaces_without_a_device():
    """wrappers only need the wrapped env's spaces: a stand-in env shows the flattening"""
    import pcbgpu
    from pcbgpu import spaces

    spec = spec_of(CONFIGS["REF_pin"])

    class Dummy:
        action_space = spaces.action_space(spec)
        observation_space = spaces.observation_space(spec)
        action_sizes = [4, 10, 10]

    w = pcbgpu.FlatteningActionWrapperRect(pcbgpu.FlatteningActionMaskObservationWrapperRect(Dummy()))
    assert w.action_space.n == 400
    assert w.observation_space["action_mask"].shape == (400,)
    assert w.observation_space["grid"].shape == (10, 10)
    assert w.action(399) == [3, 9, 9] and w.action(123) == [1, 2, 3]
