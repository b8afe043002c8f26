"""Known-answer tests taken from the reference's own unit tests (SURVEY.md section 4
table: tests/square_environment, tests/rectangular_environment, tests/pin_environment).

Every case runs against the oracle (CPU, always) and against the CUDA environment
through GymCompatEnv (marked gpu).  The expected values are the reference's published
known answers, re-expressed as (instance, action sequence) -> observation / reward.
"""
import math

import numpy as np
import pytest

import oracle
from oracle import Instance

BACKENDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]


def make_env(backend, cfg):
    if backend == "oracle":
        c = dict(cfg)
        if c["type"] not in ("square", "rectangle"):
            c["forward_weight_num_intersections"] = True
        return oracle.make_oracle_env(c)
    import pcbgpu

    return pcbgpu.GymCompatEnv(cfg)


def inst(h, w, pins=(), n_nets=0):
    """pins: (comp, rx, ry, net, pin_id) in net-major order."""
    p = np.array(pins, np.int64).reshape(-1, 5)
    return Instance(np.array(h, np.int64), np.array(w, np.int64), p[:, 0], p[:, 1], p[:, 2],
                    p[:, 3], p[:, 4], n_nets)


SQ5 = dict(type="square", height=5, width=5, component_n=2)
SQ4 = dict(type="square", height=4, width=4, component_n=2)
SQ11 = dict(type="square", height=11, width=10, component_n=3)
RECT6 = dict(type="rectangle", height=6, width=6, min_component_w=2, max_component_w=4,
             min_component_h=2, max_component_h=4, max_num_components=4, min_num_components=1)
PIN = dict(net_distribution=1, pin_spread=1, min_component_w=2, max_component_w=4,
           min_component_h=2, max_component_h=4, min_num_pins_per_net=2, reward_type="both",
           reward_beam_width=2, weight_wirelength=0.5, weight_num_intersections=0.5)
PIN6 = dict(PIN, type="rectangle_pin", height=6, width=6, max_num_components=4,
            min_num_components=2, min_num_nets=4, max_num_nets=4, max_num_pins_per_net=2)
PIN30 = dict(PIN, type="rectangle_pin", height=30, width=30, max_component_w=5, max_component_h=5,
             max_num_components=6, min_num_components=1, min_num_nets=2, max_num_nets=4,
             max_num_pins_per_net=5)


# ---- square (tests/square_environment/test_env.py:80-258) -----------------------
@pytest.mark.parametrize("backend", BACKENDS)
def test_square_reset_and_masks(backend):
    env = make_env(backend, SQ5)
    obs = env.reset()
    want = np.ones((5, 5)); want[4, :] = 0; want[:, 4] = 0
    assert np.array_equal(obs["action_mask"], want) and obs["action_mask"].dtype == np.float32
    assert not obs["grid"].any()
    for action, zero_rows, zero_cols in [((0, 0), (0, 2), (0, 2)), ((1, 0), (0, 3), (0, 2)),
                                         ((0, 2), (0, 2), (1, 4))]:
        env.reset()
        obs, reward, done, info = env.step(action)
        exp = want.copy()
        exp[zero_rows[0]:zero_rows[1], zero_cols[0]:zero_cols[1]] = 0
        assert np.array_equal(obs["action_mask"], exp), action
        assert reward == 1.0 and done is False and info == {}
    env = make_env(backend, SQ11)
    obs = env.reset()
    assert obs["action_mask"].sum() == 9 * 8 and not obs["action_mask"][9:, :].any()
    obs, *_ = env.step((3, 3))  # zeroes rows 1..5 x cols 1..5 (test_env.py:172-221)
    exp = np.ones((11, 10)); exp[9:, :] = 0; exp[:, 8:] = 0; exp[1:6, 1:6] = 0
    assert np.array_equal(obs["action_mask"], exp)


@pytest.mark.parametrize("backend", BACKENDS)
def test_square_fill_and_invalid(backend):
    env = make_env(backend, SQ4)
    assert env.reset()["action_mask"].sum() == 9
    total = 0.0
    for k, a in enumerate([(0, 0), (0, 2), (2, 0), (2, 2)]):
        obs, r, done, _ = env.step(a)
        total += r
        assert done == (k == 3)
    assert obs["grid"].all() and total == 4.0
    env.reset()
    env.step((0, 0))
    before = env.grid.copy()
    obs, r, done, info = env.step((1, 1))  # overlaps -> invalid: nothing changes, episode ends
    assert r == 0.0 and done is True and info == {} and np.array_equal(obs["grid"], before)
    obs, r, done, _ = env.step((3, 3))  # out of the mask (component would leave the grid)
    assert r == 0.0 and done is True


# ---- rectangular (tests/rectangular_environment/test_env.py:108-311) -------------
@pytest.mark.parametrize("backend", BACKENDS)
def test_rect_mask_validate_step(backend):
    env = make_env(backend, RECT6)
    obs = env.reset(instance=inst([1, 3], [2, 2]))
    assert env.validate_action(0, 0, 0) and env.validate_action(1, 2, 3)
    assert not env.validate_action(0, 4, 5) and not env.validate_action(1, 5, 4)
    assert obs["action_mask"].shape == (2, 6, 6)
    assert obs["action_mask"][0].sum() == 6 * 5 and obs["action_mask"][1].sum() == 5 * 6
    obs, r, done, info = env.step((0, 0, 0))
    assert r == 1.0 and done is False and info == {}
    assert obs["placement_mask"].tolist() == [1, 0, 0, 0]
    assert obs["component_mask"].tolist() == [1, 1, 0, 0]
    assert obs["grid"][0, :2].all() and obs["grid"].sum() == 2
    assert obs["all_components_feature"][0].tolist() == [1, 2, 0, 0, 2 / 36]
    assert obs["all_components_feature"][1].tolist() == [3, 2, -1, -1, 6 / 36]


@pytest.mark.parametrize("backend", BACKENDS)
def test_rect_done(backend):
    env = make_env(backend, RECT6)
    env.reset(instance=inst([2, 3], [2, 3]))
    _, _, done, _ = env.step((0, 0, 0))
    assert not done
    obs, r, done, _ = env.step((0, 2, 3))
    assert done and r == 1.0 and not obs["action_mask"].any()  # all placed
    env.reset(instance=inst([4, 4], [4, 4]))
    obs, r, done, _ = env.step((0, 1, 1))
    assert done and r == 1.0 and not obs["action_mask"].any()  # next component blocked


# ---- pin env (tests/pin_environment/*) -------------------------------------------
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("o,pos,want", [
    (0, (2, 0), [(2, 0), (2, 2)]), (1, (0, 0), [(0, 3), (2, 3)]),
    (2, (0, 0), [(3, 2), (3, 0)]), (3, (0, 0), [(2, 0), (0, 0)])])
def test_pin_rotation(backend, o, pos, want):
    """test_component.py:1-34: 4x3 component, pins at (0,0) and (0,2)."""
    cfg = dict(PIN30, height=10, width=10, max_component_h=4, max_component_w=4)
    env = make_env(backend, cfg)
    env.reset(instance=inst([4, 2], [3, 2], [(0, 0, 0, 0, 0), (0, 0, 2, 0, 1)], 1))
    obs, r, done, _ = env.step((o,) + pos)
    rows = obs["all_pins_num_feature"][0]
    assert [tuple(rows[0, 2:]), tuple(rows[1, 2:])] == want
    assert r == 0.0 and not done and obs["placement_mask"].tolist()[:3] == [2, 3, 0]


@pytest.mark.parametrize("backend", BACKENDS)
def test_pin_feature_rows(backend):
    """test_env.py:782-828 on the 30x30 grid."""
    pins = [(0, 0, 0, 0, 0), (2, 2, 2, 0, 1), (0, 0, 2, 1, 1), (1, 3, 1, 1, 0),
            (1, 1, 0, 2, 1), (2, 2, 0, 2, 0)]
    env = make_env(backend, PIN30)
    obs = env.reset(instance=inst([1, 4, 5], [3, 2, 5], pins, 3))
    num, cat = obs["all_pins_num_feature"], obs["all_pins_cat_feature"]
    assert num.shape == (6, 25, 4) and cat.shape == (6, 25, 1)
    assert not num[3:].any() and not cat[3:].any()
    obs, *_ = env.step((0, 28, 26))
    num, cat = obs["all_pins_num_feature"], obs["all_pins_cat_feature"]
    assert num[0, 0].tolist() == [0, 0, 28, 26] and cat[0, 0, 0] == 0
    assert num[0, 1].tolist() == [0, 2, 28, 28] and cat[0, 1, 0] == 1
    assert not num[0, 2:].any() and not cat[0, 2:].any()
    obs, r, done, info = env.step((1, 0, 0))
    num, cat = obs["all_pins_num_feature"], obs["all_pins_cat_feature"]
    assert num[1, 0].tolist() == [1, 0, 1, 0] and cat[1, 0, 0] == 1
    assert num[1, 1].tolist() == [0, 2, 0, 2] and cat[1, 1, 0] == 2
    assert num[2, 0].tolist() == [2, 0, -1, -1] and cat[2, 0, 0] == 2
    assert num[2, 1].tolist() == [2, 2, -1, -1] and cat[2, 1, 0] == 0
    assert not num[1, 2:].any() and not num[2, 2:].any()
    assert r == 0.0 and done is False and info == {}
    assert obs["placement_mask"].tolist() == [2, 2, 3, 0, 0, 0]  # test_env.py:626-658 pattern


REWARD_PINS = {0: [(0, 2), (5, 3), (8, 1)], 1: [(2, 0), (3, 4), (4, 1), (7, 5)]}


def reward_case(backend, reward_type):
    """test_env.py:199-391: the pin positions of the reference's reward tests, realised
    as seven 1x1 components (one pin each) placed on those cells; the normalisers
    (H+W = 20, min(31.5, 8) = 8) are those of env_10_10_reward."""
    cfg = dict(PIN, type="rectangle_pin", height=10, width=10, max_num_components=7,
               min_num_components=2, min_num_nets=2, max_num_nets=2, max_num_pins_per_net=4,
               min_num_pins_per_net=4, reward_type=reward_type)
    cells = [c for net in (0, 1) for c in REWARD_PINS[net]]
    pins = [(i, 0, 0, 0 if i < 3 else 1, 0) for i in range(7)]
    env = make_env(backend, cfg)
    env.reset(instance=inst([1] * 7, [1] * 7, pins, 2))
    for i, (x, y) in enumerate(cells):
        obs, r, done, info = env.step((0, x, y))
        assert done == (i == 6)
    return r, info


@pytest.mark.parametrize("backend", BACKENDS)
def test_reward_known_answers(backend):
    s = np.sqrt
    beam_wl = s(26) + s(13) + s(17) + s(10) + s(5)
    cen_wl = 13 / 3 + s(13) / 3 + s(130) / 3 + s(41) / 2 + 3 / 2 + s(61) / 2 + s(13) / 2
    for kind, wl, crossings in [("beam", beam_wl, 1), ("centroid", cen_wl, 2), ("both", beam_wl, 1)]:
        r, info = reward_case(backend, kind)
        assert np.isclose(r, -0.5 * (wl / 20 + crossings / 8)), kind
        assert np.isclose(info["wirelength"], wl / 20) and info["num_intersections"] == crossings / 8


@pytest.mark.parametrize("backend", BACKENDS)
def test_reward_not_all_placed_and_bounds(backend):
    cfg = dict(PIN, type="rectangle_pin", height=10, width=10, max_num_components=5,
               min_num_components=2, min_num_nets=4, max_num_nets=4, max_num_pins_per_net=2)
    env = make_env(backend, cfg)
    env.reset(instance=inst([2, 2], [2, 2], [(0, 0, 0, 0, 0), (1, 0, 0, 0, 0)], 1))
    obs, r, done, info = env.step((0, 9, 9))  # invalid -> the not-all-placed penalty
    assert done and np.isclose(r, -0.5 * 2 * math.sqrt(2) - 0.5 * 24 / 8)
    assert info["num_intersections"] == 24 and np.isclose(info["wirelength"], 0.5 * 8 * math.sqrt(200))
    env6 = make_env(backend, PIN6)
    assert np.isclose(env6.max_wirelength, 0.5 * 8 * math.sqrt(72))  # test_env.py:185-191
    env48 = make_env(backend, dict(PIN6, max_num_nets=3, min_num_nets=3, max_num_pins_per_net=4))
    assert env48.max_num_intersections == 48  # test_env.py:88-93


# ---- function-level known answers (oracle restatement only) ----------------------
def test_route_functions():
    from oracle import reward as rw

    assert not rw.segments_cross(((1, 1), (3, 3)), ((1, 3), (1, 5)))
    assert rw.segments_cross(((1, 1), (3, 3)), ((1, 3), (2, 1)))
    four = [[((1, 1), (3, 3))], [((2, 1), (0, 3))], [((2, 3), (0, 1))], [((3, 2), (1, 3))]]
    assert rw.count_crossings(four) == 4
    assert rw.count_crossings([[((4, 4), (3, 5))], [((3, 4), (4, 5))]]) == 1
    assert rw.centroid([(0, 0), (0, 1), (1, 0), (1, 1)]) == (0.5, 0.5)
    assert rw.route_centroid([[(0, 0), (0, 1)], [(2, 2), (3, 3), (4, 4)]]) == [
        [((0, 0), (0, 1))],
        [((2, 2), (3.0, 3.0)), ((3, 3), (3.0, 3.0)), ((4, 4), (3.0, 3.0))]]
    assert rw.outlier([(0, 0), (0, 1), (1, 0), (3, 3)]) == (3, 3)
    assert rw.beam_path((0, 0), [(2, 2), (0, 1), (1, 0), (1, 1)], 4) == [
        (0, 0), (0, 1), (1, 0), (1, 1), (2, 2)]
    assert rw.beam_path((0, 0), {(2, 2), (0, 1), (1, 0), (1, 1)}, 2) == [
        (0, 0), (0, 1), (1, 1), (1, 0), (2, 2)]
    assert rw.route_beam([[(3, 3), (3, 4)], [(0, 0), (0, 1), (1, 0), (1, 1), (2, 2)]], 2) == [
        [((3, 3), (3, 4))],
        [((2, 2), (1, 1)), ((1, 1), (0, 1)), ((0, 1), (0, 0)), ((0, 0), (1, 0))]]
    assert np.isclose(rw.wirelength([[((3, 1), (2, 2))], [((1, 2), (2, 2))], [((3, 3), (2, 2))]]),
                      1 + 2 * np.sqrt(2))
    assert rw.distance((0, 0), (1, 1)) == math.sqrt(2)


def test_flat_wrappers():
    from oracle import flat

    assert flat.flat_to_tuple(37, (5, 10)) == [3, 7]
    assert flat.flat_to_tuple(2 * 100 + 4 * 10 + 9, (4, 10, 10)) == [2, 4, 9]
    assert flat.tuple_to_flat((2, 4, 9), (4, 10, 10)) == 249
    m = np.arange(24).reshape(2, 3, 4)
    assert flat.flatten_mask(m).tolist() == list(range(24))
