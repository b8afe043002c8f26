"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/pcb_env.h declares, fails loudly without a GPU, and the host logic (config
validation, constants, instance packing, wrappers, set-order model) is right."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

import __graft_entry__ as ge
import oracle
from ref_harness import CONFIGS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pcbgpu():
    return ge.build()


def test_library_exports_header_symbols(pcbgpu):
    from pcbgpu import _lib

    header = open(os.path.join(ROOT, "include", "pcb_env.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(pcb_\w+)\s*\(", header, re.M))
    assert len(declared) >= 18
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"libpcbgpu.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.pcb_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pcbgpu):
    from pcbgpu import _lib

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pcbgpu.PcbVecEnv(CONFIGS["C"], 4)
    lib = _lib.load()
    spec = pcbgpu.make_spec(CONFIGS["C"])
    h = ctypes.c_void_p()
    rc = lib.pcb_create(ctypes.byref(spec.c), 0, 4, ctypes.byref(h))
    assert rc == -2 and b"cuda" in lib.pcb_last_error().lower()  # PCB_ERR_CUDA, nothing created
    assert not h.value


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rl-environment-for-component-placement_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "/root/reference" not in text, f


def test_config_constants_match_oracle(pcbgpu):
    for name in ("C", "C_both", "D", "E", "S_pin", "S_pin_beam3", "S_spa", "REF_pin"):
        cfg = CONFIGS[name]
        spec = pcbgpu.make_spec(cfg)
        o = oracle.make_oracle_env(dict(cfg, forward_weight_num_intersections=True))
        assert spec.c.penalty_reward == o.penalty(), name
        assert spec.max_wirelength == o.max_wirelength
        assert spec.max_num_intersections == o.max_num_intersections
        wl_norm, ni_norm = o.normalizers()
        assert spec.c.wirelength_norm == wl_norm and spec.c.intersections_norm == ni_norm
        assert spec.mask_shape == o.mask_shape


def test_constructor_validation_matches_reference_messages(pcbgpu):
    bad = [
        (dict(CONFIGS["A"], component_n=9), "Component size must be less than or equal"),
        (dict(CONFIGS["B"], max_component_w=17), "Component size must be less than or equal"),
        (dict(CONFIGS["B"], min_component_h=0), "Component size must be greater than 0"),
        (dict(CONFIGS["B"], max_num_components=0), "Number of components must be greater than 0"),
        (dict(CONFIGS["C"], min_num_pins_per_net=7), "must not be greater than max num pins"),
        (dict(CONFIGS["C"], min_num_pins_per_net=1), "min num pins per net"),
        (dict(CONFIGS["C"], min_num_nets=50, max_num_nets=50), "total minimum area"),
        (dict(CONFIGS["C"], reward_beam_width=0), "Beam width must be a positive integer"),
        (dict(CONFIGS["C"], reward_type="astar"), "Reward type must be either"),
        (dict(CONFIGS["D"], reward_beam_width=1), "beam width must be an integer greater than 2"),
        (dict(CONFIGS["D"], weight_wirelength=1), "weight of wirelength must be a float"),
    ]
    for cfg, msg in bad:
        with pytest.raises(ValueError, match=msg):
            pcbgpu.make_spec(cfg)
        kw = dict(cfg)
        if kw["type"] not in ("square", "rectangle"):
            kw["forward_weight_num_intersections"] = True
        with pytest.raises(ValueError, match=msg):
            oracle.make_oracle_env(kw)


def test_instance_bank_packing(pcbgpu):
    cfg = CONFIGS["S_pin"]
    spec = pcbgpu.make_spec(cfg)
    np.random.seed(3)
    random.seed(3)
    insts = [oracle.generate_pin(cfg) for _ in range(5)]
    bank = pcbgpu.InstanceBank.from_list(spec, insts)
    s = bank.as_struct()
    assert s.count == 5 and bank.pin_comp.shape == (5, spec.max_pins)
    for i, it in enumerate(insts):
        assert bank.n_comp[i] == it.n_comp and bank.n_pins[i] == it.n_pins
        assert np.array_equal(bank.pin_id[i, : it.n_pins], it.pin_id)
        assert np.all(np.diff(bank.pin_net[i, : it.n_pins].astype(int)) >= 0)  # net-major
    with pytest.raises(ValueError):
        bank.set(0, dict(insts[0].to_dict(), comp_h=np.ones(99), comp_w=np.ones(99)))


def test_wrapper_classes_follow_reference_decoding(pcbgpu):
    class Fake:
        action_sizes = [4, 6, 7]

        def __init__(self):
            self.action_mask = np.arange(4 * 6 * 7).reshape(4, 6, 7) % 2

        def reset(self):
            return {"action_mask": self.action_mask, "grid": np.zeros((6, 7))}

        def step(self, a):
            self.last = a
            return self.reset(), 0.0, False, {}

        def validate_action(self, o, x, y):
            return bool(self.action_mask[o, x, y])

    env = pcbgpu.FlatteningActionWrapperRect(pcbgpu.FlatteningActionMaskObservationWrapperRect(Fake()))
    obs = env.reset()
    assert obs["action_mask"].shape == (168,) and obs["grid"].shape == (6, 7)
    for flat in (0, 41, 42, 167, 100):
        env.step(flat)
        o, rem = divmod(flat, 42)
        assert env.env.env.last == [o, rem // 7, rem % 7]
        assert env.validate_action(flat) == bool(obs["action_mask"][flat])
    sq = pcbgpu.FlatteningActionWrapperSquare(Fake())
    sq.factor_sizes = [6, 7]
    assert sq.action(20) == [2, 6]


def test_pyset_model_matches_cpython_sets():
    from oracle.pyset_model import M64, build, difference_order, tuple_hash

    rnd = random.Random(11)
    assert tuple_hash(3, 5) == hash((3, 5)) & M64
    for _ in range(3000):
        g = rnd.choice([10, 32, 64])
        m = rnd.randint(1, 18)
        pts = rnd.sample([(x, y) for x in range(g) for y in range(g)], m)
        t = build(pts)
        real = set(pts)
        assert [pts[k] for k in t.order()] == list(real)
        L = rnd.randint(0, m - 1) if m > 1 else 0
        seen_idx = set(rnd.sample(range(m), L))
        got = [pts[k] for k in difference_order(t, seen_idx)]
        assert got == list(real - {pts[k] for k in seen_idx})


def test_generators_follow_the_reference_rules():
    cfg = CONFIGS["C"]
    np.random.seed(0)
    random.seed(0)
    for _ in range(200):
        it = oracle.generate_pin(cfg)
        assert it.n_comp == 10 and it.n_nets == 5 and 10 <= it.n_pins <= 30
        per_net = np.bincount(it.pin_net, minlength=5)
        assert per_net.min() >= 2 and per_net.max() <= 6
        cells = set()
        for c, x, y in zip(it.pin_comp, it.pin_rx, it.pin_ry):
            assert 0 <= x < it.comp_h[c] and 0 <= y < it.comp_w[c]
            assert (c, x, y) not in cells
            cells.add((c, x, y))


def test_library_generator_matches_reference_distribution(pcbgpu):
    """pcb_generate_host follows the reference's generate_instances construction with a
    different RNG: the instances must satisfy the same invariants and have the same
    distributions as the oracle generator (two-sample KS on the aggregate statistics)."""
    from scipy import stats

    for name in ("C", "S_pin", "S_spa"):
        cfg = CONFIGS[name]
        spatial = cfg["type"] == "rectangle_spatial_pin"
        spec = pcbgpu.make_spec(cfg)
        n = 1500
        bank = pcbgpu.generate_bank(spec, n, seed=123)
        np.random.seed(5)
        random.seed(5)
        ref = [oracle.generate_pin(cfg, spatial) for _ in range(n)]
        lib_stats = {"n_comp": [], "n_pins": [], "n_nets": [], "area": [], "max_net": [], "max_comp": [],
                     "comps_with_pins": []}
        ref_stats = {k: [] for k in lib_stats}
        for i in range(n):
            nc, npn, nn = int(bank.n_comp[i]), int(bank.n_pins[i]), int(bank.n_nets[i])
            h, w = bank.comp_h[i, :nc].astype(int), bank.comp_w[i, :nc].astype(int)
            pc, pn = bank.pin_comp[i, :npn].astype(int), bank.pin_net[i, :npn].astype(int)
            rx, ry = bank.pin_rx[i, :npn].astype(int), bank.pin_ry[i, :npn].astype(int)
            assert cfg["min_num_components"] <= nc <= cfg["max_num_components"]
            assert np.all(np.diff(pn) >= 0) and (npn == 0 or pn.max() == nn - 1)
            per_net = np.bincount(pn, minlength=nn)
            assert per_net.min() >= cfg["min_num_pins_per_net"]
            assert per_net.max() <= cfg["max_num_pins_per_net"]
            assert np.all(rx < h[pc]) and np.all(ry < w[pc])
            assert len({(c, a, b) for c, a, b in zip(pc, rx, ry)}) == npn  # one pin per cell
            per_comp = np.bincount(pc, minlength=nc)
            assert np.all(per_comp <= h * w)
            if spatial:
                assert sorted(bank.pin_id[i, :npn].tolist()) == list(range(npn))
            for dst, (a_nc, a_np, a_nn, a_area, a_pn, a_pc) in (
                    (lib_stats, (nc, npn, nn, int((h * w).sum()), per_net, per_comp)),):
                dst["n_comp"].append(a_nc); dst["n_pins"].append(a_np); dst["n_nets"].append(a_nn)
                dst["area"].append(a_area); dst["max_net"].append(int(a_pn.max()))
                dst["max_comp"].append(int(a_pc.max())); dst["comps_with_pins"].append(int((a_pc > 0).sum()))
            it = ref[i]
            pn2, pc2 = np.bincount(it.pin_net, minlength=it.n_nets), np.bincount(it.pin_comp, minlength=it.n_comp)
            ref_stats["n_comp"].append(it.n_comp); ref_stats["n_pins"].append(it.n_pins)
            ref_stats["n_nets"].append(it.n_nets); ref_stats["area"].append(int((it.comp_h * it.comp_w).sum()))
            ref_stats["max_net"].append(int(pn2.max())); ref_stats["max_comp"].append(int(pc2.max()))
            ref_stats["comps_with_pins"].append(int((pc2 > 0).sum()))
        for k in lib_stats:
            p = stats.ks_2samp(lib_stats[k], ref_stats[k]).pvalue
            assert p > 1e-3, (name, k, p, np.mean(lib_stats[k]), np.mean(ref_stats[k]))


def test_host_buffers_are_validated_before_their_address_is_taken():
    """pcb_step_host gets raw pointers: wrong dtype / layout must raise, not be misread"""
    import numpy as np
    import pytest
    import torch

    from pcbgpu.vec_env import _host_buffer

    assert _host_buffer(np.zeros(4, np.int32), "int32", "actions")
    assert _host_buffer(torch.zeros(4, dtype=torch.float64), "float64", "reward")
    with pytest.raises(ValueError):
        _host_buffer(np.zeros(4, np.int64), "int32", "actions")
    with pytest.raises(ValueError):
        _host_buffer(np.zeros((4, 6), np.int32)[:, ::2], "int32", "actions")
    with pytest.raises(ValueError):
        _host_buffer(np.zeros(4, np.float32), "float64", "reward")
    with pytest.raises(ValueError):
        _host_buffer(torch.zeros(4, dtype=torch.int32), "uint8", "done")
    ro = np.zeros(4, np.uint8)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        _host_buffer(ro, "uint8", "done")
