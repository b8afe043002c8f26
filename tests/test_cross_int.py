"""The integer pre-test of the crossing count (csrc/pcb_cross.h, used by reward_kernel) against
the float64 formula of the reference (is_intersect, dummy_env_rectangular_pin.py:687-739, as
restated in oracle/reward.py): wherever the integer test decides, it must give the
reference's answer; what it leaves undecided goes to the float64 formula on the GPU.

CPU test: the header is compiled with g++ (it has no CUDA types)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reward as oracle_reward  # noqa: E402

CSRC = os.path.join(ROOT, "rl-environment-for-component-placement_b200", "csrc")


@pytest.fixture(scope="module")
def cross_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("cross") / "libhostcross.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", so,
                           os.path.join(ROOT, "tests", "host_cross.cpp")])
    lib = ctypes.CDLL(so)
    lib.pcb_test_cross_int.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def reference_formula(x1, y1, x2, y2, x3, y3, x4, y4):
    """is_intersect vectorised: the same float64 operations in the same order."""
    shared = (((x1 == x3) & (y1 == y3)) | ((x1 == x4) & (y1 == y4)) | ((x2 == x3) & (y2 == y3)) |
              ((x2 == x4) & (y2 == y4)))
    det = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    with np.errstate(all="ignore"):
        x = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / det
        y = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / det
        inside = ((np.minimum(x1, x2) <= x) & (x <= np.maximum(x1, x2)) &
                  (np.minimum(x3, x4) <= x) & (x <= np.maximum(x3, x4)) &
                  (np.minimum(y1, y2) <= y) & (y <= np.maximum(y1, y2)) &
                  (np.minimum(y3, y4) <= y) & (y <= np.maximum(y3, y4)))
    return shared | ((det != 0) & inside)


def random_segments(rng, n, grid, max_scale, near=None):
    """pin -> centroid segments (scale = pins of the net) and pin -> pin segments (scale 1); with
    `near` the far end stays within 3 cells and the second segment starts near the first, which
    makes shared end points, collinear and axis-parallel cases frequent."""
    c = rng.integers(1, max_scale + 1, n)
    px, py = rng.integers(0, grid, n), rng.integers(0, grid, n)
    if near is not None:
        px = np.clip(near[0] + rng.integers(-4, 5, n), 0, grid - 1)
        py = np.clip(near[1] + rng.integers(-4, 5, n), 0, grid - 1)
        sx = np.clip(px * c + rng.integers(-3 * c, 3 * c + 1), 0, (grid - 1) * c)
        sy = np.clip(py * c + rng.integers(-3 * c, 3 * c + 1), 0, (grid - 1) * c)
    else:
        sx = rng.integers(0, (grid - 1) * c + 1)
        sy = rng.integers(0, (grid - 1) * c + 1)
    return px, py, sx, sy, c


@pytest.mark.parametrize("grid,max_scale,clustered", [
    (32, 6, False), (32, 6, True), (64, 6, True), (64, 16, False), (64, 16, True), (10, 6, True),
    (8, 16, True), (64, 20, True)])
def test_integer_test_agrees_with_the_reference_formula(cross_lib, grid, max_scale, clustered):
    rng = np.random.default_rng(grid * 1000 + max_scale * 10 + clustered)
    n = 400_000
    p1x, p1y, s1x, s1y, c1 = random_segments(rng, n, grid, max_scale, (rng.integers(0, grid, n),
                                             rng.integers(0, grid, n)) if clustered else None)
    p2x, p2y, s2x, s2y, c2 = random_segments(rng, n, grid, max_scale, (p1x, p1y) if clustered else None)
    rec = np.stack([p1x * c1, p1y * c1, s1x, s1y, c1, p2x * c2, p2y * c2, s2x, s2y, c2], axis=1)
    rec = np.ascontiguousarray(rec, dtype=np.int32)
    out = np.empty(n, dtype=np.int32)
    cross_lib.pcb_test_cross_int(n, rec.ctypes.data, out.ctypes.data)
    f = np.float64
    want = reference_formula(p1x.astype(f), p1y.astype(f), s1x / c1.astype(f), s1y / c1.astype(f),
                             p2x.astype(f), p2y.astype(f), s2x / c2.astype(f), s2y / c2.astype(f))
    decided = out != 2
    assert set(np.unique(out)) <= {0, 1, 2}
    assert np.array_equal(out[decided] == 1, want[decided])
    # nets above the cap are never decided; below it the test decides the large majority
    capped = (c1 > 16) | (c2 > 16)
    assert not decided[capped].any()
    assert decided[~capped].mean() > (0.6 if grid <= 10 else 0.85)
    # a sample through the oracle's own scalar restatement (plain Python floats)
    for i in rng.choice(n, 3000, replace=False):
        s = ((int(p1x[i]), int(p1y[i])), (s1x[i] / f(c1[i]), s1y[i] / f(c1[i])))
        t = ((int(p2x[i]), int(p2y[i])), (s2x[i] / f(c2[i]), s2y[i] / f(c2[i])))
        got = oracle_reward.segments_cross(s, t)
        assert got == bool(want[i])
        if out[i] != 2:
            assert got == (out[i] == 1)


def test_integer_test_known_cases(cross_lib):
    def run(a, b):
        rec = np.array([list(a) + list(b)], dtype=np.int32)
        out = np.empty(1, dtype=np.int32)
        cross_lib.pcb_test_cross_int(1, rec.ctypes.data, out.ctypes.data)
        return int(out[0])
    # pin -> pin segments (scale 1) are always decided: proper crossing, T junction, shared end
    # point, collinear overlap without a shared end point (det == 0 -> no), parallel, apart
    assert run((0, 0, 4, 4, 1), (0, 4, 4, 0, 1)) == 1
    assert run((0, 0, 4, 0, 1), (2, 0, 2, 3, 1)) == 1
    assert run((0, 0, 4, 0, 1), (4, 0, 7, 5, 1)) == 1
    assert run((0, 0, 4, 0, 1), (2, 0, 6, 0, 1)) == 0
    assert run((0, 0, 4, 0, 1), (4, 0, 6, 0, 1)) == 1
    assert run((0, 0, 4, 0, 1), (0, 1, 4, 1, 1)) == 0
    assert run((0, 0, 1, 1, 1), (3, 0, 4, 7, 1)) == 0
    # centroid of 3 pins (thirds): an oblique proper crossing is decided, a crossing on an
    # axis-parallel segment and an end point on the other line are left to float64
    assert run((0, 0, 10, 11, 3), (0 * 1, 3, 5, 0, 1)) == 1
    assert run((0, 3, 12, 3, 3), (2, 0, 2, 3, 1)) == 2
    assert run((0, 0, 9, 9, 3), (3, 3, 5, 0, 1)) == 2
    assert run((0, 0, 10, 11, 3), (20, 0, 25, 1, 1)) == 0
    # more pins per net than the cap: never decided
    assert run((0, 0, 10, 11, 17), (0, 3, 5, 0, 1)) == 2
