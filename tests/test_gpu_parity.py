"""Parity of the CUDA environments (through the C ABI) with the oracle: the
parity tests proper.  Run on the B200 box (`pytest -m gpu`)."""
import pytest

from parity import run_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["A", "A2", "S_sq"])
def test_square(name):
    run_parity(name, range(16), extra_steps=2)
    run_parity(name, range(16, 32), p_invalid=0.3)


@pytest.mark.parametrize("name", ["B", "S_rect", "S_rect_full", "S_rect_many"])
def test_rect(name):
    run_parity(name, range(64), extra_steps=2)
    run_parity(name, range(64, 96), p_invalid=0.2)


@pytest.mark.parametrize("name", ["C", "S_pin_tight", "REF_pin"])
def test_pin_centroid(name):
    st = {}
    run_parity(name, range(64), extra_steps=2, stats=st)
    run_parity(name, range(64, 96), p_invalid=0.1, stats=st)
    assert st["reward_inexact"] == 0, st  # float64 rewards bit-identical


@pytest.mark.parametrize("name", ["C_both", "C_beam", "S_pin", "S_pin_beam3"])
def test_pin_beam(name):
    st = {}
    run_parity(name, range(64), extra_steps=2, stats=st)
    run_parity(name, range(64, 96), p_invalid=0.1, stats=st)
    assert st["reward_inexact"] == 0, st


@pytest.mark.parametrize("name", ["D", "D_both", "S_spa", "S_spa_tight"])
def test_spatial(name):
    st = {}
    run_parity(name, range(48), extra_steps=2, stats=st)
    run_parity(name, range(48, 64), p_invalid=0.1, stats=st)
    assert st["reward_inexact"] == 0, st


def test_large_grid_E():
    run_parity("E", range(8))


@pytest.mark.parametrize("name", ["A", "B", "C", "D", "S_pin_tight"])
def test_autoreset(name):
    run_parity(name, range(32), autoreset=True, episodes=3)
    run_parity(name, range(32, 48), autoreset=True, episodes=3, p_invalid=0.1)


@pytest.mark.parametrize("name", ["S_wide_pin", "S_wide_spa", "S_wide_rect", "S_wide_sq"])
def test_wide_grids_runtime_shaped(name):
    """64-bit bitboard rows with the runtime-shaped kernels (grids up to 64x64, non-square)."""
    st = {}
    n = 8 if name == "S_wide_sq" else 24
    run_parity(name, range(n), extra_steps=2, stats=st)
    run_parity(name, range(n, n + 12), p_invalid=0.1, stats=st)
    run_parity(name, range(8), autoreset=True, episodes=2, stats=st)
    assert st["reward_inexact"] == 0, st


@pytest.mark.parametrize("name", ["REF_pin", "REF_spatial", "REF_rect", "REF_square"])
def test_reference_shipped_configs(name):
    """agent/config/*.json (10x10): compiled as FixedShape without the fast board path."""
    st = {}
    run_parity(name, range(32), extra_steps=2, stats=st)
    run_parity(name, range(32, 48), p_invalid=0.15, stats=st)
    run_parity(name, range(16), autoreset=True, episodes=3, stats=st)
    run_parity(name, range(8), force_generic=True, stats=st)
    assert st["reward_inexact"] == 0, st


@pytest.mark.parametrize("name", ["C", "REF_pin", "S_pin_tight", "S_pin_cen", "FZ_rect_pin_0_cen",
                                  "FZ_rect_pin_3_cen", "FZ_rect_pin_4_cen", "FZ_rect_pin_6_cen",
                                  "D", "REF_spatial", "S_spa_tight", "S_spa_cen", "C_both", "C_beam", "S_pin",
                                  "S_pin_beam3", "S_spa", "FZ_rect_pin_0", "FZ_rect_pin_3", "FZ_rect_pin_6"])
def test_pin_rollout_kernel_against_oracle(name):
    """pcb_rollout_random of the pin env in its multi-step kernel (pcb_pin_small.cuh: 8 / 16 / 32
    lanes per env, compile-time and runtime shapes): the logged actions replayed through the
    oracle, every observation of every step, rewards, dones and the episode count."""
    from parity import run_rollout_parity

    assert run_rollout_parity(name, N=27, T=45, episodes=8) == 27 * 45
