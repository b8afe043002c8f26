"""GPU tests of the API surface around the step kernel: the device random policy,
the fused rollout, host-buffer stepping, flat actions, the generic (runtime-shaped)
kernels on the canonical shapes, the rollout collector and the episode statistics."""
import numpy as np
import pytest
import torch

from parity import draw_instance, run_parity
from ref_harness import CONFIGS

pytestmark = pytest.mark.gpu


def make(name, n, **kw):
    import pcbgpu

    return pcbgpu.PcbVecEnv(CONFIGS[name], n, **kw)


@pytest.mark.parametrize("name", ["B", "C", "D"])
def test_generic_kernels_on_canonical_shapes(name):
    run_parity(name, range(24), extra_steps=2, force_generic=True)
    run_parity(name, range(24, 40), autoreset=True, episodes=2, p_invalid=0.1, force_generic=True)


@pytest.mark.parametrize("name", ["A", "B", "C", "E", "S_pin", "S_rect"])
def test_sampler_only_returns_valid_actions(name):
    env = make(name, 256, autoreset=True, seed=5)
    obs = env.reset()
    idx = torch.arange(256, device="cuda")
    for t in range(12):
        a = env.sample_random_actions(seed=3).long()
        mask = obs["action_mask"]
        picked = mask[idx, a[:, 0], a[:, 1]] if a.shape[1] == 2 else mask[idx, a[:, 0], a[:, 1], a[:, 2]]
        assert bool((picked == 1).all()), (name, t)
        obs, r, d, _ = env.step(a.int())
    env.close()


def test_sampler_is_uniform_over_valid_actions():
    cfg = CONFIGS["B"]
    inst = draw_instance(cfg, 1)
    n = 32768
    env = make("B", n, autoreset=False, instances=[inst])  # every env holds the same instance
    obs = env.reset()
    a = env.sample_random_actions(seed=11).long()
    flat = (a[:, 0] * 16 + a[:, 1]) * 16 + a[:, 2]
    valid = obs["action_mask"][0].reshape(-1).bool()
    counts = torch.bincount(flat, minlength=2 * 256).double()
    assert counts[~valid].sum() == 0
    k = int(valid.sum())
    exp = n / k
    chi2 = float((((counts[valid] - exp) ** 2) / exp).sum())
    assert chi2 < k + 6 * np.sqrt(2 * k), (chi2, k)  # mean k-1, sd sqrt(2k)
    env.close()


@pytest.mark.parametrize("name", ["C", "C_both", "B", "A"])
def test_fused_rollout_equals_sample_then_step(name):
    a, b = make(name, 512, seed=2), make(name, 512, seed=2)
    a.reset(); b.reset()
    a.rollout_random(25, seed=9)
    tot = torch.zeros(512, dtype=torch.float64, device="cuda")
    for t in range(25):
        act = b.sample_random_actions(seed=9)
        _, r, d, _ = b.step(act)
        tot += r
    torch.cuda.synchronize()
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] > 0 and sa["sum_length"] == sb["sum_length"]
    for k in sa:  # float64 atomics: the accumulation order differs between runs
        assert np.isclose(sa[k], sb[k], rtol=1e-12, atol=0), k
    assert np.array_equal(a.occupancy(), b.occupancy())
    for k in a._obs():
        assert torch.equal(a._obs()[k], b._obs()[k]), k
    a.close(); b.close()


SMALL = ["A", "A2", "B", "S_sq", "S_rect", "S_rect_full", "S_rect_many", "REF_rect", "REF_square"]


@pytest.mark.parametrize("autoreset", [True, False])
@pytest.mark.parametrize("name", SMALL)
def test_fused_small_rollout_is_bit_identical_to_per_step_launches(name, autoreset):
    """rollout_small_kernel (all steps in one launch, 32/LPE envs per warp) against the
    per-step launches: actions, every ring slice, results, state records, statistics."""
    n, R = 777, 4  # 777: the last warp / CTA is partly empty
    a = make(name, n, seed=3, autoreset=autoreset, obs_ring=R)
    b = make(name, n, seed=3, autoreset=autoreset, obs_ring=R)
    assert a.fused_rollout(True) and not b.fused_rollout(False)
    a.reset(); b.reset()
    w = a.spec.action_width
    for chunk in (1, 6, 23):  # several calls: the ring wraps, the step counter carries over
        la = torch.zeros((chunk, n, w), dtype=torch.int32, device="cuda")
        lb = torch.zeros_like(la)
        a.rollout_random(chunk, seed=9, actions_log=la)
        b.rollout_random(chunk, seed=9, actions_log=lb)
        torch.cuda.synchronize()
        assert torch.equal(la, lb), (name, chunk, "actions")
        assert a.lib.pcb_current_slice(a._h) == b.lib.pcb_current_slice(b._h)
        for k in a.obs_ring_tensors:
            assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, chunk, k)
        for k in a._res:
            assert torch.equal(a._res[k], b._res[k]), (name, chunk, k)
        sa, sb = a.state_dict(), b.state_dict()
        assert sa["step_no"] == sb["step_no"] and np.array_equal(sa["state"], sb["state"])
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] and sa["sum_length"] == sb["sum_length"]
    assert sa["sum_return"] == sb["sum_return"]  # integer-valued float64 sums: exact
    if autoreset:
        assert sa["episodes"] > 0
    # the fused kernel is one launch per call
    l0 = a.launch_count
    a.rollout_random(17, seed=1)
    assert a.launch_count == l0 + 1
    a.close(); b.close()


@pytest.mark.parametrize("name", ["A", "B", "S_rect", "S_sq", "REF_rect", "S_rect_many"])
def test_fused_small_rollout_against_the_oracle(name):
    """The logged actions of a fused rollout replayed through the oracle: every
    observation of every step (ring of T+1 slices) must be the oracle's."""
    from parity import run_rollout_parity

    assert run_rollout_parity(name) > 0


@pytest.mark.parametrize("name", ["A", "B", "C", "D"])
def test_reference_dtype_mode_equals_the_oracle_value_for_value(name):
    """obs_dtype="reference": float64 (float32 square) tensors equal to the oracle's arrays
    without any cast on the comparing side, incl. the exact float64 area ratio."""
    from ref_harness import make_oracle, pick_action

    cfg = CONFIGS[name]
    square = cfg["type"] == "square"
    inst = None if square else draw_instance(cfg, 3)
    env = make(name, 1, autoreset=False, instances=None if square else [inst], obs_dtype="reference")
    o = make_oracle(cfg)
    ref = o.reset() if square else o.reset(instance=inst)
    got = env.reset()
    rng = np.random.default_rng(0)
    for t in range(6):
        for k, v in ref.items():
            g = got[k][0].cpu().numpy()
            v = np.asarray(v)
            if k == "component_grid":
                g = g[: v.shape[0]]
            assert g.dtype == v.dtype and np.array_equal(g, v), (name, t, k)
        a = pick_action(rng, o.action_mask)
        ref, _, rd, _ = o.step(a)
        got, _, _, _ = env.step(torch.tensor([a], dtype=torch.int32, device="cuda"))
        if rd:
            break
    assert env.single_action_space == tuple(o.action_mask.shape)
    sp = env.single_observation_space
    assert set(sp.keys() if isinstance(sp, dict) else sp.spaces.keys()) == set(ref.keys())
    env.close()


def test_fused_rollout_is_refused_for_wide_boards():
    for name in ("S_wide_pin", "S_wide_spa", "S_wide_sq", "S_wide_rect"):
        env = make(name, 8)
        assert not env.fused_rollout(True)
        env.close()


def test_step_host_and_flat_actions_match_device_step():
    import pcbgpu

    a, b, c = make("C", 300, seed=4), make("C", 300, seed=4), make("C", 300, seed=4)
    a.reset(); b.reset(); c.reset()
    rew = np.zeros(300); done = np.zeros(300, np.uint8)
    for t in range(14):
        act = a.sample_random_actions(seed=1).clone()
        _, r, d, _ = a.step(act)
        torch.cuda.synchronize()
        b.step_host(act.cpu().numpy(), rew, done)
        flat = (act[:, 0].long() * 32 + act[:, 1]) * 32 + act[:, 2]
        assert torch.equal(c.decode_flat_actions(flat), act)
        _, r3, d3, _ = c.step(flat.int())
        torch.cuda.synchronize()
        assert np.array_equal(r.cpu().numpy(), rew) and np.array_equal(d.cpu().numpy(), done)
        assert torch.equal(r, r3) and torch.equal(d, d3)
    assert torch.equal(a.action_mask_flat, c.action_mask.reshape(300, -1))
    for e in (a, b, c):
        e.close()
    assert pcbgpu.PcbVecEnv is not None


def test_episode_statistics_and_partial_reset():
    env = make("C", 128, autoreset=True, seed=8)
    env.reset()
    n_done, ret = 0, 0.0
    for t in range(35):
        _, r, d, info = env.step(env.sample_random_actions(seed=2))
        torch.cuda.synchronize()
        n_done += int(d.sum())
        ret += float(r.sum())
        if d.any():
            i = int(torch.nonzero(d)[0])
            assert int(info["episode_length"][i]) == 10
    st = env.episode_stats()
    assert st["episodes"] == n_done == 128 * 3
    assert np.isclose(st["sum_return"], ret) and st["sum_length"] == 10 * n_done
    assert st["steps"] == 35 * 128
    env.close()
    env = make("B", 8, autoreset=False, seed=1)
    env.reset()
    env.rollout_random(3, seed=0)
    occ = env.occupancy()
    mask = torch.tensor([1, 0, 0, 1, 0, 0, 0, 0], dtype=torch.uint8)
    obs = env.reset(mask=mask)
    occ2 = env.occupancy()
    assert not occ2[0].any() and not occ2[3].any() and np.array_equal(occ[1], occ2[1])
    assert int(obs["placement_mask"][0].sum()) == 0 and int(obs["placement_mask"][1].sum()) == 3
    env.close()


def test_rollout_collector_zero_copy_ring():
    import pcbgpu

    env = make("D", 64, obs_ring=9, seed=3)
    col = pcbgpu.RolloutCollector(env, horizon=8, seed=5)
    assert col.zero_copy
    out = col.collect()
    assert out["obs"]["pin_grid"].shape == (9, 64, 32, 32, 6)
    assert out["actions"].shape == (8, 64, 3) and out["rewards"].shape == (8, 64)
    # the observation stored for step t is the one the action of step t was valid for
    t = 5
    a = out["actions"][t].long()
    m = out["obs"]["action_mask"][t]
    assert bool((m[torch.arange(64), a[:, 0], a[:, 1], a[:, 2]] == 1).all())
    env.close()
    env = make("B", 32, obs_ring=1)
    out = pcbgpu.RolloutCollector(env, policy=lambda o: env.sample_random_actions(seed=1), horizon=4).collect()
    assert out["obs"]["grid"].shape == (5, 32, 16, 16)
    env.close()


@pytest.mark.parametrize("name", ["B", "C", "C_both", "D", "A"])
def test_logged_rollout_equals_stepping_by_hand(name):
    """RolloutCollector without a policy = one pcb_rollout_random_logged call (one launch for
    the pin-less kinds): observations, actions, rewards and dones of every step must equal
    a hand-written {sample; step} loop, and results() must describe the last step."""
    import pcbgpu

    N, T = 300, 13
    a, b = make(name, N, obs_ring=T + 1, seed=4), make(name, N, obs_ring=T + 1, seed=4)
    col = pcbgpu.RolloutCollector(a, horizon=T, seed=6)
    assert col.zero_copy
    for rep in range(2):  # the second collect continues the episodes and wraps the ring
        out = col.collect(reset=(rep == 0))
        obs = b.reset() if rep == 0 else b._obs()
        for t in range(T):
            for k, v in obs.items():
                assert torch.equal(out["obs"][k][t], v), (name, rep, t, k)
            act = b.sample_random_actions(seed=6)
            assert torch.equal(out["actions"][t], act), (name, rep, t)
            obs, r, d, _ = b.step(act)
            assert torch.equal(out["rewards"][t], r) and torch.equal(out["dones"][t], d), (name, rep, t)
        for k, v in obs.items():
            assert torch.equal(out["obs"][k][T], v), (name, rep, "last", k)
        ra, rb = a.results(), b.results()
        for k in ("reward", "done"):
            assert torch.equal(ra[k], rb[k]), (name, rep, k)
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] > 0 and sa["sum_length"] == sb["sum_length"]
    a.close(); b.close()


@pytest.mark.parametrize("name,n", [("C", 20000), ("C_both", 6000), ("D", 6000), ("E", 3000), ("B", 20000)])
def test_long_rollouts_do_not_depend_on_the_call_pattern(name, n):
    """600 steps as one call, as many short calls, and with the reward kernels serialised:
    the overlapped reward pipeline (double-buffered jobs and results, counter ring, two
    streams) must not change a single byte of state, observations, results or statistics."""
    T = 600
    envs = [make(name, n, seed=9, obs_ring=2, bank_size=3 * n) for _ in range(3)]
    for e in envs:
        e.reset()
    envs[0].rollout_random(T, seed=4)
    done = 0
    for chunk in [1, 2, 3, 5, 7, 11, 13, 17, 64, 1, 1, 100] * 10:
        chunk = min(chunk, T - done)
        if chunk == 0:
            break
        envs[1].rollout_random(chunk, seed=4)
        done += chunk
    assert done == T
    envs[2].reward_overlap(False)
    envs[2].rollout_random(T, seed=4)
    torch.cuda.synchronize()
    ref_state, ref_stats = envs[0].state_dict(), envs[0].episode_stats()
    for e in envs[1:]:
        sd, st = e.state_dict(), e.episode_stats()
        assert sd["step_no"] == ref_state["step_no"] and np.array_equal(sd["state"], ref_state["state"])
        assert st["episodes"] == ref_stats["episodes"] > 0 and st["sum_length"] == ref_stats["sum_length"]
        for k in ("sum_return", "sum_wirelength", "sum_intersections"):  # atomics: order differs
            assert np.isclose(st[k], ref_stats[k], rtol=1e-9, atol=0), (name, k)
        for k in envs[0].obs_ring_tensors:
            assert torch.equal(e.obs_ring_tensors[k], envs[0].obs_ring_tensors[k]), (name, k)
        for k in envs[0]._res:
            assert torch.equal(e._res[k], envs[0]._res[k]), (name, k)
    for e in envs:
        e.close()


def test_host_stepper_equals_step_host():
    n, T = 200, 9
    a, b = make("C", n, seed=2), make("C", n, seed=2)
    a.reset(); b.reset()
    log = torch.zeros((T, n, 3), dtype=torch.int32, device="cuda")
    c = make("C", n, seed=2)
    c.reset(); c.rollout_random(T, seed=1, actions_log=log); c.close()
    flat = ((log[..., 0] * 32 + log[..., 1]) * 32 + log[..., 2]).int().cpu().pin_memory()
    ra, da = torch.zeros(n, dtype=torch.float64).pin_memory(), torch.zeros(n, dtype=torch.uint8).pin_memory()
    rb, db = np.zeros(n, np.float64), np.zeros(n, np.uint8)
    f = a.host_stepper(flat, ra, da)
    for t in range(T):
        f(t)
        b.step_host(flat[t].numpy(), rb, db)
        assert np.array_equal(ra.numpy(), rb) and np.array_equal(da.numpy(), db), t
    with pytest.raises(IndexError):
        f(T)
    a.close(); b.close()


def test_errors_are_loud():
    import pcbgpu
    from pcbgpu._lib import PcbError

    env = make("C", 4)
    with pytest.raises(ValueError):
        env.step(torch.zeros((4, 2), dtype=torch.int32))
    with pytest.raises(PcbError, match="n_comp"):
        env.set_instances([dict(comp_h=np.zeros(0), comp_w=np.zeros(0), pin_comp=np.zeros(0),
                                pin_rx=np.zeros(0), pin_ry=np.zeros(0), pin_net=np.zeros(0),
                                pin_id=np.zeros(0), n_nets=0)])
    env.close()
    with pytest.raises(PcbError, match="grid"):
        pcbgpu.PcbVecEnv(dict(CONFIGS["A"], height=80, width=80), 4)


@pytest.mark.parametrize("name", ["D", "C", "B", "A"])
def test_ppo_collector_end_to_end(name):
    """SURVEY 8f N2 / BASELINE config 4: a torch policy drives the env through flat
    actions and the masked logits; every sampled action must be valid."""
    from pcbgpu.ppo import PPOTrainer, compute_gae

    torch.manual_seed(0)
    env = make(name, 128, autoreset=True, seed=3)
    tr = PPOTrainer(env, horizon=12, minibatches=2, epochs=1)
    before = [p.detach().clone() for p in tr.policy.parameters()]
    batch = tr.collect()
    T, N = batch["actions"].shape
    m = batch["obs"]["action_mask"].reshape(T, N, -1)
    assert bool((m.gather(2, batch["actions"].unsqueeze(2)) == 1).all())
    assert int(batch["dones"].sum()) > 0 and torch.isfinite(batch["adv"]).all()
    loss = tr.update(batch)
    assert np.isfinite(loss)
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.policy.parameters()))
    # GAE cuts the bootstrap at episode ends
    r = torch.ones(3, 1); v = torch.zeros(3, 1); d = torch.tensor([[0], [1], [0]], dtype=torch.bool)
    adv, _ = compute_gae(r, v, d, torch.zeros(1), gamma=1.0, lam=1.0)
    assert adv[:, 0].tolist() == [2.0, 1.0, 1.0]
    env.close()


def test_vector_env_adapter_surface():
    """the RLlib VectorEnv method surface (duck-typed, ray is not installed)."""
    import pcbgpu

    ad = pcbgpu.VectorEnvAdapter(CONFIGS["REF_pin"], 6, seed=2)
    obs = ad.vector_reset()
    assert len(obs) == 6 and obs[0]["action_mask"].shape == (400,) and obs[0]["grid"].dtype == np.float64
    rng = np.random.default_rng(0)
    for t in range(5):
        acts = [int(rng.choice(np.flatnonzero(o["action_mask"]))) for o in obs]
        obs, rew, done, infos = ad.vector_step(acts)
        assert len(rew) == 6 and all(isinstance(d, bool) for d in done)
    assert all(done) and all(set(i) == {"wirelength", "num_intersections"} for i in infos)
    o3 = ad.reset_at(3)
    assert o3["grid"].sum() == 0 and obs[2]["grid"].sum() > 0
    ad.close()


def test_checkpoint_resume_is_deterministic():
    env = make("C_both", 256, seed=6)
    env.reset()
    env.rollout_random(7, seed=5)
    sd = env.state_dict()

    def run():
        out = []
        for t in range(15):
            _, r, d, _ = env.step(env.sample_random_actions(seed=5))
            out.append((r.clone(), d.clone()))
        return out

    a = run()
    env.load_state_dict(sd)
    b = run()
    assert all(torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) for x, y in zip(a, b))
    dec = env.decode_state()
    assert len(dec) == 256 and all(len(e["components"]) == 10 for e in dec)
    env.close()


def test_export_in_reference_pickle_format(tmp_path):
    import pickle

    from pcbgpu.export import export_rollouts

    comps, actions = export_rollouts(CONFIGS["REF_pin"], 5, str(tmp_path), seed=3)
    assert len(comps) == 5 and all(len(a) == 5 and len(a[0]) == 3 for a in actions)
    raw = open(tmp_path / "components.pkl", "rb").read()
    assert b"environment.dummy_env_rectangular_pin" in raw and b"Component" in raw and b"Pin" in raw
    acts = pickle.load(open(tmp_path / "actions.pkl", "rb"))
    assert acts == actions
    for ep_comps, ep_actions in zip(comps, actions):
        for c, (o, x, y) in zip(ep_comps, ep_actions):
            assert c.placed and c.position == (x, y)
            ph, pw = (c.h, c.w) if o in (0, 2) else (c.w, c.h)
            for pin in c.pins:  # pins lie inside the placed footprint
                assert x <= pin.absolute_x < x + ph and y <= pin.absolute_y < y + pw


def test_specialised_library_for_an_arbitrary_config():
    """pcbgpu can compile the kernels for exactly one configuration (-DPCB_USER_SHAPE);
    the result must be bit-identical to the oracle like every other build."""
    env = make("S_pin", 8, specialize=True)
    assert env.shape_id == 100
    env.close()
    assert make("S_pin", 8).shape_id == 0 and make("C", 8).shape_id == 2
    st = {}
    run_parity("S_pin", range(32), extra_steps=2, stats=st, specialize=True)
    run_parity("S_pin", range(16), autoreset=True, episodes=2, p_invalid=0.1, stats=st, specialize=True)
    assert st["reward_inexact"] == 0


def test_specialised_library_for_a_pin_less_config():
    """... and for a rectangular configuration: the user shape also instantiates the
    multi-env-per-warp rollout kernel (12x14 board: 16 lanes per env, dense-bit-string boards)."""
    import __graft_entry__ as ge

    assert CONFIGS["S_rect"] == ge.SPECIALISED_DEMO_RECT  # the library build() prepared
    env = make("S_rect", 100, specialize=True, obs_ring=3, seed=1)
    assert env.shape_id == 100 and env.fused_rollout(True)
    ref = make("S_rect", 100, obs_ring=3, seed=1)  # runtime-shaped kernels of the default library
    assert ref.shape_id == 0
    env.reset(); ref.reset()
    env.rollout_random(40, seed=3); ref.rollout_random(40, seed=3)
    for k in env.obs_ring_tensors:
        assert torch.equal(env.obs_ring_tensors[k], ref.obs_ring_tensors[k]), k
    assert np.array_equal(env.state_dict()["state"], ref.state_dict()["state"])
    env.close(); ref.close()
    run_parity("S_rect", range(24), extra_steps=2, p_invalid=0.2, specialize=True)


@pytest.mark.parametrize("name", ["C", "C_both", "D", "E", "B", "A", "S_rect", "S_pin", "S_wide_rect"])
@pytest.mark.parametrize("tuples", [False, True])
def test_steps_host_equals_repeated_step_host(name, tuples):
    """pcb_steps_host (T rows in one call: copies overlapped with neighbouring steps, pin-less
    kinds in one launch) == T x pcb_step_host, for every result row, the observations of every
    step (ring), the result tensors and the episode statistics."""
    n, T = 77, 19  # more than two upload chunks, odd sizes
    a = make(name, n, seed=6, obs_ring=T + 1)
    b = make(name, n, seed=6, obs_ring=T + 1)
    c = make(name, n, seed=6)
    w = a.spec.action_width
    log = torch.zeros((T, n, w), dtype=torch.int32, device="cuda")
    c.reset(); c.rollout_random(T, seed=3, actions_log=log); c.close()
    log[T // 2, ::5] = -1          # some invalid actions
    H, W = a.spec.height, a.spec.width
    if tuples:
        acts = log.cpu().pin_memory()
    else:
        flat = (log[..., 0] * H + log[..., 1]) * W + log[..., 2] if w == 3 else log[..., 0] * W + log[..., 1]
        flat[T // 2, ::5] = -7
        acts = flat.int().cpu().pin_memory()
    a.reset(); b.reset()
    R = torch.zeros((T, n), dtype=torch.float64).pin_memory()
    D = torch.zeros((T, n), dtype=torch.uint8).pin_memory()
    a.steps_host(acts, R, D)
    f = b.host_stepper(acts)        # packed result block: one D2H copy per step
    for t in range(T):
        f(t)
        assert np.array_equal(f.reward.numpy(), R[t].numpy()), (name, t)
        assert np.array_equal(f.done.numpy(), D[t].numpy()), (name, t)
    torch.cuda.synchronize()
    for k in a.obs_ring_tensors:
        assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, k)
    ra, rb = a.results(), b.results()
    fin = ra["done"].bool()
    for k in ra:  # the info tensors are defined where done (pcb_env.h)
        if k in ("reward", "done"):
            assert torch.equal(ra[k], rb[k]), (name, k)
        else:
            assert torch.equal(ra[k][fin], rb[k][fin]), (name, k)
    sa, sb = a.episode_stats(), b.episode_stats()
    for k in sa:  # the sums are float64 atomics: the order of the additions is not fixed
        assert sa[k] == sb[k] if k in ("episodes", "steps", "sum_length") else np.isclose(sa[k], sb[k], rtol=1e-12), k
    assert int(D.sum()) > 0
    # a second call continues the same trajectories (staging reused, parity of the halves)
    a.steps_host(acts[:3], R[:3], D[:3])
    for t in range(3):
        f(t)
        assert np.array_equal(f.reward.numpy(), R[t].numpy()) and np.array_equal(f.done.numpy(), D[t].numpy())
    a.close(); b.close()


def test_steps_host_without_reward_overlap_and_argument_checks():
    n, T = 64, 5
    a, b = make("C", n, seed=1), make("C", n, seed=1)
    a.reward_overlap(False)
    acts = torch.zeros((T, n), dtype=torch.int32).pin_memory()
    acts[:] = torch.arange(n, dtype=torch.int32) % 700
    Ra, Da = torch.zeros((T, n), dtype=torch.float64), torch.zeros((T, n), dtype=torch.uint8)
    Rb, Db = torch.zeros((T, n), dtype=torch.float64), torch.zeros((T, n), dtype=torch.uint8)
    a.reset(); b.reset()
    a.steps_host(acts, Ra, Da)      # pageable host memory works too
    b.steps_host(acts, Rb, Db)
    assert torch.equal(Ra, Rb) and torch.equal(Da, Db)
    with pytest.raises(ValueError):
        a.steps_host(acts, Ra[:2], Da)
    with pytest.raises(ValueError):
        a.steps_host(acts.long(), Ra, Da)
    with pytest.raises(ValueError):
        a.steps_host(acts, Ra.float(), Da)
    a.close(); b.close()


@pytest.mark.parametrize("name,ring", [("C", 4), ("C_both", 2), ("D", 5), ("S_pin", 3), ("S_wide_rect", 4)])
def test_graph_rollout_equals_plain_launches(name, ring):
    """pcb_rollout_random of the per-step kinds replays a CUDA graph of L steps; the result must
    be the plain per-step launches' bit for bit: every observation slice, the state records,
    the results of the last step, the episode statistics -- across several calls, a partial
    reset in between and a bank upload (which invalidates the captured graph)."""
    n = 200
    a = make(name, n, seed=9, obs_ring=ring, bank_size=3 * n)
    b = make(name, n, seed=9, obs_ring=ring, bank_size=3 * n)
    b.graph_rollout(False)
    a.fused_rollout(False); b.fused_rollout(False)  # the per-step launches are what a graph replays
    a.reset(); b.reset()
    torch.cuda.synchronize()
    with torch.cuda.stream(torch.cuda.Stream()):  # the default stream cannot be captured
        _graph_rollout_body(name, n, a, b)
    a.close(); b.close()


def _graph_rollout_body(name, n, a, b):
    def same(tag):
        torch.cuda.synchronize()
        for k in a.obs_ring_tensors:
            assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, tag, k)
        assert np.array_equal(a.state_dict()["state"], b.state_dict()["state"]), (name, tag)
        ra, rb = a.results(), b.results()
        for k in ("reward", "done"):
            assert torch.equal(ra[k], rb[k]), (name, tag, k)
        sa, sb = a.episode_stats(), b.episode_stats()
        assert sa["episodes"] == sb["episodes"] and sa["sum_length"] == sb["sum_length"], (name, tag)
        assert np.isclose(sa["sum_return"], sb["sum_return"], rtol=1e-12), (name, tag)

    for e in (a, b):
        e.rollout_random(71, seed=5)      # several replays + a remainder
    same("first")
    assert a.graph_replays >= 1 and b.graph_replays == 0
    r0 = a.graph_replays
    for e in (a, b):
        e.rollout_random(40, seed=5)      # the captured graph is reused when the phase allows
        e.reset(mask=(torch.arange(n, device="cuda") % 7 == 0))
        e.rollout_random(33, seed=6)      # another seed: recaptured
    same("second")
    assert a.graph_replays > r0
    for e in (a, b):
        e.generate_instances(2 * n, seed=123)   # new bank: the graph carried the old pointer
        e.rollout_random(50, seed=6)
        e.step(e.sample_random_actions(seed=1))  # plain step in between
        e.rollout_random(37, seed=6)
    same("third")


@pytest.mark.parametrize("name", ["C", "REF_pin", "S_pin_tight", "S_pin_cen", "FZ_rect_pin_0_cen",
                                  "FZ_rect_pin_3_cen", "FZ_rect_pin_4_cen", "FZ_rect_pin_6_cen",
                                  "D", "REF_spatial", "S_spa_tight", "S_spa_cen", "FZ_rect_spatial_pin_4_cen",
                                  "C_both", "C_beam", "D_both", "S_pin", "S_pin_beam3", "S_spa", "FZ_rect_pin_0",
                                  "FZ_rect_pin_1", "FZ_rect_pin_3", "FZ_rect_pin_4", "FZ_rect_pin_5", "FZ_rect_pin_6"])
@pytest.mark.parametrize("autoreset", [True, False])
def test_pin_fused_rollout_equals_per_step(name, autoreset):
    """rollout_pin_kernel (many steps per launch, 32/LPE envs per warp, rewards from a job list
    afterwards) against the per-step kernels: every observation slice, the logged actions,
    rewards and dones of every step, the state records, the result tensors and the statistics."""
    cfg = CONFIGS[name]
    if cfg["type"] not in ("rectangle_pin", "rectangle_spatial_pin") or max(cfg["height"], cfg["width"]) > 32:
        pytest.skip("not a configuration of the fused pin kernel")
    n, T = 203, 41
    a = make(name, n, seed=11, obs_ring=T + 1, bank_size=5 * n, autoreset=autoreset)
    b = make(name, n, seed=11, obs_ring=T + 1, bank_size=5 * n, autoreset=autoreset)
    assert a.fused_rollout(True)
    b.fused_rollout(False)
    a.reset(); b.reset()
    logs = []
    for e in (a, b):
        al = torch.zeros((T, n, 3), dtype=torch.int32, device="cuda")
        rl = torch.full((T, n), -7.0, dtype=torch.float64, device="cuda")
        dl = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        e.rollout_random(T, seed=3, actions_log=al, reward_log=rl, done_log=dl)
        logs.append((al, rl, dl))
    torch.cuda.synchronize()
    for x, y, what in zip(logs[0], logs[1], ("actions", "reward", "done")):
        assert torch.equal(x, y), (name, what, (x != y).nonzero()[:5].tolist())
    assert int(logs[0][2].sum()) > 0

    def same(tag):
        torch.cuda.synchronize()
        for k in a.obs_ring_tensors:
            assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, tag, k)
        assert np.array_equal(a.state_dict()["state"], b.state_dict()["state"]), (name, tag)
        ra, rb = a.results(), b.results()
        done = ra["done"].bool()
        assert torch.equal(ra["done"], rb["done"]) and torch.equal(ra["reward"], rb["reward"]), (name, tag)
        for k in ("wirelength", "num_intersections", "episode_return", "episode_length"):
            assert torch.equal(ra[k][done], rb[k][done]), (name, tag, k)
        sa, sb = a.episode_stats(), b.episode_stats()
        for k in sa:
            assert sa[k] == sb[k] if k in ("episodes", "steps", "sum_length") else np.isclose(sa[k], sb[k], rtol=1e-12), (name, tag, k)

    same("logged")
    for e in (a, b):                      # no logs, a length that needs several launches
        e.rollout_random(150, seed=4)
        e.step(e.sample_random_actions(seed=9))   # a per-step call in between shares the state
        e.rollout_random(7, seed=5)
    same("plain")
    # the caller's actions through the same kernel (pcb_steps_host)
    acts = logs[0][0][:9].cpu().pin_memory()
    Ra, Da = torch.zeros((9, n), dtype=torch.float64), torch.zeros((9, n), dtype=torch.uint8)
    Rb, Db = torch.zeros((9, n), dtype=torch.float64), torch.zeros((9, n), dtype=torch.uint8)
    a.steps_host(acts, Ra, Da); b.steps_host(acts, Rb, Db)
    assert torch.equal(Ra, Rb) and torch.equal(Da, Db), name
    same("given actions")
    a.close(); b.close()


def test_step_host_with_adjacent_result_buffers():
    """reward and done buffers that happen to be adjacent in memory (two pinned tensors out of one
    allocator block) must not be taken for a packed block: the packed copy is its own entry point."""
    n = 256
    a, b = make("C", n, seed=3), make("C", n, seed=3)
    a.reset(); b.reset()
    blk = torch.empty(n * 9 + 64, dtype=torch.uint8).pin_memory()
    rew = blk[: n * 8].view(torch.float64)
    done = blk[n * 8: n * 9]           # starts exactly where the rewards end
    acts = torch.zeros((6, n), dtype=torch.int32).pin_memory()
    acts[:] = (torch.arange(n, dtype=torch.int32) * 37) % 3000
    f = b.host_stepper(acts)           # packed block of its own
    for t in range(6):
        a.step_host(acts[t].numpy(), rew.numpy(), done.numpy())
        f(t)
        assert np.array_equal(rew.numpy(), f.reward.numpy()) and np.array_equal(done.numpy(), f.done.numpy()), t
    a.close(); b.close()


@pytest.mark.parametrize("name", ["C", "E", "C_both"])
def test_rollout_calls_reuse_the_actions_drawn_by_the_previous_call(name):
    """The last step of a pcb_rollout_random call also draws the actions of the step after it and
    the next call with the same seed starts from them instead of launching the stand-alone
    sampler.  Per-step kernels (plain launches on the default stream): one launch fewer per
    follow-up call, dropped by a reset / another seed / a step in between, and every
    trajectory equal to a hand-written {sample; step} loop, which never takes that path."""
    n = 150
    a, b = make(name, n, seed=3, obs_ring=3, bank_size=4 * n), make(name, n, seed=3, obs_ring=3, bank_size=4 * n)
    a.fused_rollout(False); b.fused_rollout(False)
    a.reset(); b.reset()
    width = a.spec.action_width

    def by_hand(steps, seed, log=None):
        for t in range(steps):
            act = b.sample_random_actions(seed=seed)
            if log is not None:
                assert torch.equal(log[t], act), (name, t)
            b.step(act)

    def same(tag):
        torch.cuda.synchronize()
        assert np.array_equal(a.state_dict()["state"], b.state_dict()["state"]), (name, tag)
        for k in a.obs_ring_tensors:
            assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, tag, k)
        ra, rb = a.results(), b.results()
        for k in ("reward", "done"):
            assert torch.equal(ra[k], rb[k]), (name, tag, k)

    def launches(fn):
        l0 = a.launch_count
        fn()
        return a.launch_count - l0

    first = launches(lambda: a.rollout_random(7, seed=5))
    by_hand(7, 5); same("first call")
    again = launches(lambda: a.rollout_random(7, seed=5))
    by_hand(7, 5); same("second call")
    assert again == first - 1                      # no sampler launch
    log = torch.zeros((9, n, width), dtype=torch.int32, device="cuda")
    logged = launches(lambda: a.rollout_random(9, seed=5, actions_log=log))
    by_hand(9, 5, log); same("logged call")        # first row of the log = the kept actions
    assert logged * 7 == again * 9
    other = launches(lambda: a.rollout_random(7, seed=6))
    by_hand(7, 6); same("another seed")
    assert other == first                          # sampled again
    a.reset(mask=(torch.arange(n, device="cuda") % 3 == 0))
    b.reset(mask=(torch.arange(n, device="cuda") % 3 == 0))
    after_reset = launches(lambda: a.rollout_random(7, seed=6))
    by_hand(7, 6); same("after a partial reset")
    assert after_reset == first
    act = a.sample_random_actions(seed=9)
    a.step(act); b.step(b.sample_random_actions(seed=9))
    after_step = launches(lambda: a.rollout_random(7, seed=6))
    by_hand(7, 6); same("after a plain step")
    assert after_step == first
    a.close(); b.close()


@pytest.mark.parametrize("name", ["C_both", "C_beam", "D_both"])
@pytest.mark.parametrize("autoreset", [True, False])
def test_list_steps_producer_equals_per_step(name, autoreset, monkeypatch):
    """Reward types beam / both at large batches: per-step kernels fill the job list of a chunk,
    router + reward kernel run over the list (list_steps_chunk, pcb_api.cu).  Forced here at a small
    batch (PCB_LIST_STEPS_MIN_ENVS) and compared with the plain per-step path: observation slices,
    logged actions / rewards / dones, state records, result tensors, statistics, and the caller's
    actions through pcb_steps_host."""
    cfg = CONFIGS[name]
    if cfg["type"] not in ("rectangle_pin", "rectangle_spatial_pin") or cfg["height"] <= 16:
        pytest.skip("boards of at most 16 rows always take rollout_pin_kernel")
    n, T = 203, 41
    monkeypatch.setenv("PCB_LIST_STEPS_MIN_ENVS", "1")
    a = make(name, n, seed=11, obs_ring=T + 1, bank_size=5 * n, autoreset=autoreset)
    monkeypatch.delenv("PCB_LIST_STEPS_MIN_ENVS")
    b = make(name, n, seed=11, obs_ring=T + 1, bank_size=5 * n, autoreset=autoreset)
    assert a.fused_rollout(True)
    b.fused_rollout(False)
    a.reset(); b.reset()
    l0 = a.launch_count
    logs = []
    for e in (a, b):
        al = torch.zeros((T, n, 3), dtype=torch.int32, device="cuda")
        rl = torch.full((T, n), -7.0, dtype=torch.float64, device="cuda")
        dl = torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        e.rollout_random(T, seed=3, actions_log=al, reward_log=rl, done_log=dl)
        logs.append((al, rl, dl))
    torch.cuda.synchronize()
    # sampler + T step kernels + router + reward list kernel: the producer really ran per step
    assert a.launch_count - l0 == T + 3, a.launch_count - l0
    for x, y, what in zip(logs[0], logs[1], ("actions", "reward", "done")):
        assert torch.equal(x, y), (name, what, (x != y).nonzero()[:5].tolist())
    assert int(logs[0][2].sum()) > 0

    def same(tag):
        torch.cuda.synchronize()
        for k in a.obs_ring_tensors:
            assert torch.equal(a.obs_ring_tensors[k], b.obs_ring_tensors[k]), (name, tag, k)
        assert np.array_equal(a.state_dict()["state"], b.state_dict()["state"]), (name, tag)
        ra, rb = a.results(), b.results()
        done = ra["done"].bool()
        assert torch.equal(ra["done"], rb["done"]) and torch.equal(ra["reward"], rb["reward"]), (name, tag)
        for k in ("wirelength", "num_intersections", "episode_return", "episode_length"):
            assert torch.equal(ra[k][done], rb[k][done]), (name, tag, k)
        sa, sb = a.episode_stats(), b.episode_stats()
        for k in sa:
            assert sa[k] == sb[k] if k in ("episodes", "steps", "sum_length") else np.isclose(sa[k], sb[k], rtol=1e-12), (name, tag, k)

    same("logged")
    for e in (a, b):                      # no logs, a length that needs several chunks
        e.rollout_random(150, seed=4)
        e.step(e.sample_random_actions(seed=9))   # a per-step call in between shares the state
        e.rollout_random(7, seed=5)
    same("plain")
    acts = logs[0][0][:9].cpu().pin_memory()
    Ra, Da = torch.zeros((9, n), dtype=torch.float64), torch.zeros((9, n), dtype=torch.uint8)
    Rb, Db = torch.zeros((9, n), dtype=torch.float64), torch.zeros((9, n), dtype=torch.uint8)
    a.steps_host(acts, Ra, Da); b.steps_host(acts, Rb, Db)
    assert torch.equal(Ra, Rb) and torch.equal(Da, Db), name
    same("given actions")
    a.close(); b.close()
