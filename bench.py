#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched placement environments on B200.

Metric (BASELINE.json): env-steps/sec, whole job, device-timed, on the random-
policy rollout { device random policy -> step -> terminal reward -> auto-reset }.
Main line: config C (dummy_env_rectangular_pin 32x32, 10 components, 5 nets,
centroid reward, 65 536 envs per GPU -- the configuration the metric is quoted on);
weak scaling over GPUs, no collective on the step path.  One JSON line on rank 0:
  value        device-timed, actions sampled on device, everything resident in HBM
  e2e          the same rollout replayed through host buffers: every step's action row
               comes from pinned host memory (H2D) and its reward + done go back (D2H)
  roofline     the step kernel: algorithmic bytes per launch / its CUDA-event time
  other_workloads   the other BASELINE.json configs measured the same way at the same N:
               B (configs[1], 4 096 envs), D (configs[3]), E (configs[4], 131 072 envs/GPU)
  cpu_baseline the reference's own Python environments (baseline/_ref, staged by build())
               on the host cores as a multiprocess vector env, bounded sample; plus
               configs[0] (A, single env, one process)
`--impl reference` times that CPU path alone with all host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C"
ENVS_PER_GPU = 65536
DEFAULT_ENVS = {"A": 65536, "B": 4096, "C": 65536, "D": 65536, "E": 131072, "C_both": 65536,
                "C_beam": 65536, "REF_pin": 65536, "REF_rect": 65536, "REF_spatial": 65536,
                "REF_square": 65536}
TRAFFIC_ENVS = {"B": 65536, "C": 65536, "D": 65536, "E": 131072}  # profiles/traffic.json
SECONDARY = ("B", "D", "E")  # BASELINE.json configs[1], [3], [4] at their BASELINE sizes
OBS_RING = 4
REPETITIONS = 5
E2E_REPETITIONS = 3  # wall-clock passes of the end-to-end legs (median reported)
HOST_ROWS = 32  # rows per pcb_steps_host call in the e2e pass
# SURVEY.md section 8d / BASELINE.md section 5: algorithmic bytes per env-step (compact contract)
ALGO_BYTES = {"A": 161, "B": 1049, "C": 7699, "D": 15775, "E": 28821,
              # the reference's shipped 10x10 pin config (agent/config/rectangle_pin_model.json):
              # grid 100 + mask 400 + features 100 + placement 5 + pins 160 + 40 + 21 + 2 * 192 state
              "REF_pin": 1210, "C_both": 7699, "C_beam": 7699,
              # rectangle_model.json: grid 100 + mask 200 + features 400 + masks 40 + 21 + 2 * 240 state
              "REF_rect": 1241,
              # spatial twin of REF_pin: + pin_grid 400 + component_grid 80, features 5 x 9 x 4,
              # pin rows (21 x 4 + 21 x 2) x 2
              "REF_spatial": 1818, "REF_square": 229}
EP_LEN = {"A": 4, "B": 6, "C": 10, "D": 10, "E": 32, "REF_pin": 5, "C_both": 10, "C_beam": 10,
          "REF_rect": 8, "REF_spatial": 5, "REF_square": 25}
WORKLOAD_DESC = {
    "A": "A: dummy_env_square 8x8, 3x3 components",
    "B": "B: dummy_env_rectangular 16x16, 6 rectangular components",
    "C": "C: dummy_env_rectangular_pin 32x32, 10 components, 5 nets, <=6 pins/net, centroid reward",
    "D": "D: dummy_env_rectangular_pin_spatial 32x32, 10 components, 5 nets, centroid reward",
    "E": "E: dummy_env_rectangular_pin 64x64, 32 components, 16 nets, centroid reward",
    "C_both": "C_both: as C with reward_type='both' (beam-search route + centroid route)",
    "C_beam": "C_beam: as C with reward_type='beam'",
    "REF_rect": "REF_rect: dummy_env_rectangular 10x10, 20 components 2..4 (agent/config/rectangle_model.json)",
    "REF_pin": "REF_pin: dummy_env_rectangular_pin 10x10, 5 components 2x2, 3 nets x 6 pins (agent/config/rectangle_pin_model.json)",
    "REF_spatial": "REF_spatial: dummy_env_rectangular_pin_spatial 10x10 (agent/config/rectangle_pin_spatial_model.json)",
    "REF_square": "REF_square: dummy_env_square 10x10, 2x2 components (agent/config/square_model.json)",
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------
# CPU arm: the reference's own environments (baseline/_ref) or, when they were not
# staged, the oracle port -- as a multiprocess vector env
# ----------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_RECT_ARGS = ("height", "width", "min_component_w", "max_component_w", "min_component_h",
              "max_component_h", "max_num_components", "min_num_components")
_PIN_ARGS = ("height", "width", "net_distribution", "pin_spread", "min_component_w",
             "max_component_w", "min_component_h", "max_component_h", "max_num_components",
             "min_num_components", "min_num_nets", "max_num_nets", "max_num_pins_per_net",
             "min_num_pins_per_net", "reward_type", "reward_beam_width", "weight_wirelength",
             "weight_num_intersections")
_REF_MODULES = {"square": "environment.dummy_env_square",
                "rectangle": "environment.dummy_env_rectangular",
                "rectangle_pin": "environment.dummy_env_rectangular_pin",
                "rectangle_spatial_pin": "environment.dummy_env_rectangular_pin_spatial"}


def cpu_kind():
    """"reference": the unmodified reference environments staged under baseline/_ref by
    __graft_entry__.build(); "port": the oracle/ restatement (nothing was staged)."""
    return "reference" if os.path.isfile(os.path.join(REF_DIR, "environment", "dummy_env_square.py")) else "port"


def load_configs():
    """pcbgpu/configs.py as a plain module (the CPU workers do not need torch or the package)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "pcbgpu_configs", os.path.join(ROOT, "rl-environment-for-component-placement_b200", "configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CONFIGS


def make_cpu_env(name):
    cfg = load_configs()[name]
    kind = cfg["type"]
    args = ([cfg[k] for k in ("height", "width", "component_n")] if kind == "square" else
            [cfg[k] for k in (_RECT_ARGS if kind == "rectangle" else _PIN_ARGS)])
    if cpu_kind() == "reference":
        import importlib

        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)  # the reference packages + the gym stand-in they import
        return importlib.import_module(_REF_MODULES[kind]).DummyPlacementEnv(*args)
    import oracle  # the CPU legs of bench.py may time the oracle port (never the product path)

    if kind == "square":
        return oracle.SquareEnv(*args)
    if kind == "rectangle":
        return oracle.RectEnv(*args)
    return oracle.PinEnv(*args, spatial=(kind == "rectangle_spatial_pin"))


def _cpu_slices(args):
    """One worker: n_slices consecutive time slices of the random-policy loop
    (agent/random/random_policy_*.py: reset -> { pick uniformly among the valid actions ->
    step } until done); returns (env-steps, seconds) of every slice."""
    name, wid, n_slices, slice_s = args
    import random

    import numpy as np

    np.random.seed(wid)
    random.seed(wid)
    rng = np.random.default_rng(wid)
    env = make_cpu_env(name)
    out = []
    env.reset()
    done = False
    for _ in range(n_slices):
        steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < slice_s:
            if done:
                env.reset()
                done = False
            mask = np.asarray(env.action_mask)
            valid = np.flatnonzero(mask)
            if len(valid):
                a = np.unravel_index(valid[rng.integers(0, len(valid))], mask.shape)
            else:
                a = (0,) * mask.ndim
            _, _, done, _ = env.step(tuple(int(v) for v in a))
            steps += 1
        out.append((steps, time.perf_counter() - t0))
    return out


def cpu_rate(name, budget_s, procs):
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_slices((name, 0, 1, budget_s))]
    else:
        with ctx.Pool(procs) as pool:
            res = pool.map(_cpu_slices, [(name, w, 1, budget_s) for w in range(procs)])
    wall = time.perf_counter() - t0
    rate = sum(r[0][0] / r[0][1] for r in res)
    return rate, sum(r[0][0] for r in res), wall


def cpu_what():
    return ("the reference's own DummyPlacementEnv classes (unmodified, baseline/_ref)"
            if cpu_kind() == "reference" else
            "oracle/ numpy port of the reference envs (the reference was not staged into baseline/_ref)")


def cpu_baseline(name, budget_s, procs=None):
    procs = procs or os.cpu_count() or 1
    rate, steps, wall = cpu_rate(name, budget_s, procs)
    out = {"value": rate, "unit": "env-steps/s", "cores": procs, "kind": cpu_kind(),
           "sample": f"config {name}, random valid policy, {procs} processes x {budget_s:.0f} s "
                     f"({steps} env-steps, wall {wall:.1f} s); {cpu_what()}"}
    # BASELINE.json configs[0]: the square env, single env, one process -- the reference path
    # of experiments/random_policy/run_policy_square.py
    a_rate, a_steps, a_wall = cpu_rate("A", min(5.0, budget_s), 1)
    out["configs0_A_single_env"] = {
        "value": a_rate, "unit": "env-steps/s", "cores": 1, "kind": cpu_kind(),
        "sample": f"{WORKLOAD_DESC['A']}, one env, one process, {a_steps} env-steps in {a_wall:.1f} s"}
    return out


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (SM clock, max SM clock, clock-event reasons) from a thread while the
    timed region runs; the region lasts only tens of milliseconds, so nvidia-smi's
    100 ms loop is too coarse."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.bits, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.thread.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(sm),
                "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b)}


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def stagger(env, name, seed):
    """De-synchronise the episodes: after this, env i is (EP_LEN-1 - i % EP_LEN) steps
    into its episode, so every step finishes ~1/EP_LEN of the envs (steady state of a
    real rollout) instead of all of them every EP_LEN-th step."""
    import torch

    L = EP_LEN[name]
    ids = torch.arange(env.num_envs, device=env.device)
    for j in range(L):
        env.rollout_random(1, seed)
        env.reset(mask=((ids % L) == j))


class Ctx:
    """rank / world / barrier / max-reduce shared by every workload of one bench run"""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.local_rank = rank, world, local_rank

    def barrier(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, values, device):
        import torch
        import torch.distributed as dist

        t = torch.tensor(values, dtype=torch.float64, device=device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()


def measure(args, ctx, name, N, main):
    """All passes of one workload at N envs per GPU.  Every rank runs them; rank 0 gets the
    result dict (times are the max over ranks), the others None."""
    import torch

    import pcbgpu
    from pcbgpu.configs import CONFIGS

    rank, world = ctx.rank, ctx.world
    cfg = CONFIGS[name]
    K, W = args.steps, args.warmup
    # The observation ring must not fit in L2 (126 MB): small workloads get more slices.
    # The pin-less kinds run the whole rollout in one launch whose CTAs are not in lock-step:
    # a short ring would be overwritten inside L2 (measured: 12 % of the stores reached DRAM
    # with 5 slices), so there the ring keeps every observation of the timed rollout -- what a
    # rollout collector needs anyway -- as long as that stays below 48 GB.
    slice_bytes = N * ALGO_BYTES[name]
    ring = max(int(os.environ.get("PCB_BENCH_RING", OBS_RING)), -(-300_000_000 // slice_bytes))  # (env: experiments)
    pinless = cfg["type"] in ("square", "rectangle")
    if pinless and not args.no_fused:
        ring = max(ring, min(K + 1, 48_000_000_000 // slice_bytes))
    env = pcbgpu.PcbVecEnv(cfg, N, device=ctx.local_rank, autoreset=True, obs_ring=ring,
                           bank_size=N * 4, seed=1000 + rank, generate_on_device=True)
    if args.no_fused:
        env.fused_rollout(False)
    seed = 7 + rank
    stream = torch.cuda.current_stream()
    barrier = ctx.barrier

    def prepare():
        env.reset()
        stagger(env, name, seed)
        env.rollout_random(W, seed)
        env.episode_stats(reset=True)

    # ---- pass 1: device-timed, device policy (value); median of REPETITIONS ------------
    prepare()
    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    l0 = env.launch_count
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gate():
        # Device time, not host latency: a short spin kernel keeps the stream busy while the host
        # enqueues {event, rollout, event}; without it the first event fires on an idle stream
        # and the ~10 us the host needs to issue the launch are counted as device time -- a
        # quarter of the whole region when the pin-less kinds run 20 steps in one 40 us launch.
        torch.cuda._sleep(200_000)

    gate()
    e0.record(stream)
    env.rollout_random(K, seed)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = env.launch_count - l0
    reps_ms = [e0.elapsed_time(e1)]
    agg, agg_ms = None, None
    if world > 1 and main:
        # C1, the only collective of the design: the per-rollout all-reduce of the episode
        # statistics (NCCL over NVLink); off the step path, its cost is reported beside it
        pcbgpu.allreduce_episode_stats(env)  # first call sets up the communicator
        torch.cuda.synchronize()
        t_a = time.perf_counter()
        agg = pcbgpu.allreduce_episode_stats(env)
        torch.cuda.synchronize()
        agg_ms = (time.perf_counter() - t_a) * 1e3
    stats = env.episode_stats(reset=True)
    for _ in range(REPETITIONS - 1):  # the rollout just goes on
        barrier()
        gate()
        e0.record(stream)
        env.rollout_random(K, seed)
        e1.record(stream)
        barrier()
        reps_ms.append(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    env.episode_stats(reset=True)

    # ---- pass 2: same rollout with per-kernel CUDA events (roofline) + action log --------
    width = env.spec.action_width
    log = torch.zeros((K, N, width), dtype=torch.int32, device=env.device)
    prepare()
    env.kernel_timing(True)
    barrier()
    env.rollout_random(K, seed, actions_log=log)
    barrier()
    kt = env.kernel_times()
    fz = env.fused_rollout_times()
    fused = fz[1] > 0  # pin-less kinds: the whole rollout is one launch of rollout_small_kernel
    if fused:
        kt["step"] = (fz[0], fz[2])  # device time, step-equivalents
    stats2 = env.episode_stats(reset=True)
    # same again with the kernels strictly one after the other: the step kernel on its own
    prepare()
    env.reward_overlap(False)
    barrier()
    env.rollout_random(K, seed)
    barrier()
    kt_alone = env.kernel_times()
    fz = env.fused_rollout_times()
    if fused:
        kt_alone["step"] = (fz[0], fz[2])
    env.reward_overlap(True)
    env.kernel_timing(False)
    env.episode_stats(reset=True)

    # ---- pass 3: end to end through host buffers ------------------------------------------
    # The recorded actions of pass 2 are replayed from pinned host memory as flat Discrete
    # actions (what the reference's PPO loop sends through FlatteningActionWrapper*; the step
    # kernel decodes them), reward + done of every step are read back into pinned host memory.
    #   e2e_sync : pcb_step_host, one call and one host synchronisation per step
    #   e2e      : pcb_steps_host, HOST_ROWS rows per call -- every step still has its own
    #              action row copied up and its result rows copied down inside the timed
    #              region, but the copies overlap the neighbouring steps and the host waits
    #              once per call
    H_, W_ = env.spec.height, env.spec.width
    flat = (log[..., 0] * H_ + log[..., 1]) * W_ + log[..., 2] if width == 3 else log[..., 0] * W_ + log[..., 1]
    h_actions = torch.empty((K, N), dtype=torch.int32, pin_memory=True)
    h_actions.copy_(flat.int())
    del log, flat
    step_host = env.host_stepper(h_actions)  # packed pinned result block, pointers bound once
    d_np = step_host.done.numpy()
    checksum = 0.0

    def timed_from_same_state(body):
        # E2E_REPETITIONS passes of the same K steps, each from the state the recorded actions
        # belong to (prepare() is untimed); the wall-clock region of a pass covers the K steps
        # and the final synchronisation
        out = []
        for _ in range(E2E_REPETITIONS):
            prepare()
            barrier()
            t0 = time.perf_counter()
            body()
            torch.cuda.synchronize()
            out.append((time.perf_counter() - t0) * 1e3)
            barrier()
        return out

    def sync_pass():
        nonlocal checksum
        for t in range(K):
            step_host(t)
            checksum += float(d_np[0])

    sync_reps = timed_from_same_state(sync_pass)
    stats3 = env.episode_stats(reset=True)
    rows = min(HOST_ROWS, K)
    h_reward = torch.empty((K, N), dtype=torch.float64, pin_memory=True)
    h_done = torch.empty((K, N), dtype=torch.uint8, pin_memory=True)
    prepare()
    env.steps_host(h_actions[:rows], h_reward[:rows], h_done[:rows])  # staging allocated untimed

    def batch_pass():
        nonlocal checksum
        for r0 in range(0, K, rows):
            r1 = min(K, r0 + rows)
            env.steps_host(h_actions[r0:r1], h_reward[r0:r1], h_done[r0:r1])
            checksum += float(h_done[r1 - 1, 0])

    batch_reps = timed_from_same_state(batch_pass)
    stats4 = env.episode_stats(reset=True)

    # ---- reduce over ranks ------------------------------------------------------------------
    R, E = len(reps_ms), E2E_REPETITIONS
    all_max = ctx.max_over_ranks(reps_ms + sync_reps + batch_reps, env.device)
    reps_max = all_max[:R]
    sync_ms_max = sorted(all_max[R:R + E])[E // 2]     # median pass, each the max over ranks
    batch_ms_max = sorted(all_max[R + E:])[E // 2]
    env.close()
    del env, h_actions, h_reward, h_done
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms_max = sorted(reps_max)[len(reps_max) // 2]  # median repetition, each the max over ranks
    total_steps = float(N) * K * world
    value = total_steps / (ms_max * 1e-3)
    peak, peak_src = load_peaks()
    step_ms, step_n = kt["step"]
    term_ms, _ = kt["reward"]
    samp_ms, _ = kt["sample"]
    bytes_per_launch = ALGO_BYTES[name] * float(N)  # every env's observation is written every step
    avg_step_s = step_ms * 1e-3 / max(step_n, 1)
    achieved = bytes_per_launch / avg_step_s / 1e9
    alone_s = kt_alone["step"][0] * 1e-3 / max(kt_alone["step"][1], 1)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and N == TRAFFIC_ENVS.get(name):  # env counts the ncu captures were taken at
        with open(tpath) as f:
            traffic = json.load(f).get(f"{name}:step_kernel")
    e2e_path = ("pcb_steps_host (C ABI, PcbVecEnv.steps_host), %d action rows per call: each step's flat "
                "int32 action row is copied from pinned host memory (H2D) and its reward + done rows are "
                "copied back (D2H) inside the timed region, overlapped with the neighbouring steps; one "
                "host synchronisation per call; observations stay on device (DLPack); wall clock around the %d steps "
                "and the final synchronisation, median of %d passes" % (rows, K, E2E_REPETITIONS))
    out = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD_DESC[name], "envs_per_gpu": N, "policy": "uniform random over valid "
            "actions, sampled on device (counter-based RNG)", "autoreset": True,
            "instances": f"{N * 4} per GPU from the library's device generator (distribution of the reference's generate_instances)",
            "episode_phases": "staggered (1/%d of the envs finish every step)" % EP_LEN[name],
            "l2": "working set larger than L2: every step writes a fresh %.1f MB observation slice "
                  "(ring of %d slices)" % (N * (ALGO_BYTES[name]) / 1e6, ring),
            "timing": "CUDA events on the launching stream around %d steps (enqueued behind a ~0.1 ms spin "
                      "kernel so that host launch latency is not counted as device time), max over ranks; "
                      "median of %d back-to-back repetitions" % (K, REPETITIONS),
        },
        "repetitions_ms": reps_max,
        "e2e": {"value": total_steps / (batch_ms_max * 1e-3), "unit": "env-steps/s",
                "h2d_bytes_per_step": int(N * 4), "d2h_bytes_per_step": int(N * 9),
                "ms_per_step": batch_ms_max / K, "path": e2e_path},
        "e2e_sync": {"value": total_steps / (sync_ms_max * 1e-3), "unit": "env-steps/s",
                     "h2d_bytes_per_step": int(N * 4), "d2h_bytes_per_step": int(N * 9),
                     "ms_per_step": sync_ms_max / K,
                     "path": "pcb_step_host (C ABI, PcbVecEnv.host_stepper): one call per step -- H2D of the "
                             "action row, step, one packed D2H of reward + done, host synchronisation every step"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "rollout_small_kernel (all steps in one launch; per step-equivalent)" if fused else "step_kernel",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_env_step": ALGO_BYTES[name],
                     "envs_written_per_launch": float(N),
                     "avg_launch_us": avg_step_s * 1e6,
                     "alone": {"avg_launch_us": alone_s * 1e6,
                               "frac": bytes_per_launch / alone_s / 1e9 / peak,
                               "note": "step_kernel timed without the reward kernel running beside it"},
                     "whole_step_frac": value / world * ALGO_BYTES[name] / 1e9 / peak,
                     "kernel_share_ms": {"step": step_ms, "reward": term_ms, "sample": samp_ms}},
        "episodes": {"device_pass": stats["episodes"], "timed_kernel_pass": stats2["episodes"],
                     "e2e_sync_pass": stats3["episodes"], "e2e_pass": stats4["episodes"],
                     "mean_return": stats["sum_return"] / max(stats["episodes"], 1),
                     "mean_length": stats["sum_length"] / max(stats["episodes"], 1)},
        "wall_s_timed_region": wall,
    }
    if agg is not None:
        out["episode_stats_allreduce"] = {"ms": agg_ms, "episodes_all_ranks": agg["episodes"],
                                          "steps_all_ranks": agg["steps"],
                                          "note": "one NCCL all-reduce of 6 float64 per rollout, outside the timed region"}
    return out


def gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # An explicit stream: on the legacy default stream the overlapped reward kernels of the
    # beam reward types run in one of two regimes per call (194 or ~300 us per step, measured:
    # profiles/README.md), on a created stream they do not.
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = Ctx(rank, world, local_rank)
    name = args.workload
    out = measure(args, ctx, name, args.envs or DEFAULT_ENVS[name], main=True)
    others = []
    if name == WORKLOAD and not args.no_secondary:
        # the other BASELINE.json configurations at their BASELINE sizes, at this N too
        for sec in SECONDARY:
            r = measure(args, ctx, sec, DEFAULT_ENVS[sec], main=False)
            if r is not None:
                keep = ("value", "unit", "ms_per_step", "repetitions_ms", "e2e", "e2e_sync",
                        "gpu_launches", "clocks", "roofline", "n_gpus")
                others.append({"workload": r["config"]["workload"],
                               "envs_per_gpu": r["config"]["envs_per_gpu"],
                               "l2": r["config"]["l2"], **{k: r[k] for k in keep}})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    if others:
        out["other_workloads"] = others
    out["cpu_baseline"] = cpu_baseline(name, args.cpu_seconds) if (world == 1 and not args.no_cpu) else None
    print(json.dumps(out), flush=True)


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path on the host cores: the reference's own
    environment classes (pure Python, staged unmodified under baseline/_ref by build(); the
    oracle port only if nothing was staged), one process per core, random-valid policy.  A
    "step" is a bounded time slice of the rollout; W warm-up slices, K timed slices, one pool."""
    if rank != 0:
        return
    name = args.workload
    procs = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    slice_s = max(0.02, min(2.0, 45.0 / max(K + W, 1)))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_slices, [(name, w, K + W, slice_s) for w in range(procs)])
    wall = time.perf_counter() - t0
    value = sum(sum(s for s, _ in r[W:]) / sum(t for _, t in r[W:]) for r in res)
    total_steps = sum(s for r in res for s, _ in r[W:])
    out = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": slice_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[name], "policy": "uniform random over valid actions",
                   "note": "CPU multiprocess vector env; a step = a %.3f s slice of the rollout "
                           "in each of %d processes" % (slice_s, procs)},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": procs, "kind": cpu_kind(),
                         "sample": f"{K} timed slices of {slice_s:.3f} s x {procs} processes, config {name}, "
                                   f"{total_steps} env-steps, wall {wall:.0f} s; {cpu_what()}"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the BASELINE size of the workload)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the lines of the other BASELINE configs (B 4 096 envs, D, E 131 072 envs)")
    ap.add_argument("--no-fused", action="store_true",
                    help="pin-less kinds: per-step launches instead of the multi-step rollout kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
