#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched placement environments on B200.

Metric (BASELINE.json): env-steps/sec, whole job, device-timed, on the random-
policy rollout { device random policy -> step -> terminal reward -> auto-reset }
of config C (dummy_env_rectangular_pin 32x32, 10 components, 5 nets, centroid
reward, 65 536 envs per GPU); weak scaling over GPUs, no collective on the step
path.  One JSON line on rank 0:
  value        device-timed, actions sampled on device, everything resident in HBM
  e2e          same rollout through pcb_step_host: actions from pinned host memory
               (H2D every step), reward + done read back (D2H every step)
  roofline     the step kernel: algorithmic bytes per launch / its CUDA-event time
  cpu_baseline the oracle port (numpy restatement of the reference) on the host
               cores as a multiprocess vector env, bounded sample
`--impl reference` times that CPU path alone with all host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "C"
ENVS_PER_GPU = 65536
OBS_RING = 4
REPETITIONS = 5
# SURVEY.md section 8d / BASELINE.md section 5: algorithmic bytes per env-step (compact contract)
ALGO_BYTES = {"A": 161, "B": 1049, "C": 7699, "D": 15775, "E": 28821,
              # the reference's shipped 10x10 pin config (agent/config/rectangle_pin_model.json):
              # grid 100 + mask 400 + features 100 + placement 5 + pins 160 + 40 + 21 + 2 * 192 state
              "REF_pin": 1210, "C_both": 7699, "C_beam": 7699,
              # rectangle_model.json: grid 100 + mask 200 + features 400 + masks 40 + 21 + 2 * 240 state
              "REF_rect": 1241}
EP_LEN = {"A": 4, "B": 6, "C": 10, "D": 10, "E": 32, "REF_pin": 5, "C_both": 10, "C_beam": 10,
          "REF_rect": 8}
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
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------
# CPU baseline: the oracle port as a multiprocess vector env
# ----------------------------------------------------------------------------
def _cpu_worker(args):
    name, wid, budget = args
    import random

    import numpy as np

    from ref_harness import CONFIGS, make_oracle

    np.random.seed(wid)
    random.seed(wid)
    rng = np.random.default_rng(wid)
    env = make_oracle(CONFIGS[name])
    steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget:
        env.reset()
        done = False
        while not done:
            valid = np.flatnonzero(env.action_mask)
            a = np.unravel_index(valid[rng.integers(0, len(valid))], env.action_mask.shape)
            _, _, done, _ = env.step(tuple(int(v) for v in a))
            steps += 1
    return steps, time.perf_counter() - t0


def cpu_baseline(name, budget_s, procs=None):
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(name, w, budget_s) for w in range(procs)])
    wall = time.perf_counter() - t0
    rate = sum(s / t for s, t in res)
    return {"value": rate, "unit": "env-steps/s", "cores": procs, "kind": "port",
            "sample": f"config {name}, random valid policy, {procs} processes x {budget_s:.0f} s "
                      f"({sum(s for s, _ in res)} env-steps, wall {wall:.1f} s); oracle/ numpy port "
                      "of the reference envs (reference itself is Python and absent on the GPU box)"}


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


def gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.build()
    import pcbgpu
    from ref_harness import CONFIGS

    name = args.workload
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = args.envs
    cfg = CONFIGS[name]
    # The observation ring must not fit in L2 (126 MB): small workloads get more slices.
    # The pin-less kinds run the whole rollout in one launch whose CTAs are not in lock-step:
    # a short ring would be overwritten inside L2 (measured: 12 % of the stores reached DRAM
    # with 5 slices), so there the ring keeps every observation of the timed rollout -- what a
    # rollout collector needs anyway -- as long as that stays below 48 GB.
    slice_bytes = N * ALGO_BYTES[name]
    ring = max(OBS_RING, -(-300_000_000 // slice_bytes))
    if cfg["type"] in ("square", "rectangle") and not args.no_fused:
        ring = max(ring, min(args.steps + 1, 48_000_000_000 // slice_bytes))
    env = pcbgpu.PcbVecEnv(cfg, N, device=local_rank, autoreset=True, obs_ring=ring,
                           bank_size=N * 4, seed=1000 + rank, generate_on_device=True)
    if args.no_fused:
        env.fused_rollout(False)
    seed = 7 + rank
    K, W = args.steps, args.warmup
    # An explicit stream: on the legacy default stream the overlapped reward kernels of the
    # beam reward types run in one of two regimes per call (194 or ~300 us per step, measured:
    # profiles/README.md), on a created stream they do not.
    torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def prepare():
        env.reset()
        stagger(env, name, seed)
        env.rollout_random(W, seed)
        env.episode_stats(reset=True)

    # ---- pass 1: device-timed, device policy (value) -------------------------
    prepare()
    width = env.spec.action_width
    log = torch.zeros((K, N, width), dtype=torch.int32, device=env.device)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    l0 = env.launch_count
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    env.rollout_random(K, seed)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = env.launch_count - l0
    reps_ms = [e0.elapsed_time(e1)]
    # C1, the only collective of the design: the per-rollout all-reduce of the episode
    # statistics (NCCL over NVLink); off the step path, its cost is reported beside it
    agg, agg_ms = None, None
    if world > 1:
        pcbgpu.allreduce_episode_stats(env)  # first call sets up the communicator
        torch.cuda.synchronize()
        t_a = time.perf_counter()
        agg = pcbgpu.allreduce_episode_stats(env)
        torch.cuda.synchronize()
        agg_ms = (time.perf_counter() - t_a) * 1e3
    stats = env.episode_stats(reset=True)
    # median of 5 repetitions of the same K-step region (the rollout just goes on)
    for _ in range(REPETITIONS - 1):
        barrier()
        e0.record(stream)
        env.rollout_random(K, seed)
        e1.record(stream)
        barrier()
        reps_ms.append(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    env.episode_stats(reset=True)

    # ---- pass 2: same rollout with per-kernel CUDA events (roofline) + action log ----
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

    # ---- pass 3: end to end through host buffers (e2e) ---------------------------
    import numpy as np

    # the reference's PPO loop sends flat Discrete actions (FlatteningActionWrapper*); the
    # step kernel decodes them, so the host buffer holds one int32 per env
    H_, W_ = env.spec.height, env.spec.width
    if width == 3:
        flat = (log[..., 0] * H_ + log[..., 1]) * W_ + log[..., 2]
    else:
        flat = log[..., 0] * W_ + log[..., 1]
    h_actions = torch.empty((K, N), dtype=torch.int32, pin_memory=True)
    h_actions.copy_(flat.int())
    del log, flat
    h_reward = torch.empty((N,), dtype=torch.float64, pin_memory=True)
    h_done = torch.empty((N,), dtype=torch.uint8, pin_memory=True)
    d_np = h_done.numpy()
    prepare()
    step_host = env.host_stepper(h_actions, h_reward, h_done)  # pointers bound once: ~2 us per call
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for t in range(K):
        step_host(t)
        checksum += float(d_np[0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    stats3 = env.episode_stats(reset=True)

    # ---- reduce over ranks ------------------------------------------------------
    vals = torch.tensor(reps_ms + [e2e_s * 1e3], dtype=torch.float64, device=env.device)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    *reps_max, e2e_ms_max = vals.tolist()
    ms_max = sorted(reps_max)[len(reps_max) // 2]  # median repetition, each the max over ranks
    total_steps = float(N) * K * world
    value = total_steps / (ms_max * 1e-3)
    e2e_value = total_steps / (e2e_ms_max * 1e-3)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = load_peaks()
    step_ms, step_n = kt["step"]
    term_ms, term_n = kt["reward"]
    samp_ms, samp_n = kt["sample"]
    done_frac = stats2["episodes"] / (float(N) * K)
    # with the fused terminal path (centroid reward) the step kernel also writes the first
    # observation of every restarting env; otherwise terminal_kernel writes those
    written = float(N)  # step_kernel writes every env's observation (restarting envs included)
    bytes_per_launch = ALGO_BYTES[name] * written
    avg_step_s = step_ms * 1e-3 / max(step_n, 1)
    achieved = bytes_per_launch / avg_step_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and N == ENVS_PER_GPU:  # the captures were taken at 65 536 envs
        with open(tpath) as f:
            traffic = json.load(f).get(f"{name}:step_kernel")
    out = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD_DESC[name], "envs_per_gpu": N, "policy": "uniform random over valid "
            "actions, sampled on device (Philox)", "autoreset": True,
            "instances": f"{N * 4} per GPU from the library's device generator (distribution of the reference's generate_instances)",
            "episode_phases": "staggered (1/%d of the envs finish every step)" % EP_LEN[name],
            "l2": "working set larger than L2: every step writes a fresh %.1f MB observation slice "
                  "(ring of %d slices)" % (N * (ALGO_BYTES[name]) / 1e6, ring),
            "timing": "CUDA events on the launching stream around %d steps, max over ranks; "
                      "median of %d back-to-back repetitions" % (K, REPETITIONS),
        },
        "repetitions_ms": reps_max,
        "e2e": {"value": e2e_value, "unit": "env-steps/s",
                "h2d_bytes_per_step": int(N * 4), "d2h_bytes_per_step": int(N * 9),
                "path": "pcb_step_host (C ABI, called through PcbVecEnv.host_stepper): pinned host flat actions "
                        "(int32 per env) -> H2D -> step -> D2H reward+done, sync every step; observations stay "
                        "on device (DLPack)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "rollout_small_kernel (all steps in one launch; per step-equivalent)" if fused else "step_kernel",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_env_step": ALGO_BYTES[name],
                     "envs_written_per_launch": written,
                     "avg_launch_us": avg_step_s * 1e6,
                     "alone": {"avg_launch_us": kt_alone["step"][0] * 1e3 / max(kt_alone["step"][1], 1),
                               "frac": bytes_per_launch / (kt_alone["step"][0] * 1e-3 / max(kt_alone["step"][1], 1)) / 1e9 / peak,
                               "note": "step_kernel timed without the reward kernel running beside it"},
                     "whole_step_frac": value / world * ALGO_BYTES[name] / 1e9 / peak,
                     "kernel_share_ms": {"step": step_ms, "reward": term_ms, "sample": samp_ms}},
        "episodes": {"device_pass": stats["episodes"], "timed_kernel_pass": stats2["episodes"],
                     "e2e_pass": stats3["episodes"],
                     "mean_return": stats["sum_return"] / max(stats["episodes"], 1),
                     "mean_length": stats["sum_length"] / max(stats["episodes"], 1)},
        "wall_s_timed_region": wall,
    }
    if agg is not None:
        out["episode_stats_allreduce"] = {"ms": agg_ms, "episodes_all_ranks": agg["episodes"],
                                          "steps_all_ranks": agg["steps"],
                                          "note": "one NCCL all-reduce of 6 float64 per rollout, outside the timed region"}
    if world == 1 and name == WORKLOAD and not args.no_secondary:
        env.close()
        del env
        torch.cuda.empty_cache()
        out["other_workloads"] = [secondary(args, "B", 4096), secondary(args, "D", ENVS_PER_GPU)]
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(name, args.cpu_seconds)
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)


def secondary(args, name, N):
    """Device-timed value of another BASELINE.json configuration (B = configs[1] at its 4 096
    envs, D = configs[3]); same method as the main line, one repetition, no e2e / roofline."""
    import torch

    import pcbgpu
    from ref_harness import CONFIGS

    K, W = args.steps, args.warmup
    slice_bytes = N * ALGO_BYTES[name]
    ring = max(OBS_RING, -(-300_000_000 // slice_bytes))
    if CONFIGS[name]["type"] in ("square", "rectangle"):
        ring = max(ring, min(K + 1, 48_000_000_000 // slice_bytes))
    env = pcbgpu.PcbVecEnv(CONFIGS[name], N, device=torch.cuda.current_device(), autoreset=True,
                           obs_ring=ring, bank_size=N * 4, seed=1000, generate_on_device=True)
    env.reset()
    stagger(env, name, 7)
    env.rollout_random(W, 7)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = env.launch_count
    e0.record(stream)
    env.rollout_random(K, 7)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    peak, _ = load_peaks()
    value = float(N) * K / (ms * 1e-3)
    res = {"workload": WORKLOAD_DESC[name], "envs_per_gpu": N, "value": value, "unit": "env-steps/s",
           "ms_per_step": ms / K, "gpu_launches": env.launch_count - l0, "obs_ring": ring,
           "whole_step_frac": value * ALGO_BYTES[name] / 1e9 / peak}
    env.close()
    return res


def _cpu_slices(args):
    """One worker of the reference arm: (warmup + steps) consecutive time slices of the
    random-policy loop; returns (env-steps, seconds) of every slice."""
    name, wid, n_slices, slice_s = args
    import random

    import numpy as np

    from ref_harness import CONFIGS, make_oracle

    np.random.seed(wid)
    random.seed(wid)
    rng = np.random.default_rng(wid)
    env = make_oracle(CONFIGS[name])
    out = []
    env.reset()
    done = False
    for _ in range(n_slices):
        steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < slice_s:
            if done:
                env.reset()
                done = False
            valid = np.flatnonzero(env.action_mask)
            a = np.unravel_index(valid[rng.integers(0, len(valid))], env.action_mask.shape)
            _, _, done, _ = env.step(tuple(int(v) for v in a))
            steps += 1
        out.append((steps, time.perf_counter() - t0))
    return out


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path on the host cores.  The reference
    is pure Python and is not present on the GPU box, so this times the oracle port
    (oracle/: same numpy/scipy operations per step) with one process per core.  A "step"
    is a bounded time slice of the rollout; W warm-up slices, K timed slices, one pool."""
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
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": procs, "kind": "port",
                         "sample": f"{K} timed slices of {slice_s:.3f} s x {procs} processes, config {name}, "
                                   f"{total_steps} env-steps, wall {wall:.0f} s; oracle/ numpy port of the "
                                   "reference envs (the reference is Python and absent on the GPU box)"},
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
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the short device-timed lines of configs B (4 096 envs) and D")
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
