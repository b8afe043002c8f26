"""C_both: time of consecutive rollout_random(K) calls, overlapped vs serial reward kernels."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import pcbgpu  # noqa: E402
from bench import stagger  # noqa: E402
from pcbgpu.configs import CONFIGS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C_both"
N, K = 65536, 128
env = pcbgpu.PcbVecEnv(CONFIGS[name], N, autoreset=True, obs_ring=4, bank_size=4 * N, seed=1000,
                       generate_on_device=True)
if os.environ.get("USE_STREAM"):
    torch.cuda.set_stream(torch.cuda.Stream())
st = torch.cuda.current_stream()


def reps(label, n, k):
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        env.rollout_random(k, 7)
        b.record(st)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / k * 1e3)
    print(label, " ".join("%.0f" % v for v in out), "us/step", flush=True)


for overlap in (True,):
    env.reward_overlap(overlap)
    env.reset()
    stagger(env, name, 7)
    env.rollout_random(16, 7)
    reps(f"overlap={overlap} K=128 x8:", 8, 128)
    reps(f"overlap={overlap} K=32 x12:", 12, 32)
