"""BASELINE config 4: the spatial pin env driving a PPO rollout collector end to end.

    python examples/ppo_spatial.py [--envs 4096] [--iters 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import __graft_entry__ as ge  # noqa: E402

pcbgpu = ge.build()
from pcbgpu.ppo import PPOTrainer  # noqa: E402
from pcbgpu.configs import CONFIGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--horizon", type=int, default=16)
args = ap.parse_args()
env = pcbgpu.PcbVecEnv(CONFIGS["D"], args.envs, autoreset=True, bank_size=args.envs * 8, seed=1)
import time  # noqa: E402

import torch  # noqa: E402

trainer = PPOTrainer(env, horizon=args.horizon)
for i in range(args.iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    batch = trainer.collect()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss = trainer.update(batch)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = env.episode_stats(reset=True)
    n = args.envs * args.horizon
    print(f"iter {i}: loss {loss:.4f}  episodes {st['episodes']:.0f}  "
          f"mean return {st['sum_return'] / max(st['episodes'], 1.0):.4f}  "
          f"collect {n / (t1 - t0):.3e} env-steps/s (policy forward + env step)  "
          f"update {t2 - t1:.3f} s")
# the env's share of a collection step: the same steps with the actions replayed, no policy
acts = batch["actions"].int()
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(args.horizon):
    env.step(acts[t])
torch.cuda.synchronize()
print(f"env only: {args.envs * args.horizon / (time.perf_counter() - t0):.3e} env-steps/s "
      f"(pcb_step per step, flat actions decoded in the kernel)")
