"""BASELINE config 4: the spatial pin env driving a PPO rollout collector end to end.

    python examples/ppo_spatial.py [--envs 4096] [--iters 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import __graft_entry__ as ge  # noqa: E402

pcbgpu = ge.build()
from pcbgpu.ppo import PPOTrainer  # noqa: E402
from ref_harness import CONFIGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--horizon", type=int, default=16)
args = ap.parse_args()
env = pcbgpu.PcbVecEnv(CONFIGS["D"], args.envs, autoreset=True, bank_size=args.envs * 8, seed=1)
trainer = PPOTrainer(env, horizon=args.horizon)
for i, rec in enumerate(trainer.train(args.iters)):
    print(f"iter {i}: loss {rec['loss']:.4f}  episodes {rec['episodes']:.0f}  "
          f"mean return {rec['mean_return']:.4f}")
