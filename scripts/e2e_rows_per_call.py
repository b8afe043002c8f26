"""Wall time of one pcb_steps_host call against the rows per call (config C, 65 536 envs by default;
argv: workload envs): median of 9 calls each, and the least-squares fit  time = fixed + rows * per_row.
The fixed part is what a short call (the driver's --steps 20) pays on top of the steady-state rate."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import pcbgpu  # noqa: E402
from pcbgpu.configs import CONFIGS  # noqa: E402
from bench import stagger  # noqa: E402  (1 / EP_LEN of the envs finish every step, as in the bench)

NAME = sys.argv[1] if len(sys.argv) > 1 else "C"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = 128
env = pcbgpu.PcbVecEnv(CONFIGS[NAME], N, autoreset=True, obs_ring=4, bank_size=4 * N, seed=1,
                       generate_on_device=True)


def prep():
    # the state the recorded actions belong to (deterministic: same seeds, same trajectory)
    env.reset()
    stagger(env, NAME, 3)
    env.rollout_random(8, seed=3)


prep()
H_, W_, AW = env.spec.height, env.spec.width, env.spec.action_width
log = torch.zeros((K, N, AW), dtype=torch.int32, device="cuda")
env.rollout_random(K, seed=3, actions_log=log)
flat = ((log[..., 0] * H_ + log[..., 1]) * W_ + log[..., 2]) if AW == 3 else (log[..., 0] * W_ + log[..., 1])
h_act = torch.empty((K, N), dtype=torch.int32, pin_memory=True)
h_act.copy_(flat.int())
h_r = torch.empty((K, N), dtype=torch.float64, pin_memory=True)
h_d = torch.empty((K, N), dtype=torch.uint8, pin_memory=True)
env.steps_host(h_act[:K], h_r[:K], h_d[:K])  # staging
print('episodes finished per step in the replay:', float(h_d.float().sum()) / K)
rows_list = [1, 2, 4, 8, 16, 20, 32, 64, 128]
med = []
for rows in rows_list:
    ts = []
    for rep in range(9):
        prep()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        env.steps_host(h_act[:rows], h_r[:rows], h_d[:rows])
        ts.append((time.perf_counter() - t0) * 1e6)
    m = float(np.median(ts))
    med.append(m)
    print(f"rows {rows:4d}  call {m:9.1f} us  per row {m / rows:7.1f} us  min {min(ts):9.1f}", flush=True)
A = np.stack([np.ones(len(rows_list)), np.array(rows_list, float)], 1)
fixed, per = np.linalg.lstsq(A, np.array(med), rcond=None)[0]
print(f"fit: fixed {fixed:.1f} us per call + {per:.2f} us per row")
# device-only reference: rollout_random of the same length, one sync per call
for rows in (20, 128):
    ts = []
    for rep in range(9):
        prep()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        env.rollout_random(rows, seed=3)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    print(f"rollout_random {rows:4d} steps: {np.median(ts):9.1f} us wall, per step {np.median(ts) / rows:7.1f}")
