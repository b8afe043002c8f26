"""Where the end-to-end step time goes (default config C, 65 536 envs; argv: workload envs): device step with a sync per
step, + H2D of the actions, + D2H of reward/done, and the full pcb_step_host."""
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
from pcbgpu.configs import CONFIGS  # noqa: E402

if os.environ.get("USE_STREAM"):
    torch.cuda.set_stream(torch.cuda.Stream())
NAME = sys.argv[1] if len(sys.argv) > 1 else "C"
N, K = int(sys.argv[2]) if len(sys.argv) > 2 else 65536, 192
env = pcbgpu.PcbVecEnv(CONFIGS[NAME], N, autoreset=True, obs_ring=4, bank_size=4 * N, seed=1,
                       generate_on_device=True)
env.reset()
H_, W_, AW = env.spec.height, env.spec.width, env.spec.action_width
log = torch.zeros((K, N, AW), dtype=torch.int32, device="cuda")
env.rollout_random(K, seed=3, actions_log=log)
if AW == 3:
    flat = ((log[..., 0] * H_ + log[..., 1]) * W_ + log[..., 2]).int().contiguous()
else:
    flat = (log[..., 0] * W_ + log[..., 1]).int().contiguous()
h_act = torch.empty((K, N), dtype=torch.int32, pin_memory=True)
h_act.copy_(flat)
h_r = torch.empty((N,), dtype=torch.float64, pin_memory=True)
h_d = torch.empty((N,), dtype=torch.uint8, pin_memory=True)
d_act = torch.empty((N,), dtype=torch.int32, device="cuda")


def run(label, body):
    env.reset()
    env.rollout_random(16, seed=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(K):
        body(t)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e6
    print(f"{label:52s} {dt:8.1f} us/step  {N / dt * 1e6:.3e} env-steps/s", flush=True)


def dev_only(t):
    env.step_async(flat[t])
    torch.cuda.synchronize()


def with_h2d(t):
    d_act.copy_(h_act[t], non_blocking=True)
    env.step_async(d_act)
    torch.cuda.synchronize()


def with_both(t):
    d_act.copy_(h_act[t], non_blocking=True)
    env.step_async(d_act)
    r = env.results()
    h_r.copy_(r["reward"], non_blocking=True)
    h_d.copy_(r["done"], non_blocking=True)
    torch.cuda.synchronize()


a_np, r_np, d_np = h_act.numpy(), h_r.numpy(), h_d.numpy()


def step_host(t):
    env.step_host(a_np[t], r_np, d_np)


def no_sync(t):
    env.step_async(flat[t])


if os.environ.get("NO_FUSED"):
    env.fused_rollout(False)
run("pcb_step, device actions, no sync", no_sync)
run("pcb_step, device actions, sync every step", dev_only)
run("+ H2D of the actions (torch copy)", with_h2d)
run("+ D2H of reward and done (torch copies)", with_both)
run("pcb_step_host, two result buffers (two D2H copies)", step_host)
stepper = env.host_stepper(h_act, h_r, h_d)
run("pcb_step_host through host_stepper (pointers bound once)", stepper)
packed = env.host_stepper(h_act)
run("pcb_step_host, packed result block (one D2H copy)", packed)
# pcb_steps_host: ROWS action rows per call, copies overlapped with the neighbouring steps
ROWS = 32
h_R = torch.empty((K, N), dtype=torch.float64, pin_memory=True)
h_D = torch.empty((K, N), dtype=torch.uint8, pin_memory=True)
env.steps_host(h_act[:ROWS], h_R[:ROWS], h_D[:ROWS])


def rows_host(t):
    if t % ROWS == 0:
        env.steps_host(h_act[t:t + ROWS], h_R[t:t + ROWS], h_D[t:t + ROWS])


run("pcb_steps_host, %d rows per call" % ROWS, rows_host)
