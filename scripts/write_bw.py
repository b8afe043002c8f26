"""Microbench: pure-write vs copy HBM bandwidth on this GPU (context for the roofline:
the step kernel is ~93 % writes, MEASURED_PEAKS.json's peak is a copy = read + write)."""
import json
import torch

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3

out = {}
for mb in (505, 1024, 2048):
    n = mb * 1024 * 1024
    x = torch.empty(n, dtype=torch.uint8, device="cuda")
    y = torch.empty(n, dtype=torch.uint8, device="cuda")
    xi = x.view(torch.int32)
    out[f"fill_{mb}MB_GBs"] = n / timeit(lambda: xi.fill_(1)) / 1e9
    out[f"zero_{mb}MB_GBs"] = n / timeit(lambda: x.zero_()) / 1e9
    out[f"copy_{mb}MB_GBs_rw"] = 2 * n / timeit(lambda: y.copy_(x)) / 1e9
    out[f"read_sum_{mb}MB_GBs"] = n / timeit(lambda: xi.sum()) / 1e9
print(json.dumps(out, indent=1))
