#!/usr/bin/env python3
"""HBM direction check for the roofline denominators: write-only (fill), read-only (sum) and copy rates of this GPU,
next to MEASURED_PEAKS.json's copy figure. The minimizer and direction-store passes are write-dominated."""
import json
import torch
dev = torch.device("cuda", 0)
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
ai = a.view(torch.int32); bi = b.view(torch.int32)
out = {"write_only_gbs": n / timeit(lambda: ai.fill_(7)) / 1e9,
       "memset_gbs": n / timeit(lambda: a.zero_()) / 1e9,
       "read_only_gbs": n / timeit(lambda: ai.sum()) / 1e9,
       "copy_gbs_read_plus_write": 2 * n / timeit(lambda: bi.copy_(ai)) / 1e9}
print(json.dumps(out))
