#!/usr/bin/env python3
"""Host-path tuning harness: b200_align_batch_packed on the config-2 batch for several pipeline chunk sizes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, seqgen
from bioinfo1_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
qb, qo, tb, to = seqgen.short_pairs(1000, n)
ctx = capi.Context(0); L = capi.lib()
hq, ht = torch.from_numpy(qb).pin_memory(), torch.from_numpy(tb).pin_memory()
cap = 64 * n + (1 << 20)
hs, hb = torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
hc, ho = torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(n + 1, dtype=torch.int64).pin_memory()
def step():
    capi.check(L.b200_align_batch_packed(ctx.h, n, hq.data_ptr(), qo.ctypes.data, ht.data_ptr(), to.ctypes.data, 0, 1, -1, -1,
                                         hs.data_ptr(), hb.data_ptr(), hc.data_ptr(), ho.data_ptr(), cap))
for chunk in (0, n // 4, 151552, n // 8, 98304, 75776, n // 16):
    ctx.set_option("chunk_pairs", chunk)
    for _ in range(3): step()
    ctx.set_option("profile", 1); ctx.set_option("reset_counters", 1)
    step()
    prof = {k: ctx.counter(k + "_ns") / 1e6 for k in ("fill", "walk", "emit", "other")}
    ctx.set_option("profile", 0)
    t0 = time.perf_counter()
    for _ in range(10): step()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print(f"chunk_pairs={chunk:8d}  e2e {ms:6.2f} ms  kernel ms {prof}")
