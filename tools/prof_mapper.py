#!/usr/bin/env python3
"""Profiling workload for the config-3 / config-4 kernels: MinimizeBatch over 16 384 ONT-like reads (the judged read
batch) and one b200_map_batch of 2 048 reads against the 4.6 Mbp reference. Run under ncu with a kernel filter."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from bioinfo1_b200 import capi

what = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = torch.device("cuda", 0)
ctx = capi.Context(0); L = capi.lib()
ref = synth.dna(1, 4_600_000)
if what in ("minimize", "both"):
    buf, off = synth.ont_reads(2, ref, n=16384)
    d_buf = torch.from_numpy(buf).to(dev)
    plan = C.c_void_p()
    capi.check(L.b200_min_plan_create(ctx.h, len(off) - 1, off.ctypes.data, 15, 5, None, C.byref(plan)))
    tot = int(L.b200_min_plan_tuples(plan))
    d_h = torch.empty(tot, dtype=torch.int32, device=dev); d_p = torch.empty(tot, dtype=torch.int32, device=dev)
    d_f = torch.empty(tot, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(3):
        capi.check(L.b200_min_plan_run(plan, d_buf.data_ptr(), d_h.data_ptr(), d_p.data_ptr(), d_f.data_ptr(), st.cuda_stream))
    torch.cuda.synchronize()
    if os.environ.get("PROF_TIME"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            capi.check(L.b200_min_plan_run(plan, d_buf.data_ptr(), d_h.data_ptr(), d_p.data_ptr(), d_f.data_ptr(), st.cuda_stream))
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        alg = int(off[-1]) + 9 * tot
        print(f"minimize 16384 reads: {ms:.4f} ms, {alg / ms / 1e6:.1f} GB/s algorithmic, {alg / ms / 1e6 / 6548.8:.3f} of the HBM copy peak")
if what in ("map", "both"):
    index = capi.Index(ctx, ref[:4_600_000].tobytes(), 15, 5, 0.001)
    rb, ro = synth.ont_reads(2, ref, n=2048)
    for _ in range(2):
        out, cig, coff = index.map_packed(rb, ro, True, 2, 1, -1, -1, True)
    torch.cuda.synchronize()
    print("mapped", int(out["mapped"].sum()))
