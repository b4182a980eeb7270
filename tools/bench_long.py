#!/usr/bin/env python3
"""Kernel-tuning harness for the long-pair classes (BASELINE configs 4/5 shapes): device-resident
plan_run over ONT-like pairs, per-kind kernel times from the context's event brackets, optional
parity check of a few pairs against the CPU oracle.  Not the judged bench (that is bench.py)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2048)
    ap.add_argument("--mean", type=int, default=8000)
    ap.add_argument("--fixed", type=int, default=0)
    ap.add_argument("--type", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=2)
    ap.add_argument("--no-cigar", action="store_true")
    ap.add_argument("--force-generic", action="store_true")
    args = ap.parse_args()
    import torch
    import seqgen
    from bioinfo1_b200 import capi
    from cpu_checkers import load_oracle

    t0 = time.time()
    base_q, base_t = seqgen.ont_like_pairs(5, min(args.pairs, 256), mean_len=args.mean, fixed=args.fixed or None)
    reps = (args.pairs + len(base_q) - 1) // len(base_q)
    qs = (base_q * reps)[:args.pairs]
    ts = (base_t * reps)[:args.pairs]
    qb, qo = seqgen.pack_arrays(qs)
    tb, to = seqgen.pack_arrays(ts)
    gen_s = time.time() - t0
    ctx = capi.Context(0)
    if args.force_generic:
        ctx.set_option("force_generic", 1)
    L = capi.lib()
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(qb).to(dev)
    d_t = torch.from_numpy(tb).to(dev)
    n = args.pairs
    want = 0 if args.no_cigar else 1
    plan = C.c_void_p()
    capi.check(L.b200_align_plan_create(ctx.h, n, qo.ctypes.data, to.ctypes.data, args.type, 1, -1, -1, want, C.byref(plan)))
    cells = int(L.b200_align_plan_cells(plan))
    cap = int(L.b200_align_plan_cigar_bound(plan)) if want else 0
    d_score = torch.empty(n, dtype=torch.int32, device=dev)
    d_tb = torch.empty(n, dtype=torch.int32, device=dev)
    d_cig = torch.empty(max(cap, 16), dtype=torch.uint8, device=dev)
    d_coff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()

    def step():
        capi.check(L.b200_align_plan_run(plan, d_q.data_ptr(), d_t.data_ptr(), d_score.data_ptr(), d_tb.data_ptr(),
                                         d_cig.data_ptr() if want else None, d_coff.data_ptr() if want else None, cap,
                                         st.cuda_stream))
    step()
    torch.cuda.synchronize()
    ctx.set_option("profile", 1)
    ctx.set_option("reset_counters", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    out = {"pairs": n, "cells": cells, "type": args.type, "gen_s": gen_s, "ms_per_step": ms, "gcups": cells / ms / 1e6,
           "fill_ms": ctx.counter("fill_ns") / args.steps / 1e6, "walk_ms": ctx.counter("walk_ns") / args.steps / 1e6,
           "emit_ms": ctx.counter("emit_ns") / args.steps / 1e6, "other_ms": ctx.counter("other_ns") / args.steps / 1e6}
    out["fill_gcups"] = cells / max(out["fill_ms"], 1e-9) / 1e6
    if args.check:
        oracle = load_oracle()
        sc = d_score.cpu().numpy(); tbg = d_tb.cpu().numpy()
        cig = d_cig.cpu().numpy(); coff = d_coff.cpu().numpy()
        ok = True
        for k in range(min(args.check, n)):
            exp = oracle.align(qs[k].tobytes(), ts[k].tobytes(), args.type, 1, -1, -1, bool(want))
            got = (int(sc[k]), int(tbg[k]) & 0xffffffff, cig[int(coff[k]):int(coff[k + 1])].tobytes() if want else None)
            ok &= (got == exp)
        out["parity_checked"] = min(args.check, n)
        out["parity_ok"] = bool(ok)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
