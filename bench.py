#!/usr/bin/env python3
"""bench.py -- the headline measurement: batched pairwise alignment GCUPS (score + CIGAR).

Workload (BASELINE.json configs[1]): 1 048 576 pairs of 150 bp x 150 bp, global NW, scores
1/-1/-1, CIGAR + target_begin produced, synthetic data (tests/seqgen.short_pairs). One "step"
is one pass of the whole hot path (classify -> DP fill -> traceback walk -> CIGAR emit) over the
batch. With N GPUs every rank owns its own batch of that size (weak scaling, no collective on
the data path); `value` = cells of all ranks / max-over-ranks time.

  value      inputs resident in HBM, device-resident C-ABI (b200_align_plan_run), CUDA events
  e2e        same batch through the host-buffer C-ABI (b200_align_batch_packed): pinned host
             buffers in, host arrays out, H2D/D2H inside the timed region (wall clock + sync)
  roofline   the DP fill kernel against the measured integer-ALU issue peak
  cpu_baseline / --impl reference: the reference's own CPU Align on the host cores

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OPS_PER_CELL = 9  # 1 cmp + 1 sel + 3 add + 2 x (cmp + sel) for global / semiGlobal
TYPE_NAMES = {0: "global", 1: "local", 2: "semiGlobal"}


def load_peaks():
    out = {"hbm_gbs": 6650.0, "source": "fallback"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            out.update(json.load(open(p)))
            out["source"] = "measured"
        except Exception:
            pass
    # integer-ALU issue peak measured on this pool by tools/ubench (profiles/int_peak_r01.json):
    # VIMNMX/VIADDMNMX/LOP3 = 64 lanes/clk/SM -> 18.4 T lane-ops/s at 1965 MHz
    ip = os.path.join(ROOT, "profiles", "int_peak_r01.json")
    out["int_tops"] = 18.4
    out["int_source"] = "nominal 148 SM x 64 lanes x 1.965 GHz"
    if os.path.exists(ip):
        try:
            r = json.load(open(ip))["results"]
            out["int_tops"] = float(r["VIADDMNMX.s32"]["Tops"])
            out["int_source"] = "measured (tools/ubench, VIADDMNMX.s32 alu-pipe issue rate)"
        except Exception:
            pass
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.t_begin = None   # wall-clock window of interest (set by mark_begin / stop)

    def mark_begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        import datetime
        t_end = datetime.datetime.now()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            if self.t_begin is not None:   # keep only samples taken while the GPU was under our load
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                    if ts < self.t_begin or ts > t_end:
                        continue
                except ValueError:
                    pass
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_arm(args, qb, qo, tb, to, typ, threads, budget_s, kind_pref="reference"):
    """Times the reference's CPU Align (oracle/_ref when present, else the C port) on a bounded
    sample of the same batch; returns (gcups, dict)."""
    from cpu_checkers import load_oracle, load_ref
    chk = load_ref() if kind_pref == "reference" else None
    if chk is None:
        chk = load_oracle()
    n_total = len(qo) - 1
    # calibrate on a small slice, then size the sample for ~budget_s seconds per step
    t0 = time.perf_counter()
    chk.align_batch(qb, qo[:257], tb, to[:257], typ)
    per_pair = (time.perf_counter() - t0) / 256
    n = int(min(n_total, max(threads * 64, budget_s / per_pair * threads)))
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)

    def one_step():
        out = [None] * threads

        def work(k):
            a, b = int(bounds[k]), int(bounds[k + 1])
            if b > a:
                out[k] = chk.align_batch(qb, qo[a:b + 1], tb, to[a:b + 1], typ)
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    cells = float(np.sum((qo[1:n + 1] - qo[:n]).astype(np.float64) * (to[1:n + 1] - to[:n]).astype(np.float64)))
    return chk, n, cells, one_step


def extra_workloads(ctx, capi, torch, dev, peaks):
    """Secondary, device-resident measurements on rank 0 (not the headline): the long-pair kernels on
    ONT-like pairs (BASELINE configs 4/5 shapes, bounded counts) and MinimizeBatch (config 3 shape)."""
    import ctypes as C
    import seqgen
    L = capi.lib()
    res = {}
    st = torch.cuda.current_stream()

    def run_align(tag, qs, ts, typ, steps=3):
        qb, qo = seqgen.pack_arrays(qs)
        tb, to = seqgen.pack_arrays(ts)
        n = len(qs)
        d_q, d_t = torch.from_numpy(qb).to(dev), torch.from_numpy(tb).to(dev)
        plan = C.c_void_p()
        capi.check(L.b200_align_plan_create(ctx.h, n, qo.ctypes.data, to.ctypes.data, typ, 1, -1, -1, 1, C.byref(plan)))
        cells = int(L.b200_align_plan_cells(plan))
        cap = int(L.b200_align_plan_cigar_bound(plan))
        d_s = torch.empty(n, dtype=torch.int32, device=dev)
        d_b = torch.empty(n, dtype=torch.int32, device=dev)
        d_c = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_o = torch.empty(n + 1, dtype=torch.int64, device=dev)

        def step():
            capi.check(L.b200_align_plan_run(plan, d_q.data_ptr(), d_t.data_ptr(), d_s.data_ptr(), d_b.data_ptr(),
                                             d_c.data_ptr(), d_o.data_ptr(), cap, st.cuda_stream))
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ctx.set_option("profile", 1); ctx.set_option("reset_counters", 1)
        step()
        torch.cuda.synchronize()
        fill_ms = ctx.counter("fill_ns") / 1e6
        walk_ms = ctx.counter("walk_ns") / 1e6
        ctx.set_option("profile", 0)
        L.b200_align_plan_destroy(plan)
        ops = 11 if typ == 1 else 9
        res[tag] = {"pairs": n, "cells": cells, "alignment": TYPE_NAMES[typ], "gcups": cells / ms / 1e6, "ms_per_step": ms,
                    "fill_ms": fill_ms, "walk_ms": walk_ms, "fill_gcups": cells / fill_ms / 1e6,
                    "roofline_frac_int_alu": cells * ops / (fill_ms * 1e-3) / 1e12 / peaks["int_tops"],
                    "ops_per_cell": ops}

    base_q, base_t = seqgen.ont_like_pairs(4242, 256, mean_len=8000)
    run_align("semiglobal_ont8kb_2048pairs", (base_q * 8), (base_t * 8), 2)
    fq, ft = seqgen.ont_like_pairs(4343, 64, fixed=10000)
    run_align("local_10kbx10kb_512pairs", fq * 8, ft * 8, 1, steps=2)

    # end-to-end mapping (BASELINE config 4 shape, bounded read count): 4.6 Mbp random reference, ONT-like reads
    # drawn from both strands at ~12 % indel-heavy error; index build timed separately from the per-read pipeline
    rng = np.random.default_rng(1)
    ref = seqgen.random_dna(rng, 4_600_000)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    mreads = []
    for i in range(512):
        Lr = int(np.clip(rng.lognormal(np.log(8000) - 0.125, 0.5), 1000, 40000))
        s0 = int(rng.integers(0, len(ref) - Lr))
        q = seqgen.mutate(rng, ref[s0:s0 + Lr], sub=0.024, ins=0.048, dele=0.048).tobytes()
        mreads.append(q.translate(comp)[::-1] if i % 2 else q)
    mreads = mreads * 4
    t0 = time.perf_counter()
    index = capi.Index(ctx, ref.tobytes(), 15, 5, 0.001)
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0          # first build of the process: module loading and first allocations included
    index.close()
    t0 = time.perf_counter()
    index = capi.Index(ctx, ref.tobytes(), 15, 5, 0.001)
    torch.cuda.synchronize()
    t_index_warm = time.perf_counter() - t0
    index.map_batch(mreads, True, 2, 1, -1, -1, True)   # warm-up at full size: the context's scratch buffers grow once
    t0 = time.perf_counter()
    mres, _ = index.map_batch(mreads, True, 2, 1, -1, -1, True)
    torch.cuda.synchronize()
    t_map = time.perf_counter() - t0
    res["map_2048_ont_reads_4.6Mbp_ref"] = {
        "reads": len(mreads), "bases": int(sum(len(r) for r in mreads)), "mapped": int(mres["mapped"].sum()),
        "index_build_s": t_index_warm, "index_build_first_call_s": t_index, "map_s": t_map, "mapped_reads_per_s": float(mres["mapped"].sum()) / t_map,
        "note": "host buffers in, PAF fields + CIGAR out (b200_map_batch), semiGlobal, k=15 w=5 f=0.001"}
    index.close()

    # MinimizeBatch, k=15 w=5: the reference (both strands) and ONT-like reads
    reads, _ = seqgen.ont_like_pairs(2, 256, mean_len=8000)
    reads = reads * 64
    for tag, seqs in (("minimize_ref_4.6Mbp_x2", [ref, ref[::-1].copy()]), ("minimize_16384_ont_reads", reads)):
        buf, off = seqgen.pack_arrays(seqs)
        d_buf = torch.from_numpy(buf).to(dev)
        plan = C.c_void_p()
        capi.check(L.b200_min_plan_create(ctx.h, len(seqs), off.ctypes.data, 15, 5, None, C.byref(plan)))
        tot = int(L.b200_min_plan_tuples(plan))
        d_h = torch.empty(tot, dtype=torch.int32, device=dev)
        d_p = torch.empty(tot, dtype=torch.int32, device=dev)
        d_f = torch.empty(tot, dtype=torch.uint8, device=dev)

        def mstep():
            capi.check(L.b200_min_plan_run(plan, d_buf.data_ptr(), d_h.data_ptr(), d_p.data_ptr(), d_f.data_ptr(),
                                           st.cuda_stream))
        for _ in range(3):
            mstep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            mstep()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        bases = int(off[-1])
        alg_bytes = bases + 9 * tot
        L.b200_min_plan_destroy(plan)
        res[tag] = {"bases": bases, "tuples": tot, "ms": ms, "gbases_per_s": bases / ms / 1e6,
                    "hbm_gbs": alg_bytes / ms / 1e6, "roofline_frac_hbm": alg_bytes / ms / 1e6 / peaks["hbm_gbs"],
                    "algorithmic_bytes": alg_bytes}
    return res


_OUT = sys.stdout


def emit(obj):
    _OUT.write(json.dumps(obj) + "\n")
    _OUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--length", type=int, default=150)
    ap.add_argument("--type", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="skip the e2e and CPU arms (kernel tuning runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (configs 3/4/5 shapes)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner does) goes to stderr
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    workload = (f"{args.pairs} pairs x {args.length}x{args.length} bp, {TYPE_NAMES[args.type]} NW, match 1 "
                f"mismatch -1 gap -1, score+CIGAR+target_begin (BASELINE.json configs[1])")
    config = {"workload": workload, "pairs_per_gpu": args.pairs, "query_len": args.length, "target_len": args.length,
              "alignment": TYPE_NAMES[args.type], "scores": [1, -1, -1], "cigar": True,
              "l2_policy": "inputs (300 MB ASCII + direction matrix 6 GB per step) exceed the 126 MB L2",
              "parallelism": f"{world} x independent shard, no collective"}

    import seqgen

    if args.impl == "reference":
        if rank != 0:
            return 0
        qb, qo, tb, to = seqgen.short_pairs(1000, args.pairs, args.length)
        threads = os.cpu_count() or 1
        chk, n, cells, one_step = cpu_arm(args, qb, qo, tb, to, args.type, threads, args.cpu_seconds)
        for _ in range(args.warmup):
            one_step()
        t = sum(one_step() for _ in range(args.steps))
        v = cells * args.steps / t / 1e9
        emit(({
            "impl": "reference", "metric": "alignment GCUPS (score+CIGAR)", "value": v, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": chk.kind,
                             "sample": f"{n} of {args.pairs} pairs per step, one slice per host thread"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    from bioinfo1_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # one rank per GPU: run on (and first-touch the pinned staging buffers from) the CPUs next to this GPU, so
        # that eight concurrent uploads do not all cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            config["cpu_affinity"] = "GPU-local (nvmlDeviceSetCpuAffinity)"
        except Exception as e:   # affinity is an optimisation, never a requirement
            config["cpu_affinity"] = "unchanged (%s)" % type(e).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    qb, qo, tb, to = seqgen.short_pairs(1000 + rank, args.pairs, args.length)
    n = args.pairs
    ctx = capi.Context(local_rank)
    L = capi.lib()
    sampler = ClockSampler(local_rank)   # nvidia-smi needs ~0.1 s to start: launch it before the uploads
    sampler.start()

    # ---- device-resident arm ---------------------------------------------------------
    d_q = torch.from_numpy(qb).to(dev)
    d_t = torch.from_numpy(tb).to(dev)
    import ctypes as C
    plan = C.c_void_p()
    capi.check(L.b200_align_plan_create(ctx.h, n, qo.ctypes.data, to.ctypes.data, args.type, 1, -1, -1, 1,
                                        C.byref(plan)))
    cells = int(L.b200_align_plan_cells(plan))
    cigar_cap = int(min(L.b200_align_plan_cigar_bound(plan), 64 * n + (1 << 20)))
    d_score = torch.empty(n, dtype=torch.int32, device=dev)
    d_tb = torch.empty(n, dtype=torch.int32, device=dev)
    d_cig = torch.empty(cigar_cap, dtype=torch.uint8, device=dev)
    d_coff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        capi.check(L.b200_align_plan_run(plan, d_q.data_ptr(), d_t.data_ptr(), d_score.data_ptr(), d_tb.data_ptr(),
                                         d_cig.data_ptr(), d_coff.data_ptr(), cigar_cap, stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler.mark_begin()   # the sampler has been running since start-up; count samples from the warm-up on
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.set_option("reset_counters", 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    launches = ctx.counter("kernel_launches")
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    t_dev = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_max = float(t_dev.item())
    value = cells * world * args.steps / (ms_max * 1e-3) / 1e9

    # dominant-kernel time, measured live with CUDA events on the launching stream (separate
    # passes so the brackets do not perturb the headline number)
    ctx.set_option("profile", 1)
    ctx.set_option("reset_counters", 1)
    prof_steps = max(2, min(args.steps, 5))
    for _ in range(prof_steps):
        step_device()
    torch.cuda.synchronize()
    fill_ns, fill_l = ctx.counter("fill_ns"), max(1, ctx.counter("fill_launches"))
    walk_ns, emit_ns, other_ns = ctx.counter("walk_ns"), ctx.counter("emit_ns"), ctx.counter("other_ns")
    ctx.set_option("profile", 0)
    peaks = load_peaks()
    fill_s = fill_ns * 1e-9 / prof_steps   # all fill launches of one step (a step is cut into waves)
    # DRAM bytes of one fill launch from the committed `ncu --set full` capture of this kernel on this workload
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01b.json")))
        if args.pairs == tr["fill_short_kernel"]["pairs"] and args.length == 150 and args.type == 0:
            traffic = tr["fill_short_kernel"]["dram_bytes_per_launch"]
    except Exception:
        pass
    achieved_tops = cells * OPS_PER_CELL / fill_s / 1e12
    roofline = {"bound": "int-alu", "kernel": "DP fill", "achieved": achieved_tops, "peak": peaks["int_tops"],
                "unit": "Tint-op/s", "frac": achieved_tops / peaks["int_tops"], "traffic": traffic,
                "ops_per_cell": OPS_PER_CELL, "fill_gcups": cells / fill_s / 1e9, "fill_ms_per_step": fill_s * 1e3,
                "fill_launches_per_step": fill_l / prof_steps,
                "peak_source": peaks["int_source"],
                "step_breakdown_ms": {"fill": fill_ns / prof_steps / 1e6, "walk": walk_ns / prof_steps / 1e6,
                                      "emit": emit_ns / prof_steps / 1e6, "other": other_ns / prof_steps / 1e6},
                "hbm": {"dir_bytes_written_per_step": cells / 4, "hbm_peak_gbs": peaks["hbm_gbs"],
                        "dir_store_gbs": cells / 4 / fill_s / 1e9, "hbm_source": peaks["source"]}}

    if args.device_only:
        emit({"value": value, "ms_per_step": ms_max / args.steps, "roofline": roofline, "clocks": clocks})
        return 0

    # ---- end-to-end arm: host buffers through the public host C-ABI ---------------------
    hq = torch.from_numpy(qb).pin_memory()
    ht = torch.from_numpy(tb).pin_memory()
    h_score = torch.empty(n, dtype=torch.int32).pin_memory()
    h_tb = torch.empty(n, dtype=torch.int32).pin_memory()
    h_cig = torch.empty(cigar_cap, dtype=torch.uint8).pin_memory()
    h_coff = torch.empty(n + 1, dtype=torch.int64).pin_memory()

    def step_host():
        capi.check(L.b200_align_batch_packed(ctx.h, n, hq.data_ptr(), qo.ctypes.data, ht.data_ptr(), to.ctypes.data,
                                             args.type, 1, -1, -1, h_score.data_ptr(), h_tb.data_ptr(),
                                             h_cig.data_ptr(), h_coff.data_ptr(), cigar_cap))

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        step_host()
    ctx.set_option("reset_counters", 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    h2d = ctx.counter("h2d_bytes") // e2e_steps
    d2h = ctx.counter("d2h_bytes") // e2e_steps
    t_e = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_val = cells * world * e2e_steps / float(t_e.item()) / 1e9
    # the host and device arms must agree bit for bit
    assert torch.equal(h_score, d_score.cpu()), "host and device arms disagree"

    out = {"metric": "alignment GCUPS (score+CIGAR)", "value": value, "unit": "GCUPS", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": t_e2e / e2e_steps * 1e3, "steps": e2e_steps},
           "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        chk, ns, ccells, one_step = cpu_arm(args, qb, qo, tb, to, args.type, 1, args.cpu_seconds)
        t = one_step()
        out["cpu_baseline"] = {"value": ccells / t / 1e9, "unit": "GCUPS", "cores": 1, "kind": chk.kind,
                               "sample": f"first {ns} of {n} pairs, single thread, {t:.1f} s",
                               "host_cpus": os.cpu_count()}
        # parity spot check of the sampled pairs against the GPU result
        sc, tbeg, _ = chk.align_batch(qb, qo[:2049], tb, to[:2049], args.type)
        assert np.array_equal(sc, h_score.numpy()[:2048]), "GPU scores differ from the CPU reference"
    L.b200_align_plan_destroy(plan)
    if rank == 0 and not args.no_extra:
        try:
            out["extra"] = extra_workloads(ctx, capi, torch, dev, peaks)
        except Exception as e:  # the headline line must survive a failure of the side measurements
            out["extra"] = {"error": repr(e)}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
