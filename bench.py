#!/usr/bin/env python3
"""bench.py -- the headline measurement: batched pairwise alignment GCUPS (score + CIGAR), plus the sharded
configs `north_star` names (mapped reads/s, local 10 kb x 10 kb) at the same N.

Headline workload (BASELINE.json configs[1]): 1 048 576 pairs of 150 bp x 150 bp, global NW, scores 1/-1/-1,
CIGAR + target_begin produced, synthetic data (tests/seqgen.short_pairs). One "step" is one pass of the whole
hot path (2-bit pack -> DP fill -> traceback walk -> CIGAR emit) over the batch. With N GPUs every rank owns its
own batch of that size (weak scaling, no collective on the data path); `value` = cells of all ranks /
max-over-ranks device time.

  value        inputs resident in HBM, device-resident C-ABI (b200_align_plan_run), CUDA events
  e2e          same batch through the host-buffer C-ABI (b200_align_batch_packed): pinned host buffers in, host
               arrays out, H2D/D2H inside the timed region; with the box's concurrent H2D ceiling measured in
               the same run (`h2d_ceiling_gbs`) and the pointer-array entry point the reference-shaped callers use
  roofline     the DP fill kernel against the measured integer-ALU issue peak, with the kernel's own
               issue-slot ceiling and the alu-pipe utilisation of the committed ncu capture beside it
  extra.c4_strong / extra.c5_strong   BASELINE configs[3] and [4] at their stated sizes, STRONG scaling: the
               100 000 reads / 10 000 pairs are dealt to the ranks by cost (bioinfo1_b200.shard.partition), every
               rank maps / aligns its shard, numbers = work of all ranks / max-over-ranks time
  cpu_baseline / --impl reference: the reference's own CPU Align on the host cores

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import ctypes as C
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
DTYPE = "int16x2 packed (bit-exact to the reference's int32)"
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")   # falls back to the round-1 capture
NCU_SUMMARY_OLD = os.path.join(ROOT, "profiles", "ncu_traffic_r01b.json")


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


def cpu_arm(qb, qo, tb, to, typ, threads, budget_s, kind_pref="reference"):
    """Times the reference's CPU Align (oracle/_ref when present, else the C port) on a bounded
    sample of the same batch; returns (checker, sample size, cells, step function)."""
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
        def work(k):
            a, b = int(bounds[k]), int(bounds[k + 1])
            if b > a:
                chk.align_batch(qb, qo[a:b + 1], tb, to[a:b + 1], typ)
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    cells = float(np.sum((qo[1:n + 1] - qo[:n]).astype(np.float64) * (to[1:n + 1] - to[:n]).astype(np.float64)))
    return chk, n, cells, one_step


class Dist:
    """The few collectives the bench needs (all outside the data path): barrier, max / sum of a scalar."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, v, op):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._red(v, self.dist.ReduceOp.MAX)

    def min(self, v):
        return self._red(v, self.dist.ReduceOp.MIN)

    def sum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM)


# ------------------------------------------------------------------ strong-scaling configs (all ranks) ----
class Phases:
    """Local work of one rank in guarded phases: after a failure the later phases are skipped, but the rank still
    takes part in every collective (a rank that raised in the middle would leave the others waiting in NCCL)."""

    def __init__(self):
        self.err = None

    def run(self, fn):
        if self.err is None:
            try:
                fn()
            except Exception as e:
                self.err = repr(e)

    def failed_anywhere(self, D):
        return D.max(1.0 if self.err else 0.0) > 0


def c4_strong(args, D, rank, world, local_rank, ctx, capi, torch):
    """BASELINE configs[3]: end-to-end mapping of 100 000 ONT-like reads (4.6 Mbp reference, k=15 w=5 f=0.001,
    semiGlobal + CIGAR) through b200_index_build / b200_map_batch with host buffers in and out. The reads are dealt to
    the ranks by cost (span^2: the alignment dominates), each rank builds its own index and keeps two 8 192-read
    batches in flight (two contexts / host threads, the same setting at every N)."""
    import synth
    from bioinfo1_b200 import shard
    L = capi.lib()
    n_reads = args.c4_reads
    S = {"t_index": 0.0, "t_index_first": 0.0, "t_local": 0.0, "n_workers": 1, "batch_reads": 0}
    ph = Phases()

    def setup():
        ref = synth.dna(1, 4_600_000)
        spans = synth.ont_spans(2, n_reads).astype(np.float64)
        mine = shard.partition(spans * spans, world)[rank]
        S["buf"], S["off"] = synth.ont_reads(2, ref, idx=mine)          # this rank's reads only, input order kept
        S["n_mine"] = len(mine)
        S["ref"] = ref[:4_600_000].tobytes()
    ph.run(setup)
    D.barrier()

    def build():
        t0 = time.perf_counter()
        index = capi.Index(ctx, S["ref"], 15, 5, 0.001)
        torch.cuda.synchronize()
        S["t_index_first"] = time.perf_counter() - t0
        index.close()
        t0 = time.perf_counter()
        S["index"] = capi.Index(ctx, S["ref"], 15, 5, 0.001)
        torch.cuda.synchronize()
        S["t_index"] = time.perf_counter() - t0
        n_mine, off = S["n_mine"], S["off"]
        # The same rule at every N: two batches in flight per GPU, the rank's reads cut into an EVEN number of equal
        # batches of at most --c4-batch reads (so that both workers get the same share whatever the shard size).
        n_batches = 2 * max(1, -(-n_mine // (2 * args.c4_batch)))
        per = -(-n_mine // n_batches)
        batches = [(a, min(n_mine, a + per)) for a in range(0, n_mine, per)]
        S["batch_reads"] = per
        n_workers = max(1, min(2, len(batches)))
        S["batches"], S["n_workers"] = batches, n_workers
        S["ctxs"] = [ctx] + [capi.Context(local_rank) for _ in range(n_workers - 1)]
        max_n = max((b - a for a, b in batches), default=1)
        S["max_cap"] = max((int(4 * int(off[b] - off[a]) + 64 * (b - a) + 64) for a, b in batches), default=64)
        S["outs"] = [(np.zeros(max_n, dtype=capi.MAPPING_DTYPE), np.empty(S["max_cap"], dtype=np.uint8),
                      np.zeros(max_n + 1, dtype=np.uint64)) for _ in range(n_workers)]
        S["mapped_of"] = np.zeros(max(n_mine, 1), dtype=np.uint8)
        S["score_sum"] = [0] * n_workers
        S["cigar_bytes"] = [0] * n_workers
    ph.run(build)

    def run_all(which, record):
        nxt = [0]
        lock = threading.Lock()
        errors = []
        buf, off, index, ctxs, outs, max_cap = S["buf"], S["off"], S["index"], S["ctxs"], S["outs"], S["max_cap"]

        def worker(wi):
            out, cig, coff = outs[wi]
            try:
                while True:
                    with lock:
                        bi = nxt[0]; nxt[0] += 1
                    if bi >= len(which):
                        return
                    a, b = which[bi]
                    n = b - a
                    boff = np.ascontiguousarray(off[a:b + 1])
                    capi.check(L.b200_map_batch(ctxs[wi].h, index.h, n, buf.ctypes.data, boff.ctypes.data, 1, 2, 1, -1, -1, 1,
                                                out.ctypes.data, cig.ctypes.data, coff.ctypes.data, max_cap))
                    if record:
                        S["mapped_of"][a:b] = out["mapped"][:n]
                        S["score_sum"][wi] += int(out["score"][:n][out["mapped"][:n] != 0].astype(np.int64).sum())
                        S["cigar_bytes"][wi] += int(coff[n])
            except Exception as e:   # surfaced after the join
                errors.append(e)
        th = [threading.Thread(target=worker, args=(wi,)) for wi in range(S["n_workers"])]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if errors:
            raise errors[0]
    # warm-up: one batch per context, so the scratch buffers have their size
    ph.run(lambda: run_all(S["batches"][:S["n_workers"]], False))
    D.barrier()

    def timed():
        t0 = time.perf_counter()
        run_all(S["batches"], True)
        torch.cuda.synchronize()
        S["t_local"] = time.perf_counter() - t0
    ph.run(timed)
    ok = ph.err is None
    t = D.max(S["t_local"])
    t_min = D.min(S["t_local"])
    mapped = D.sum(int(S["mapped_of"][:S["n_mine"]].sum()) if ok else 0)
    bases = D.sum(int(S["off"][S["n_mine"]]) if ok else 0)
    t_index, t_index_first = D.max(S["t_index"]), D.max(S["t_index_first"])
    checksum = D.sum(sum(S["score_sum"]) if ok else 0)
    cig_total = D.sum(sum(S["cigar_bytes"]) if ok else 0)
    failed = ph.failed_anywhere(D)
    if "index" in S:
        S["index"].close()
    for c in S.get("ctxs", [ctx])[1:]:
        c.close()
    if failed:
        return {"error": ph.err or "another rank failed"}
    return {"reads": n_reads, "mapped": int(mapped), "bases": int(bases), "map_s": t, "reads_per_s": mapped / t,
            "index_build_s": t_index, "index_build_first_call_s": t_index_first,
            "batch_reads": S.get("batch_reads"), "batch_rule": f"even number of equal batches of <= {args.c4_batch} reads per rank",
            "batches_in_flight": S["n_workers"], "split": "by cost (span^2), shard.partition",
            "rank_seconds_min_max": [t_min, t], "score_checksum": int(checksum), "cigar_bytes": int(cig_total),
            "note": "host buffers in, PAF fields + CIGAR out (b200_map_batch), semiGlobal 1/-1/-1, k=15 w=5 f=0.001, "
                    "FASTQ-flavoured lookup (both strands); time = wall clock of the slowest rank"}


def c5_strong(args, D, rank, world, ctx, capi, torch, dev, peaks):
    """BASELINE configs[4]: 10 000 local pairs 10 kb x 10 kb with full traceback, dealt to the ranks by cells."""
    import synth
    from bioinfo1_b200 import shard
    L = capi.lib()
    n_pairs, length = args.c5_pairs, 10000
    S = {"t_dev": 0.0, "t_host": 0.0, "cells": 0, "checksum": 0, "cig": 0}
    ph = Phases()

    def setup():
        # equal costs: LPT deals equal counts; the data set is seeded per pair, so ranks take contiguous blocks of those sizes
        sizes = [len(p) for p in shard.partition(np.full(n_pairs, float(length) * length), world)]
        lo, n = int(sum(sizes[:rank])), int(sizes[rank])
        S["n"] = n
        S["qb"], S["qo"], S["tb"], S["to"] = synth.long_pairs(5, n, length, first=lo)
        S["d_q"], S["d_t"] = torch.from_numpy(S["qb"]).to(dev), torch.from_numpy(S["tb"]).to(dev)
        plan = C.c_void_p()
        capi.check(L.b200_align_plan_create(ctx.h, n, S["qo"].ctypes.data, S["to"].ctypes.data, 1, 1, -1, -1, 1, C.byref(plan)))
        S["plan"] = plan
        S["cells"] = int(L.b200_align_plan_cells(plan))
        S["cap"] = int(L.b200_align_plan_cigar_bound(plan))
        S["d_s"] = torch.empty(n, dtype=torch.int32, device=dev); S["d_b"] = torch.empty(n, dtype=torch.int32, device=dev)
        S["d_c"] = torch.empty(S["cap"], dtype=torch.uint8, device=dev); S["d_o"] = torch.empty(n + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()

    def step():
        capi.check(L.b200_align_plan_run(S["plan"], S["d_q"].data_ptr(), S["d_t"].data_ptr(), S["d_s"].data_ptr(), S["d_b"].data_ptr(),
                                         S["d_c"].data_ptr(), S["d_o"].data_ptr(), S["cap"], st.cuda_stream))
    ph.run(setup)
    ph.run(step)
    D.barrier()

    def timed():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        e0.record(st)
        for _ in range(reps):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        S["t_dev"] = e0.elapsed_time(e1) / reps * 1e-3
    ph.run(timed)
    D.barrier()

    def host_arm():   # the same shard through the host-buffer entry point (pageable numpy buffers in, host arrays out)
        L.b200_align_plan_destroy(S.pop("plan"))
        S.pop("d_c")
        ctx.align_packed(S["qb"], S["qo"], S["tb"], S["to"], 1, 1, -1, -1, True)   # warm-up: the host path's staging buffers grow once
        t0 = time.perf_counter()
        hs, hb, hc, ho = ctx.align_packed(S["qb"], S["qo"], S["tb"], S["to"], 1, 1, -1, -1, True)
        S["t_host"] = time.perf_counter() - t0
        assert np.array_equal(hs, S["d_s"].cpu().numpy()), "config 5: host and device arms disagree"
        S["checksum"] = int(hs.astype(np.int64).sum())
        S["cig"] = int(ho[S["n"]])
    ph.run(host_arm)
    t, t_host = D.max(S["t_dev"]), D.max(S["t_host"])
    total_cells = D.sum(S["cells"])
    checksum, cig = D.sum(S["checksum"]), D.sum(S["cig"])
    if "plan" in S:
        L.b200_align_plan_destroy(S.pop("plan"))
    if ph.failed_anywhere(D):
        return {"error": ph.err or "another rank failed"}
    return {"pairs": n_pairs, "cells": total_cells, "s_per_pass": t, "gcups": total_cells / t / 1e9,
            "e2e_s_per_pass": t_host, "e2e_gcups": total_cells / t_host / 1e9, "dir_bytes": total_cells / 4,
            "dir_store_gbs_per_gpu": total_cells / 4 / t / 1e9 / world, "hbm_peak_gbs": peaks["hbm_gbs"],
            "score_checksum": int(checksum), "cigar_bytes": int(cig), "split": "by cells, shard.partition",
            "note": "local 10 kb x 10 kb at 12 % indel-heavy error, score + CIGAR + target_begin; gcups device-resident "
                    "(CUDA events, max over ranks), e2e_gcups through b200_align_batch_packed (wall clock, max over ranks)"}


# ------------------------------------------------------------------ kernel-level side measurements (rank 0) ----
def single_gpu_extras(ctx, capi, torch, dev, peaks):
    """Device-resident kernel measurements on rank 0: the long-pair fill on ONT-like pairs (the `north_star` target
    kernel) and MinimizeBatch on a read batch (config 3 shape)."""
    import seqgen
    import synth
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

    # MinimizeBatch, k=15 w=5: the reference (both strands) and 16 384 distinct ONT-like reads
    ref = synth.dna(1, 4_600_000)
    rbuf, roff = synth.ont_reads(2, ref, n=16384)
    both = np.concatenate([ref[:4_600_000], ref[:4_600_000][::-1], np.zeros(1, np.uint8)])
    for tag, (buf, off) in (("minimize_ref_4.6Mbp_x2", (both, np.array([0, 4_600_000, 9_200_000], dtype=np.uint64))),
                            ("minimize_16384_ont_reads", (rbuf, roff))):
        d_buf = torch.from_numpy(buf).to(dev)
        plan = C.c_void_p()
        capi.check(L.b200_min_plan_create(ctx.h, len(off) - 1, off.ctypes.data, 15, 5, None, C.byref(plan)))
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
    ap.add_argument("--no-extra", action="store_true", help="skip the kernel-level side measurements on rank 0")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling configs 4 and 5")
    ap.add_argument("--c4-reads", type=int, default=100_000)
    ap.add_argument("--c4-batch", type=int, default=8192)
    ap.add_argument("--c5-pairs", type=int, default=10_000)
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
        chk, n, cells, one_step = cpu_arm(qb, qo, tb, to, args.type, threads, args.cpu_seconds)
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    D = Dist(torch, dist, world, dev)

    qb, qo, tb, to = seqgen.short_pairs(1000 + rank, args.pairs, args.length)
    n = args.pairs
    ctx = capi.Context(local_rank)
    L = capi.lib()
    sampler = ClockSampler(local_rank)   # nvidia-smi needs ~0.1 s to start: launch it before the uploads
    sampler.start()

    # ---- device-resident arm ---------------------------------------------------------
    d_q = torch.from_numpy(qb).to(dev)
    d_t = torch.from_numpy(tb).to(dev)
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

    sampler.mark_begin()   # the sampler has been running since start-up; count samples from the warm-up on
    for _ in range(args.warmup):
        step_device()
    D.barrier()
    ctx.set_option("reset_counters", 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    D.barrier()
    launches = ctx.counter("kernel_launches")
    ms_max = D.max(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
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
    # DRAM bytes and pipe utilisation of one fill launch from the committed `ncu --set full` capture of this kernel on
    # this workload (a run cannot read DRAM counters itself)
    traffic = alu_pipe = issue_active = None
    traffic_src = None
    for path in (NCU_SUMMARY, NCU_SUMMARY_OLD):
        try:
            tr = json.load(open(path))["fill_short_kernel"]
            if args.pairs == tr["pairs"] and args.length == 150 and args.type == 0:
                traffic, alu_pipe, issue_active = tr["dram_bytes_per_launch"], tr.get("alu_pipe_pct"), tr.get("issue_active_pct")
                traffic_src = os.path.relpath(path, ROOT)
            break
        except Exception:
            continue
    achieved_tops = cells * OPS_PER_CELL / fill_s / 1e12
    # the kernel's own ceiling: alu-pipe issue slots (64 lanes / clk / SM) per pair of cells of the packed recurrence
    slots = ctx.counter("alu_slots_per_cell_pair_x10") / 10.0 if ctx.counter("alu_slots_per_cell_pair_x10") > 0 else 5.0
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    ceiling_tcups = sm_count * 64 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / (slots / 2.0) / 1e12
    roofline = {"bound": "int-alu", "kernel": "DP fill (fill_short_kernel)", "achieved": achieved_tops, "peak": peaks["int_tops"],
                "unit": "Tint-op/s", "frac": achieved_tops / peaks["int_tops"], "traffic": traffic, "traffic_source": traffic_src,
                "ops_per_cell": OPS_PER_CELL, "fill_gcups": cells / fill_s / 1e9, "fill_ms_per_step": fill_s * 1e3,
                "fill_launches_per_step": fill_l / prof_steps,
                "peak_source": peaks["int_source"],
                "alu_pipe_frac": None if alu_pipe is None else alu_pipe / 100.0,
                "issue_active_frac": None if issue_active is None else issue_active / 100.0,
                "issue_slots_per_cell_pair": slots, "ceiling_tcups": ceiling_tcups,
                "frac_of_kernel_ceiling": cells / fill_s / 1e12 / ceiling_tcups,
                "note": ("frac uses SURVEY 8(d)'s 9 int32 ops per cell against the measured alu-pipe issue peak; the packed "
                         "tagged recurrence spends `issue_slots_per_cell_pair` alu-pipe slots per TWO cells, so frac > 1 is "
                         "expected and frac_of_kernel_ceiling / alu_pipe_frac say how close the kernel is to its own bound"),
                "algorithmic_bytes": cells / 4 + 2 * n * args.length,
                "step_breakdown_ms": {"fill": fill_ns / prof_steps / 1e6, "walk": walk_ns / prof_steps / 1e6,
                                      "emit": emit_ns / prof_steps / 1e6, "other": other_ns / prof_steps / 1e6},
                "hbm": {"dir_bytes_written_per_step": cells / 4, "hbm_peak_gbs": peaks["hbm_gbs"],
                        "dir_store_gbs": cells / 4 / fill_s / 1e9, "hbm_source": peaks["source"]}}

    if args.device_only:
        if rank == 0:
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

    e2e_steps = args.steps
    for _ in range(max(2, min(args.warmup, 3))):
        step_host()
    ctx.set_option("reset_counters", 1)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e_local = time.perf_counter() - t0
    h2d = ctx.counter("h2d_bytes") // e2e_steps
    d2h = ctx.counter("d2h_bytes") // e2e_steps
    t_e2e = D.max(t_e2e_local)
    e2e_val = cells * world * e2e_steps / t_e2e / 1e9
    # the host and device arms must agree bit for bit: scores, target_begin, CIGAR offsets and every CIGAR byte
    total_cig = int(h_coff[n])
    assert torch.equal(h_score, d_score.cpu()), "host and device arms disagree (score)"
    assert torch.equal(h_tb, d_tb.cpu()), "host and device arms disagree (target_begin)"
    assert torch.equal(h_coff, d_coff.cpu()), "host and device arms disagree (CIGAR offsets)"
    assert torch.equal(h_cig[:total_cig], d_cig[:total_cig].cpu()), "host and device arms disagree (CIGAR bytes)"

    # The box's host-to-device ceiling with all N ranks uploading at once: the same pinned buffers, plain
    # cudaMemcpyAsync of the whole 2 x 157 MB per repetition, timed on the device, all ranks between barriers.
    cp_stream = torch.cuda.Stream()
    reps = 8
    with torch.cuda.stream(cp_stream):
        d_q.copy_(hq, non_blocking=True); d_t.copy_(ht, non_blocking=True)
    D.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cp_stream):
        c0.record(cp_stream)
        for _ in range(reps):
            d_q.copy_(hq, non_blocking=True); d_t.copy_(ht, non_blocking=True)
        c1.record(cp_stream)
    D.barrier()
    copy_bytes = (hq.numel() + ht.numel()) * reps
    t_copy = D.max(c0.elapsed_time(c1) * 1e-3)
    h2d_ceiling = copy_bytes * world / t_copy / 1e9            # aggregate over the ranks, GB/s
    h2d_bound_ms = h2d / (h2d_ceiling / world * 1e9) * 1e3     # what one step's upload alone costs at that rate
    e2e = {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": t_e2e / e2e_steps * 1e3, "steps": e2e_steps,
           "h2d_ceiling_gbs": h2d_ceiling, "h2d_ceiling_gbs_per_gpu": h2d_ceiling / world,
           "h2d_bound_ms_per_step": h2d_bound_ms, "frac_of_h2d_bound": h2d_bound_ms / (t_e2e / e2e_steps * 1e3),
           "h2d_ceiling_note": f"{world} ranks x {reps} x {copy_bytes // reps} B pinned cudaMemcpyAsync at once, device-timed, max over ranks"}

    # The reference-shaped entry point (arrays of pointers and lengths, pageable memory; team::AlignBatch and
    # INTEGRATION.md's stub call it): same batch, same outputs, its own per-thread default context.
    ptr_steps = max(2, min(args.steps, 5))
    qptr = (qb.ctypes.data + qo[:n]).astype(np.uint64)
    tptr = (tb.ctypes.data + to[:n]).astype(np.uint64)
    qlen = (qo[1:] - qo[:n]).astype(np.uint32)
    tlen = (to[1:] - to[:n]).astype(np.uint32)
    p_score = np.empty(n, dtype=np.int32); p_tb = np.empty(n, dtype=np.uint32)
    p_cig = np.empty(cigar_cap, dtype=np.uint8); p_coff = np.zeros(n + 1, dtype=np.uint64)

    def step_ptr():
        capi.check(L.b200_align_batch(local_rank, n, qptr.ctypes.data, qlen.ctypes.data, tptr.ctypes.data, tlen.ctypes.data,
                                      args.type, 1, -1, -1, p_score.ctypes.data, p_tb.ctypes.data, p_cig.ctypes.data,
                                      p_coff.ctypes.data, cigar_cap))
    for _ in range(2):
        step_ptr()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(ptr_steps):
        step_ptr()
    torch.cuda.synchronize()
    t_ptr = D.max(time.perf_counter() - t0)
    assert np.array_equal(p_score, h_score.numpy()) and np.array_equal(p_coff.astype(np.int64), h_coff.numpy()), \
        "pointer-array and packed entry points disagree"
    assert np.array_equal(p_cig[:total_cig], h_cig.numpy()[:total_cig]), "pointer-array and packed entry points disagree (CIGAR)"
    e2e["pointer_api"] = {"entry": "b200_align_batch (pointer arrays, pageable host memory)", "ms_per_step": t_ptr / ptr_steps * 1e3,
                          "value": cells * world * ptr_steps / t_ptr / 1e9, "unit": "GCUPS", "steps": ptr_steps,
                          "vs_packed": (t_ptr / ptr_steps) / (t_e2e / e2e_steps)}

    out = {"metric": "alignment GCUPS (score+CIGAR)", "value": value, "unit": "GCUPS", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config,
           "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline}

    # parity of the timed batch against the CPU reference: score, target_begin and CIGAR bytes of the first 2 048 pairs
    from cpu_checkers import load_oracle, load_ref
    chk = load_ref() or load_oracle()
    m = min(n, 2048)
    sc, tbeg, cg, co = chk.align_batch_full(qb, qo[:m + 1], tb, to[:m + 1], args.type, 1, -1, -1, threads=min(8, os.cpu_count() or 1))
    assert np.array_equal(sc, h_score.numpy()[:m]), "GPU scores differ from the CPU reference"
    assert np.array_equal(tbeg, h_tb.numpy()[:m].astype(np.uint32)), "GPU target_begin differs from the CPU reference"
    assert np.array_equal(co.astype(np.int64), h_coff.numpy()[:m + 1]), "GPU CIGAR lengths differ from the CPU reference"
    assert np.array_equal(cg, h_cig.numpy()[:int(co[m])]), "GPU CIGAR bytes differ from the CPU reference"
    out["parity_spot_check"] = {"pairs": m, "against": chk.kind, "fields": ["score", "target_begin", "cigar_off", "cigar bytes"], "ok": True}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        chk, ns, ccells, one_step = cpu_arm(qb, qo, tb, to, args.type, 1, args.cpu_seconds)
        t = one_step()
        out["cpu_baseline"] = {"value": ccells / t / 1e9, "unit": "GCUPS", "cores": 1, "kind": chk.kind,
                               "sample": f"first {ns} of {n} pairs, single thread, {t:.1f} s",
                               "host_cpus": os.cpu_count()}
    L.b200_align_plan_destroy(plan)
    del d_cig, d_q, d_t, h_cig, hq, ht
    extra = {}
    if not args.no_strong:
        for name, fn in (("c4_strong", lambda: c4_strong(args, D, rank, world, local_rank, ctx, capi, torch)),
                         ("c5_strong", lambda: c5_strong(args, D, rank, world, ctx, capi, torch, dev, peaks))):
            try:
                extra[name] = fn()
            except Exception as e:   # the headline line must survive a failure of the side measurements
                extra[name] = {"error": repr(e)}
            extra[name]["n_gpus"] = world
    if rank == 0 and not args.no_extra:
        try:
            extra["single_gpu"] = single_gpu_extras(ctx, capi, torch, dev, peaks)
        except Exception as e:
            extra["single_gpu"] = {"error": repr(e)}
    out["extra"] = extra
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
