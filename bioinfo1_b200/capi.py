"""ctypes binding of include/b200map.h (harness for tests/ and bench.py).

Every call goes to the CUDA library; if it cannot be loaded, or no sm_100 GPU is visible,
the functions raise -- nothing here computes an alignment or a minimizer on the CPU.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

GLOBAL, LOCAL, SEMIGLOBAL = 0, 1, 2
E_TYPE, E_NOMEM, E_CUDA, E_CAP, E_ARG, E_NOGPU = -1, -2, -3, -4, -5, -6


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200map error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.lib_path()
        if not os.path.exists(path):
            raise B200Error(E_NOGPU, f"{path} is missing: run __graft_entry__.build() first (no fallback exists)")
        L = C.CDLL(path)
        vp, u64, u32, i32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
        L.b200_last_error.restype = C.c_char_p
        L.b200_device_count.restype = C.c_int
        L.b200_version.restype = C.c_int
        L.b200_ctx_create.argtypes = [i32, C.POINTER(vp)]
        L.b200_ctx_destroy.argtypes = [vp]
        L.b200_ctx_destroy.restype = None
        L.b200_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.b200_ctx_get_counter.argtypes = [vp, C.c_char_p]
        L.b200_ctx_get_counter.restype = C.c_int64
        L.b200_align_batch.argtypes = [i32, sz, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, u64]
        L.b200_align_batch_packed.argtypes = [vp, sz, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, u64]
        L.b200_align_plan_create.argtypes = [vp, sz, vp, vp, i32, i32, i32, i32, i32, C.POINTER(vp)]
        L.b200_align_plan_destroy.argtypes = [vp]
        L.b200_align_plan_destroy.restype = None
        L.b200_align_plan_cells.argtypes = [vp]
        L.b200_align_plan_cells.restype = u64
        L.b200_align_plan_cigar_bound.argtypes = [vp]
        L.b200_align_plan_cigar_bound.restype = u64
        L.b200_align_plan_run.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64, vp]
        L.b200_minimize_count.argtypes = [u32, u32, u32]
        L.b200_minimize_count.restype = u64
        L.b200_minimize_batch.argtypes = [i32, sz, vp, vp, u32, u32, vp, vp, vp, vp, vp, u64]
        L.b200_minimize_batch_packed.argtypes = [vp, sz, vp, vp, u32, u32, vp, vp, vp, vp, vp, u64]
        L.b200_min_plan_create.argtypes = [vp, sz, vp, u32, u32, vp, C.POINTER(vp)]
        L.b200_min_plan_destroy.argtypes = [vp]
        L.b200_min_plan_destroy.restype = None
        L.b200_min_plan_tuples.argtypes = [vp]
        L.b200_min_plan_tuples.restype = u64
        L.b200_min_plan_out_off.argtypes = [vp]
        L.b200_min_plan_out_off.restype = C.POINTER(u64)
        L.b200_min_plan_run.argtypes = [vp, vp, vp, vp, vp, vp]
        L.b200_index_build.argtypes = [vp, vp, u64, u32, u32, C.c_double, C.POINTER(vp)]
        L.b200_index_destroy.argtypes = [vp]
        L.b200_index_destroy.restype = None
        L.b200_index_stats.argtypes = [vp, vp]
        L.b200_map_batch.argtypes = [vp, vp, sz, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, u64]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data


def pack(seqs):
    """list of bytes -> (uint8 buffer, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8).copy()
    return buf, off


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().b200_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().b200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        check(lib().b200_ctx_set_option(self.h, key.encode(), int(value)))

    def counter(self, key):
        return int(lib().b200_ctx_get_counter(self.h, key.encode()))

    # ---- host-buffer entry points -------------------------------------------------
    def align_packed(self, qbuf, qoff, tbuf, toff, typ, match=1, mismatch=-1, gap=-1, want_cigar=True,
                     cigar_cap=None):
        n = len(qoff) - 1
        score = np.empty(max(n, 1), dtype=np.int32)
        tb = np.empty(max(n, 1), dtype=np.uint32)
        if want_cigar:
            if cigar_cap is None:
                cigar_cap = int(2 * (int(qoff[-1] - qoff[0]) + int(toff[-1] - toff[0])) + 2 * n + 16)
            cig = np.empty(cigar_cap, dtype=np.uint8)
            coff = np.zeros(n + 1, dtype=np.uint64)
        else:
            cig, coff, cigar_cap = None, None, 0
        check(lib().b200_align_batch_packed(self.h, n, _ptr(qbuf), _ptr(qoff), _ptr(tbuf), _ptr(toff), typ, match,
                                            mismatch, gap, _ptr(score), _ptr(tb), _ptr(cig), _ptr(coff), cigar_cap))
        return score[:n], tb[:n], cig, coff

    def align(self, queries, targets, typ, match=1, mismatch=-1, gap=-1, want_cigar=True):
        """lists of bytes -> list of (score, target_begin, cigar bytes | None)"""
        qbuf, qoff = pack(queries)
        tbuf, toff = pack(targets)
        score, tb, cig, coff = self.align_packed(qbuf, qoff, tbuf, toff, typ, match, mismatch, gap, want_cigar)
        out = []
        for i in range(len(queries)):
            c = bytes(cig[int(coff[i]):int(coff[i + 1])]) if want_cigar else None
            out.append((int(score[i]), int(tb[i]), c))
        return out

    def minimize_packed(self, buf, off, k, w, is_fwd=None):
        n = len(off) - 1
        lens = (off[1:] - off[:-1]).astype(np.uint64)
        cap = int(sum(int(lib().b200_minimize_count(int(v), k, w)) for v in lens))
        h = np.empty(max(cap, 1), dtype=np.uint32)
        p = np.empty(max(cap, 1), dtype=np.uint32)
        f = np.empty(max(cap, 1), dtype=np.uint8)
        ooff = np.zeros(n + 1, dtype=np.uint64)
        fw = None if is_fwd is None else np.ascontiguousarray(is_fwd, dtype=np.uint8)
        check(lib().b200_minimize_batch_packed(self.h, n, _ptr(buf), _ptr(off), k, w, _ptr(fw), _ptr(h), _ptr(p),
                                               _ptr(f), _ptr(ooff), cap))
        return h[:cap], p[:cap], f[:cap], ooff

    def minimize(self, seqs, k, w, is_fwd=None):
        """list of bytes -> list of (hash[], pos[], flag[])"""
        buf, off = pack(seqs)
        h, p, f, ooff = self.minimize_packed(buf, off, k, w, is_fwd)
        return [(h[int(ooff[i]):int(ooff[i + 1])], p[int(ooff[i]):int(ooff[i + 1])], f[int(ooff[i]):int(ooff[i + 1])])
                for i in range(len(seqs))]


MAPPING_DTYPE = np.dtype([("mapped", np.uint32), ("strand_fwd", np.uint32), ("q_begin", np.uint32),
                          ("q_end", np.uint32), ("t_begin", np.uint32), ("t_end", np.uint32), ("score", np.int32),
                          ("target_begin", np.uint32)])


class Index:
    """b200_index: the reference's minimizer index (team_mapper.cpp:412-477) resident on the GPU."""

    def __init__(self, ctx, ref: bytes, k=15, w=5, f=0.001):
        self.ctx = ctx
        self.h = C.c_void_p()
        self._ref = np.frombuffer(ref, dtype=np.uint8)
        check(lib().b200_index_build(ctx.h, self._ref.ctypes.data, len(ref), k, w, float(f), C.byref(self.h)))
        self.ref_len = len(ref)

    def stats(self):
        a = np.zeros(6, dtype=np.uint64)
        check(lib().b200_index_stats(self.h, a.ctypes.data))
        return a

    def close(self):
        if self.h:
            lib().b200_index_destroy(self.h)
            self.h = C.c_void_p()

    def map_packed(self, buf, off, fastq_semantics=True, typ=0, match=1, mismatch=-1, gap=-1, want_cigar=True,
                   ctx=None, out=None, cig=None, coff=None):
        """Packed reads (uint8 buffer + uint64 offsets[n+1]) -> (MAPPING_DTYPE array, cigar bytes, cigar_off).
        `ctx` may be another context of the same device (two batches in flight share one index); the result
        arrays may be passed in to be reused."""
        n = len(off) - 1
        if out is None:
            out = np.zeros(max(n, 1), dtype=MAPPING_DTYPE)
        cap = int(2 * (int(off[n] - off[0]) + n) + 64) * 2 if want_cigar else 0
        if want_cigar and (cig is None or len(cig) < cap):
            cig = np.empty(max(cap, 1), dtype=np.uint8)
        if want_cigar and coff is None:
            coff = np.zeros(n + 1, dtype=np.uint64)
        check(lib().b200_map_batch((ctx or self.ctx).h, self.h, n, buf.ctypes.data, off.ctypes.data,
                                   1 if fastq_semantics else 0, typ, match, mismatch, gap, 1 if want_cigar else 0,
                                   out.ctypes.data, cig.ctypes.data if want_cigar else None,
                                   coff.ctypes.data if want_cigar else None, len(cig) if want_cigar else 0))
        return out[:n], (cig if want_cigar else None), (coff if want_cigar else None)

    def map_batch(self, reads, fastq_semantics=True, typ=0, match=1, mismatch=-1, gap=-1, want_cigar=True):
        """list of bytes -> (structured array of MAPPING_DTYPE, list of cigar bytes | None)"""
        buf, off = pack(reads)
        n = len(reads)
        out, cig, coff = self.map_packed(buf, off, fastq_semantics, typ, match, mismatch, gap, want_cigar)
        cigs = [bytes(cig[int(coff[i]):int(coff[i + 1])]) for i in range(n)] if want_cigar else None
        return out, cigs


def align_batch_pointers(device, queries, targets, typ, match=1, mismatch=-1, gap=-1, want_cigar=True):
    """Exercises the reference-shaped pointer-array entry point b200_align_batch."""
    n = len(queries)
    qa = (C.c_char_p * max(n, 1))(*queries)
    ta = (C.c_char_p * max(n, 1))(*targets)
    ql = np.array([len(q) for q in queries], dtype=np.uint32)
    tl = np.array([len(t) for t in targets], dtype=np.uint32)
    score = np.empty(max(n, 1), dtype=np.int32)
    tb = np.empty(max(n, 1), dtype=np.uint32)
    cap = int(2 * (int(ql.sum()) + int(tl.sum())) + 2 * n + 16)
    cig = np.empty(cap, dtype=np.uint8) if want_cigar else None
    coff = np.zeros(n + 1, dtype=np.uint64) if want_cigar else None
    check(lib().b200_align_batch(device, n, C.cast(qa, C.c_void_p), _ptr(ql), C.cast(ta, C.c_void_p), _ptr(tl), typ,
                                 match, mismatch, gap, _ptr(score), _ptr(tb), _ptr(cig), _ptr(coff),
                                 cap if want_cigar else 0))
    return [(int(score[i]), int(tb[i]), bytes(cig[int(coff[i]):int(coff[i + 1])]) if want_cigar else None)
            for i in range(n)]
