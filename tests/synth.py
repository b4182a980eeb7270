"""ctypes front end of tests/synth.c: seeded synthetic inputs of the BASELINE.json shapes (test / bench
infrastructure). Every item is generated from its own splitmix64 stream, so a rank can make just its shard
and the bytes do not depend on numpy, the thread count or the slice boundaries."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "synth.c")
_SO = os.path.join(HERE, "libsynth.so")
_lib = None

ONT_ERR = dict(sub=0.024, ins=0.048, dele=0.048)   # 12 % error, 40 % del / 40 % ins / 20 % sub


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u64, u32, dbl, vp = C.c_uint64, C.c_uint32, C.c_double, C.c_void_p
        L.synth_dna.argtypes = [u64, u64, u64, vp]
        L.synth_dna.restype = None
        L.synth_ont_spans.argtypes = [u64, u64, u64, dbl, u32, u32, vp]
        L.synth_ont_spans.restype = None
        L.synth_ont_reads.argtypes = [u64, u64, u64, vp, u64, vp, dbl, dbl, dbl, vp, vp, vp, vp, vp]
        L.synth_ont_reads.restype = None
        L.synth_pairs.argtypes = [u64, u64, u64, u32, dbl, dbl, dbl, vp, vp, vp]
        L.synth_pairs.restype = None
        _lib = L
    return _lib


def _threads():
    return max(1, min(16, os.cpu_count() or 1))


def _parallel(n, fn):
    """fn(a, b) over [0, n) cut into one slice per thread (ctypes calls release the GIL)."""
    t = _threads()
    if n < 4 * t:
        fn(0, n)
        return
    cuts = [n * k // t for k in range(t + 1)]
    th = [threading.Thread(target=fn, args=(cuts[k], cuts[k + 1])) for k in range(t)]
    for x in th:
        x.start()
    for x in th:
        x.join()


def dna(seed, n):
    """n uniform ACGT bases (+ one NUL) of stream `seed`."""
    out = np.zeros(n + 1, dtype=np.uint8)
    L = lib()
    _parallel(n, lambda a, b: L.synth_dna(seed, a, b - a, out.ctypes.data + a))
    return out


def ont_spans(seed, n, mean=8000.0, lo=1000, hi=40000):
    span = np.empty(max(n, 1), dtype=np.uint32)
    lib().synth_ont_spans(seed, 0, n, mean, lo, hi, span.ctypes.data)
    return span[:n]


def ont_reads(seed, ref, idx=None, n=None, mean=8000.0, lo=1000, hi=40000, sub=ONT_ERR["sub"], ins=ONT_ERR["ins"],
              dele=ONT_ERR["dele"], with_truth=False):
    """Reads `idx` (index array, default range(n)) of the data set `seed` over reference `ref` (uint8 array, the
    NUL at the end not counted). Returns (buf, off[len(idx)+1]) packed like the C ABI wants them."""
    L = lib()
    ref_len = len(ref) - 1 if len(ref) and ref[-1] == 0 else len(ref)
    if idx is None:
        idx = np.arange(n, dtype=np.int64)
    idx = np.asarray(idx, dtype=np.int64)
    m = len(idx)
    lens = np.zeros(max(m, 1), dtype=np.uint32)
    starts = np.zeros(max(m, 1), dtype=np.uint64)
    strands = np.zeros(max(m, 1), dtype=np.uint8)
    spans = np.zeros(max(m, 1), dtype=np.uint32)
    for k, i in enumerate(idx):   # spans are per-index streams: cheap
        L.synth_ont_spans(seed, int(i), int(i) + 1, mean, lo, hi, spans.ctypes.data + 4 * k)

    def runs(a, b):
        """maximal runs of consecutive indices inside [a, b)"""
        k = a
        while k < b:
            e = k + 1
            while e < b and idx[e] == idx[e - 1] + 1:
                e += 1
            yield k, e
            k = e

    def dry(a, b):
        for k, e in runs(a, b):
            L.synth_ont_reads(seed, int(idx[k]), int(idx[k]) + (e - k), ref.ctypes.data, ref_len, spans.ctypes.data + 4 * k,
                              sub, ins, dele, lens.ctypes.data + 4 * k, None, None, starts.ctypes.data + 8 * k,
                              strands.ctypes.data + k)
    _parallel(m, dry)
    off = np.zeros(m + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens[:m], dtype=np.uint64)
    buf = np.zeros(int(off[m]) + 1, dtype=np.uint8)

    def fill(a, b):
        for k, e in runs(a, b):
            L.synth_ont_reads(seed, int(idx[k]), int(idx[k]) + (e - k), ref.ctypes.data, ref_len, spans.ctypes.data + 4 * k,
                              sub, ins, dele, None, buf.ctypes.data, off.ctypes.data + 8 * k, None, None)
    _parallel(m, fill)
    if with_truth:
        return buf, off, starts[:m], strands[:m], spans[:m]
    return buf, off


def pairs(seed, n, length, sub, ins, dele, first=0):
    """Pairs [first, first+n) of the fixed-length pair data set `seed`: (qbuf, qoff, tbuf, toff)."""
    L = lib()
    qbuf = np.zeros(n * length + 1, dtype=np.uint8)
    tbuf = np.zeros(n * length + 1, dtype=np.uint8)

    def work(a, b):
        scratch = np.empty(2 * length + 8, dtype=np.uint8)
        L.synth_pairs(seed, first + a, first + b, length, sub, ins, dele, qbuf.ctypes.data + a * length,
                      tbuf.ctypes.data + a * length, scratch.ctypes.data)
    _parallel(n, work)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    return qbuf, off, tbuf, off.copy()


def short_pairs(seed, n, length=150, first=0):
    """BASELINE config 2 shape: 5 % substitutions + 1 % indels."""
    return pairs(seed, n, length, 0.05, 0.005, 0.005, first)


def long_pairs(seed, n, length=10000, first=0):
    """BASELINE config 5 shape: both sequences `length` long, query = target at 12 % indel-heavy error."""
    return pairs(seed, n, length, ONT_ERR["sub"], ONT_ERR["ins"], ONT_ERR["dele"], first)
