"""ctypes bindings for the two CPU checkers (test infrastructure only).

* ``oracle``  -- oracle/liboracle.so, our C restatement (always present after build()).
* ``ref``     -- oracle/_ref/libref.so, the unmodified reference translation units behind
                 oracle/ref_shim.cpp (present when built in the authoring container; it
                 travels to the GPU box as a prebuilt file).

Both expose the same call shapes so tests can be parametrised over them.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLOBAL, LOCAL, SEMI = 0, 1, 2
TYPE_NAMES = {GLOBAL: "global", LOCAL: "local", SEMI: "semiGlobal"}


class _Checker:
    def __init__(self, path, prefix):
        self.path = path
        self.lib = C.CDLL(path)
        self._align = getattr(self.lib, prefix + "_align")
        self._align.restype = C.c_int
        self._align.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_char_p,
                                C.c_uint64, C.POINTER(C.c_uint64)]
        self._minimize = getattr(self.lib, prefix + "_minimize")
        self._minimize.restype = C.c_int64
        self._minimize.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_uint64]

        self._align_batch = getattr(self.lib, prefix + "_align_batch")
        self._align_batch.restype = C.c_int64
        self._align_batch.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        self._align_batch_cigar = getattr(self.lib, prefix + "_align_batch_cigar")
        self._align_batch_cigar.restype = C.c_int64
        self._align_batch_cigar.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        self._minimize_batch = getattr(self.lib, prefix + "_minimize_batch")
        self._minimize_batch.restype = C.c_int64
        self._minimize_batch.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        self.kind = "reference" if prefix == "ref" else "port"

    def align_batch_full(self, qbuf, qoff, tbuf, toff, typ, match=1, mismatch=-1, gap=-1, threads=1):
        """Packed numpy buffers -> (score[], target_begin[], cigar bytes (uint8[]), cigar_off[n+1]) with every
        CIGAR kept; the batch is cut into one slice per thread (Align is re-entrant; calls release the GIL)."""
        import threading
        import numpy as np
        n = len(qoff) - 1
        score = np.empty(max(n, 1), dtype=np.int32)
        tb = np.empty(max(n, 1), dtype=np.uint32)
        threads = max(1, min(threads, n)) if n else 1
        cuts = [n * k // threads for k in range(threads + 1)]
        parts = [None] * threads
        err = []

        def work(k):
            a, b = cuts[k], cuts[k + 1]
            ql = qoff[a:b + 1]
            tl = toff[a:b + 1]
            cap = int(2 * ((ql[-1] - ql[0]) + (tl[-1] - tl[0])) + 16 * (b - a) + 64)
            buf = np.empty(cap, dtype=np.uint8)
            off = np.zeros(b - a + 1, dtype=np.uint64)
            rc = self._align_batch_cigar(b - a, qbuf.ctypes.data, ql.ctypes.data, tbuf.ctypes.data, tl.ctypes.data, typ,
                                         match, mismatch, gap, score.ctypes.data + 4 * a, tb.ctypes.data + 4 * a,
                                         buf.ctypes.data, cap, off.ctypes.data)
            if rc != b - a:
                err.append(rc)
            parts[k] = (buf, off)
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            raise ValueError(f"align_batch_cigar rc={err[0]}")
        coff = np.zeros(n + 1, dtype=np.uint64)
        base = 0
        for k in range(threads):
            a, b = cuts[k], cuts[k + 1]
            coff[a + 1:b + 1] = parts[k][1][1:] + np.uint64(base)
            base += int(parts[k][1][-1])
        cig = np.concatenate([parts[k][0][:int(parts[k][1][-1])] for k in range(threads)]) if n else np.zeros(0, np.uint8)
        return score[:n], tb[:n], cig, coff

    def minimize_batch(self, buf, off, k, w, is_fwd=True):
        """Packed sequences -> (hash[], pos[], flag[], out_off[n+1]); single thread (the reference's Minimize keeps
        process-global state). `buf` must be readable a few bytes past every sequence (callers pass packed buffers)."""
        import numpy as np
        n = len(off) - 1
        cap = int(sum(int(off[i + 1] - off[i]) + w for i in range(n))) + 8
        h = np.empty(cap, dtype=np.uint32)
        p = np.empty(cap, dtype=np.uint32)
        f = np.empty(cap, dtype=np.uint8)
        oo = np.zeros(n + 1, dtype=np.uint64)
        tot = self._minimize_batch(n, buf.ctypes.data, off.ctypes.data, k, w, 1 if is_fwd else 0, h.ctypes.data,
                                   p.ctypes.data, f.ctypes.data, cap, oo.ctypes.data)
        if tot < 0:
            raise ValueError(f"minimize_batch rc={tot}")
        return h[:tot], p[:tot], f[:tot], oo

    def align_batch(self, qbuf, qoff, tbuf, toff, typ, match=1, mismatch=-1, gap=-1, want_cigar=True):
        """Packed numpy buffers -> (score[], target_begin[], total cigar bytes); releases the GIL."""
        import numpy as np
        n = len(qoff) - 1
        score = np.empty(max(n, 1), dtype=np.int32)
        tb = np.empty(max(n, 1), dtype=np.uint32)
        nbytes = C.c_uint64(0)
        rc = self._align_batch(n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data, toff.ctypes.data, typ, match,
                               mismatch, gap, 1 if want_cigar else 0, score.ctypes.data, tb.ctypes.data,
                               C.byref(nbytes))
        if rc != n:
            raise ValueError(f"align_batch rc={rc}")
        return score[:n], tb[:n], nbytes.value

    def align(self, q: bytes, t: bytes, typ: int, match=1, mismatch=-1, gap=-1, want_cigar=True):
        """-> (score, target_begin, cigar_bytes | None)"""
        score = C.c_int32(0)
        tb = C.c_uint32(0)
        cap = 8 * (len(q) + len(t)) + 16
        buf = C.create_string_buffer(cap)
        clen = C.c_uint64(0)
        rc = self._align(q, len(q), t, len(t), typ, match, mismatch, gap, 1 if want_cigar else 0,
                         C.byref(score), C.byref(tb), buf, cap, C.byref(clen))
        if rc != 0:
            raise ValueError(f"align rc={rc}")
        return score.value, tb.value, (buf.raw[:clen.value] if want_cigar else None)

    def minimize(self, seq: bytes, k: int, w: int, is_fwd=True):
        """-> list of (hash, pos, flag)"""
        import numpy as np
        padded = seq + b"\0" * (k + w + 8)
        n = self._minimize(padded, len(seq), k, w, 1 if is_fwd else 0, None, None, None, 0)
        if n < 0:
            raise ValueError(f"minimize rc={n}")
        h = np.empty(max(n, 1), dtype=np.uint32)
        p = np.empty(max(n, 1), dtype=np.uint32)
        f = np.empty(max(n, 1), dtype=np.uint8)
        n2 = self._minimize(padded, len(seq), k, w, 1 if is_fwd else 0, h.ctypes.data, p.ctypes.data,
                            f.ctypes.data, n)
        assert n2 == n
        return h[:n], p[:n], f[:n]


def load_oracle():
    return _Checker(os.path.join(ROOT, "oracle", "liboracle.so"), "oracle")


def load_ref():
    p = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    return _Checker(p, "ref") if os.path.exists(p) else None
