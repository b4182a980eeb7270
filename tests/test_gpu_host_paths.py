"""GPU tests of the host-side pipelines and of the kernel variants behind one ABI:

* the host-buffer entry point with its shrinking tail of waves / wave-aligned upload slices, at sizes on both sides
  of every switch (two waves, many waves, tail longer than the batch), against the CPU checker and the device path;
* the pointer-array entry point with its threaded gather (several block counts, one thread and many);
* both substitution-term variants of the 2-bit fill kernels (PRMT and the shared-memory table) on the same batches;
* the stripe protocol of the long-pair fill under stress: thousands of multi-stripe pairs, tiny direction budget
  (many waves), concurrent walkers on and off.
"""
import ctypes as C
import os

import numpy as np
import pytest

from cpu_checkers import load_oracle, load_ref
import seqgen
import synth

pytestmark = pytest.mark.gpu
CHK = load_ref() or load_oracle()
THREADS = os.cpu_count() or 1


@pytest.fixture()
def ctx():
    from bioinfo1_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _equal(got, exp, what):
    n = len(exp[0])
    assert np.array_equal(got[0][:n], exp[0]), what + ": scores"
    assert np.array_equal(got[1][:n], exp[1]), what + ": target_begin"
    assert np.array_equal(np.asarray(got[3][:n + 1], dtype=np.uint64), np.asarray(exp[3], dtype=np.uint64)), what + ": CIGAR offsets"
    tot = int(exp[3][n])
    assert np.array_equal(got[2][:tot], exp[2][:tot]), what + ": CIGAR bytes"


@pytest.mark.parametrize("n", [8192, 70_000, 262_144, 400_001])
def test_host_pipeline_wave_schedules(ctx, n):
    qb, qo, tb, to = synth.short_pairs(77, n, 150)
    exp = CHK.align_batch_full(qb, qo, tb, to, 0, 1, -1, -1, threads=THREADS)
    for taper, stream in ((1, 1), (0, 1), (1, 0), (0, 0)):     # shrinking tail of waves; one persistent fill fed by a watermark
        ctx.set_option("taper_tail", taper)
        ctx.set_option("stream_fill", stream)
        for _ in range(2):   # twice: the second run reuses every workspace (stale flags / counters would show)
            got = ctx.align_packed(qb, qo, tb, to, 0, 1, -1, -1, True, cigar_cap=64 * n)
            _equal(got, exp, f"n={n} taper={taper} stream={stream}")
    ctx.set_option("taper_tail", 0)
    ctx.set_option("stream_fill", 0)
    # score-only through the same pipeline
    s, t_, _, _ = ctx.align_packed(qb, qo, tb, to, 0, 1, -1, -1, False)
    assert np.array_equal(s, exp[0]) and np.array_equal(t_, exp[1])


def test_streaming_fill_with_flagged_pairs_and_other_types(ctx):
    """Pairs that are not pure ACGT inside a streamed batch: flagged by the pack kernels while the persistent fill is
    already running, skipped by it, repaired afterwards -- all three alignment types."""
    n = 420_000
    qb, qo, tb, to = synth.short_pairs(81, n, 150)
    qb = qb.copy(); tb = tb.copy()
    rng = np.random.default_rng(5)
    for i in rng.integers(0, n, size=300):
        qb[int(i) * 150 + int(rng.integers(0, 150))] = ord("N")
    for i in rng.integers(0, n, size=200):
        tb[int(i) * 150 + int(rng.integers(0, 150))] = ord("-")
    tb[(n - 1) * 150 + 149] = ord("n")          # the very last base of the batch
    for typ in (0, 1, 2):
        exp = CHK.align_batch_full(qb, qo, tb, to, typ, 1, -1, -1, threads=THREADS)
        for stream in (1, 0):
            ctx.set_option("stream_fill", stream)
            got = ctx.align_packed(qb, qo, tb, to, typ, 1, -1, -1, True, cigar_cap=64 * n)
            _equal(got, exp, f"type {typ} stream={stream}")
    ctx.set_option("stream_fill", 0)


def test_host_pipeline_other_lengths_and_types(ctx):
    for length, typ in ((100, 2), (64, 1), (33, 0)):
        n = 150_000
        qb, qo, tb, to = synth.pairs(5, n, length, 0.05, 0.01, 0.01)
        exp = CHK.align_batch_full(qb, qo, tb, to, typ, 1, -1, -1, threads=THREADS)
        got = ctx.align_packed(qb, qo, tb, to, typ, 1, -1, -1, True, cigar_cap=64 * n)
        _equal(got, exp, f"length {length} type {typ}")


def _in_fresh_thread(env, fn):
    """The pointer-array entry point uses a per-thread default context whose option defaults come from the environment
    when it is created: a new thread is a new context."""
    import threading
    out = {}
    old = {k: os.environ.get(k) for k in env}

    def work():
        try:
            fn()
        except BaseException as e:   # surfaced in the caller
            out["err"] = e
    os.environ.update(env)
    try:
        t = threading.Thread(target=work)
        t.start()
        t.join()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if "err" in out:
        raise out["err"]


@pytest.mark.parametrize("host_pack", ["0", "1"])
@pytest.mark.parametrize("n", [1, 300, 9000, 150_000])
def test_pointer_array_entry_point_threaded_gather(n, host_pack):
    """host_pack = 1: the gather pass packs uniform batches to 2 bits on the host (host_pack.hpp) instead of copying."""
    _in_fresh_thread({"B200_HOST_PACK": host_pack}, lambda: _pointer_array_checks(n))


def _pointer_array_checks(n):
    from bioinfo1_b200 import capi
    L = capi.lib()
    qb, qo, tb, to = synth.short_pairs(78, n, 150)
    exp = CHK.align_batch_full(qb, qo, tb, to, 0, 1, -1, -1, threads=THREADS)
    qptr = (qb.ctypes.data + qo[:n]).astype(np.uint64)
    tptr = (tb.ctypes.data + to[:n]).astype(np.uint64)
    qlen = (qo[1:] - qo[:n]).astype(np.uint32)
    tlen = (to[1:] - to[:n]).astype(np.uint32)
    cap = 64 * n + 64
    score = np.empty(n, dtype=np.int32); tbeg = np.empty(n, dtype=np.uint32)
    cig = np.empty(cap, dtype=np.uint8); coff = np.zeros(n + 1, dtype=np.uint64)
    for _ in range(2):   # second call: recycled staging and plan
        capi.check(L.b200_align_batch(0, n, qptr.ctypes.data, qlen.ctypes.data, tptr.ctypes.data, tlen.ctypes.data, 0, 1, -1, -1,
                                      score.ctypes.data, tbeg.ctypes.data, cig.ctypes.data, coff.ctypes.data, cap))
        _equal((score, tbeg, cig, coff), exp, f"pointer arrays n={n}")
    if n >= 9000:
        # foreign bytes inside a uniform batch: the gather pass (which packs such batches to 2 bits on the host) flags the
        # pairs, uploads their raw bytes and the repair pass aligns them with the byte-compare kernel
        qb2, tb2 = qb.copy(), tb.copy()
        rng = np.random.default_rng(9)
        for i in rng.integers(0, n, size=60):
            qb2[int(i) * 150 + int(rng.integers(0, 150))] = ord("N")
        for i in rng.integers(0, n, size=40):
            tb2[int(i) * 150 + int(rng.integers(0, 150))] = ord("-")
        tb2[(n - 1) * 150 + 149] = ord("a")
        qptr2 = (qb2.ctypes.data + qo[:n]).astype(np.uint64)
        tptr2 = (tb2.ctypes.data + to[:n]).astype(np.uint64)
        for typ in (0, 2):
            exp2 = CHK.align_batch_full(qb2, qo, tb2, to, typ, 1, -1, -1, threads=THREADS)
            capi.check(L.b200_align_batch(0, n, qptr2.ctypes.data, qlen.ctypes.data, tptr2.ctypes.data, tlen.ctypes.data, typ, 1, -1, -1,
                                          score.ctypes.data, tbeg.ctypes.data, cig.ctypes.data, coff.ctypes.data, cap))
            _equal((score, tbeg, cig, coff), exp2, f"pointer arrays with foreign bytes n={n} type={typ}")
        # score only
        capi.check(L.b200_align_batch(0, n, qptr2.ctypes.data, qlen.ctypes.data, tptr2.ctypes.data, tlen.ctypes.data, 2, 1, -1, -1,
                                      score.ctypes.data, tbeg.ctypes.data, None, None, 0))
        assert np.array_equal(score, exp2[0]) and np.array_equal(tbeg, exp2[1])
    # ragged lengths (not a uniform batch), some empty sequences, pointers in scattered order
    rng = np.random.default_rng(3)
    m = min(n, 20_000)
    seqs_q = [bytes(qb[int(qo[i]):int(qo[i]) + int(rng.integers(0, 151))]) for i in range(m)]
    seqs_t = [bytes(tb[int(to[i]):int(to[i]) + int(rng.integers(0, 151))]) for i in range(m)]
    got = capi.align_batch_pointers(0, seqs_q, seqs_t, 2, 1, -1, -1, True)
    pq, pqo = capi.pack(seqs_q); pt, pto = capi.pack(seqs_t)
    es, et, ec, eo = CHK.align_batch_full(pq, pqo, pt, pto, 2, 1, -1, -1, threads=THREADS)
    for i in range(m):
        assert got[i] == (int(es[i]), int(et[i]), bytes(ec[int(eo[i]):int(eo[i + 1])])), i


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_fill_kernel_variants_agree(ctx, typ):
    """Every variant of K1 and K3 (substitution by PRMT / shared table, plain / software-pipelined columns) against the
    checker: unit scores and larger ones. (The context is the test's own.)"""
    n = 20_000
    qb, qo, tb, to = synth.short_pairs(79, n, 150)
    qs, ts = seqgen.ont_like_pairs(31 + typ, 6, mean_len=3000, min_len=2100, max_len=4500)
    lq, lqo = seqgen.pack_arrays(qs); lt, lto = seqgen.pack_arrays(ts)
    for (m, x, g) in ((1, -1, -1), (2, -3, -2), (5, -4, -6)):
        exp_s = CHK.align_batch_full(qb, qo[:4097], tb, to[:4097], typ, m, x, g, threads=THREADS)
        exp_l = CHK.align_batch_full(lq, lqo, lt, lto, typ, m, x, g, threads=min(6, THREADS))
        first = None
        # subst_lds: bit 0 = K1, bit 1 = K3; fill_pipe: K1 pipelined columns
        for lds, pipe in ((3, 1), (3, 0), (0, 1), (0, 0), (2, 1)):
            for key, val in (("subst_lds", lds), ("fill_pipe", pipe)):
                ctx.set_option(key, val)
            what = f"lds={lds} pipe={pipe} scores {(m, x, g)}"
            got = ctx.align_packed(qb, qo, tb, to, typ, m, x, g, True, cigar_cap=80 * n)
            _equal(got, exp_s, "K1 " + what)   # (the checker saw the first 4 096 pairs)
            if first is None:
                first = got
            else:
                _equal(got, first, "K1 against the first variant, every pair: " + what)
            got_l = ctx.align_packed(lq, lqo, lt, lto, typ, m, x, g, True)
            _equal(got_l, exp_l, "K3 " + what)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_stripe_protocol_under_stress(ctx, typ):
    """3 000 three-stripe pairs (4.2-6 kb queries): far more stripes than resident warps, tiny direction budget so the
    batch runs as dozens of waves on two streams, walkers concurrent with the fill and after it -- every variant must
    give the same bytes, and a sample is checked against the CPU."""
    rng = np.random.default_rng(40 + typ)
    base_q, base_t = [], []
    for k in range(40):
        L = int(rng.integers(4200, 6000))
        t = seqgen.random_dna(rng, int(L * rng.uniform(0.3, 1.1)))
        q = seqgen.random_dna(rng, L)
        w = min(len(t), L) // 2
        q[:w] = t[:w]                                   # related prefix, unrelated tail
        base_q.append(q); base_t.append(t)
    qs = [base_q[i % 40] for i in range(3000)]
    ts = [base_t[(i * 7) % 40] for i in range(3000)]   # 280 distinct pairings
    qb, qo = seqgen.pack_arrays(qs); tb, to = seqgen.pack_arrays(ts)
    ref = None
    for cw, budget in ((1, 1 << 28), (0, 1 << 28), (1, 48 << 30)):
        ctx.set_option("concurrent_walk", cw)
        ctx.set_option("dir_budget_bytes", budget)
        got = ctx.align_packed(qb, qo, tb, to, typ, 1, -1, -1, True)
        if ref is None:
            ref = got
            pick = list(range(0, 3000, 431))
            for i in pick:
                e = CHK.align(qs[i].tobytes(), ts[i].tobytes(), typ, 1, -1, -1, True)
                assert (int(got[0][i]), int(got[1][i]), bytes(got[2][int(got[3][i]):int(got[3][i + 1])])) == e, i
        else:
            _equal(got, ref, f"concurrent_walk={cw} budget={budget}")
    ctx.set_option("concurrent_walk", 1)
