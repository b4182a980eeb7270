"""GPU parity tests (bit-exact): every result that crosses the C ABI is compared with the CPU
oracle (oracle/liboracle.so) and with the committed golden vectors from the reference."""
import json
import os
import random

import numpy as np
import pytest

from cpu_checkers import ROOT, load_oracle, load_ref
import seqgen

pytestmark = pytest.mark.gpu
ORACLE = load_oracle()


@pytest.fixture(scope="module")
def ctx():
    from bioinfo1_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _check_batch(ctx, qs, ts, typ, m=1, x=-1, g=-1, checker=ORACLE):
    got = ctx.align(qs, ts, typ, m, x, g, True)
    got_s = ctx.align(qs, ts, typ, m, x, g, False)
    for k, (q, t) in enumerate(zip(qs, ts)):
        exp = checker.align(q, t, typ, m, x, g, True)
        assert got[k] == exp, (k, typ, m, x, g, len(q), len(t), q[:40], t[:40], got[k][:2], exp[:2])
        assert got_s[k][:2] == exp[:2]


def test_golden_vectors_through_the_abi(ctx):
    with open(os.path.join(ROOT, "tests", "golden", "align_golden.json")) as f:
        cases = json.load(f)["cases"]
    groups = {}
    for c in cases:
        groups.setdefault((c["type"], c["match"], c["mismatch"], c["gap"]), []).append(c)
    n = 0
    for (typ, m, x, g), cs in groups.items():
        qs = [bytes.fromhex(c["q"]) for c in cs]
        ts = [bytes.fromhex(c["t"]) for c in cs]
        got = ctx.align(qs, ts, typ, m, x, g, True)
        for c, r in zip(cs, got):
            assert r == (c["score"], c["target_begin"], bytes.fromhex(c["cigar"])), (c["tag"], typ, m, x, g, c["q"], c["t"])
            n += 1
    assert n == len(cases)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_random_batches_vs_oracle(ctx, typ):
    rng = random.Random(100 + typ)
    for ab, (m, x, g) in ((b"ACGT", (1, -1, -1)), (b"ACGTN-", (2, -3, -2)), (b"AC", (1, -1, 1)), (b"ACGTacgt-", (0, -2, -1)),
                          (b"ACGT", (3, 2, -4)), (b"ACGT", (-1, 1, 0))):
        qs = [bytes(rng.choice(ab) for _ in range(rng.randint(0, 90))) for _ in range(300)]
        ts = [bytes(rng.choice(ab) for _ in range(rng.randint(0, 90))) for _ in range(300)]
        _check_batch(ctx, qs, ts, typ, m, x, g)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_related_pairs_across_tile_edges(ctx, typ):
    rng = np.random.default_rng(7 + typ)
    qs, ts = [], []
    for n in (1, 15, 16, 17, 31, 32, 33, 150, 511, 512, 513, 1023, 1024, 1025, 1700):
        t = seqgen.random_dna(rng, n)
        q = seqgen.mutate(rng, t, sub=0.03, ins=0.05, dele=0.05)
        qs.append(q.tobytes()); ts.append(t.tobytes())
        qs.append(t.tobytes()); ts.append(q.tobytes())
    _check_batch(ctx, qs, ts, typ)
    _check_batch(ctx, qs, ts, typ, 2, -3, -2)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_long_pair_multi_stripe(ctx, typ):
    qs, ts = seqgen.ont_like_pairs(11 + typ, 3, mean_len=2500, min_len=1800, max_len=3200)
    qs = [q.tobytes() for q in qs] + [b"A" * 700, b"ACGT" * 300]
    ts = [t.tobytes() for t in ts] + [b"ACGT" * 400, b"A" * 900]
    _check_batch(ctx, qs, ts, typ)


def test_short_pair_batch_config2_shape(ctx):
    """BASELINE config 2 at reduced count: 150 x 150 global, score + CIGAR."""
    qb, qo, tb, to = seqgen.short_pairs(3, 4096)
    score, tbeg, cig, coff = ctx.align_packed(qb, qo, tb, to, 0)
    for k in range(0, 4096, 7):
        q = qb[int(qo[k]):int(qo[k + 1])].tobytes()
        t = tb[int(to[k]):int(to[k + 1])].tobytes()
        exp = ORACLE.align(q, t, 0)
        assert (int(score[k]), int(tbeg[k]), cig[int(coff[k]):int(coff[k + 1])].tobytes()) == exp
    # size-independent property on the full batch: CIGAR op counts rebuild both lengths
    for k in range(4096):
        c = cig[int(coff[k]):int(coff[k + 1])].tobytes().decode()
        num, tot = "", {"M": 0, "I": 0, "D": 0}
        for ch in c:
            if ch.isdigit():
                num += ch
            else:
                tot[ch] += int(num); num = ""
        assert tot["M"] + tot["D"] == 150 and tot["M"] + tot["I"] == 150


def test_multiple_waves_give_identical_results(ctx):
    rng = np.random.default_rng(5)
    qs, ts = [], []
    for _ in range(200):
        t = seqgen.random_dna(rng, int(rng.integers(50, 400)))
        q = seqgen.mutate(rng, t, sub=0.05, ins=0.03, dele=0.03)
        qs.append(q.tobytes()); ts.append(t.tobytes())
    one = ctx.align(qs, ts, 2)
    ctx.set_option("dir_budget_bytes", 1 << 20)
    try:
        many = ctx.align(qs, ts, 2)
    finally:
        ctx.set_option("dir_budget_bytes", 8 << 30)
    assert one == many


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_repair_pass_in_chunks_under_a_small_budget(ctx, typ):
    """Long pairs that are not pure ACGT, with a direction budget of about one matrix: the flagged pairs of every wave are
    repaired by the generic kernel in several chunks that reuse the first slot's buffer."""
    rng = np.random.default_rng(4100 + typ)
    qs, ts = [], []
    for k in range(9):
        t = seqgen.random_dna(rng, int(rng.integers(1500, 2300)))
        q = seqgen.mutate(rng, t, sub=0.04, ins=0.03, dele=0.03)
        if k % 3 != 1:
            q = q.copy(); q[len(q) // 3] = ord("N-n"[k % 3])
        qs.append(q.tobytes()); ts.append(t.tobytes())
    ctx.set_option("dir_budget_bytes", 3 << 20)
    try:
        got = ctx.align(qs, ts, typ)
    finally:
        ctx.set_option("dir_budget_bytes", 48 << 30)
    for q, t, g in zip(qs, ts, got):
        assert g == ORACLE.align(q, t, typ), (typ, len(q), len(t))


def test_pointer_array_entry_point_and_errors(ctx):
    from bioinfo1_b200 import capi
    qs = [b"GTACC", b"", b"ACG", b"TGACGTACATGGACA"]
    ts = [b"GATACGTTA", b"", b"", b"CGTACATGGA"]
    got = capi.align_batch_pointers(0, qs, ts, capi.SEMIGLOBAL)
    assert got == [ORACLE.align(q, t, 2) for q, t in zip(qs, ts)]
    assert got[0] == (2, 0, b"5I1M1I2M2D") and got[1] == (0, 0, b"1\0")
    # too-small CIGAR buffer is an error, not a truncated success
    qb, qo = capi.pack(qs)
    tb, to = capi.pack(ts)
    with pytest.raises(capi.B200Error) as e:
        ctx.align_packed(qb, qo, tb, to, 0, cigar_cap=3)
    assert e.value.code == capi.E_CAP
    with pytest.raises(capi.B200Error) as e:
        ctx.align_packed(qb, qo, tb, to, 5)
    assert e.value.code == capi.E_TYPE
    assert ctx.align([], [], 0) == []


def test_live_reference_spot_check(ctx):
    ref = load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/libref.so not present")
    rng = random.Random(9)
    qs = [bytes(rng.choice(b"ACGT-N") for _ in range(rng.randint(0, 200))) for _ in range(100)]
    ts = [bytes(rng.choice(b"ACGT-N") for _ in range(rng.randint(0, 200))) for _ in range(100)]
    for typ in (0, 1, 2):
        _check_batch(ctx, qs, ts, typ, 2, -1, -2, checker=ref)


# ---- K1: the thread-per-pair int16 kernel (needs >= 8192 eligible pairs to be selected) -------

def _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, m, x, g, stride):
    score, tbeg, cig, coff = ctx.align_packed(qb, qo, tb, to, typ, m, x, g)
    s_only, tb_only, _, _ = ctx.align_packed(qb, qo, tb, to, typ, m, x, g, want_cigar=False)
    assert np.array_equal(score, s_only) and np.array_equal(tbeg, tb_only)
    n = len(qo) - 1
    sc_ref, tb_ref, _ = ORACLE.align_batch(qb, qo, tb, to, typ, m, x, g)
    assert np.array_equal(score, sc_ref), np.nonzero(score != sc_ref)[0][:10]
    assert np.array_equal(tbeg, tb_ref)
    for k in range(0, n, stride):
        q = qb[int(qo[k]):int(qo[k + 1])].tobytes()
        t = tb[int(to[k]):int(to[k + 1])].tobytes()
        exp = ORACLE.align(q, t, typ, m, x, g)
        got = (int(score[k]), int(tbeg[k]), cig[int(coff[k]):int(coff[k + 1])].tobytes())
        assert got == exp, (k, len(q), len(t), got, exp)


def test_short_kernel_uniform_150(ctx):
    qb, qo, tb, to = seqgen.short_pairs(21, 16384)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, 0, 1, -1, -1, 5)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, 0, 2, -3, -2, 37)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, 0, 1, -1, 1, 37)     # positive gap


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_short_kernel_trimmed_last_block(ctx, typ):
    """Uniform batches whose last row block is trimmed to 8 / 16 / 24 rows (and not at all): every pair against the
    oracle's scores, a sample against its CIGARs."""
    for length in (8, 40, 44, 50, 64, 90):
        qb, qo, tb, to = seqgen.short_pairs(300 + length, 8192, length=length)
        _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, 1, -1, -1, 211)


def test_uniform_batch_with_flagged_pairs_goes_through_the_repair_pass(ctx):
    """Uniform (device-built plan) K1 batch in which a few pairs are not pure ACGT: they are flagged, skipped by the wave
    and repaired by the generic kernel; a clean batch right after must not see the patched descriptors."""
    qb, qo, tb, to = seqgen.short_pairs(5150, 16384)
    qb = qb.copy(); tb = tb.copy()
    for k in (0, 63, 64, 777, 9000, 16383):
        qb[int(qo[k]) + 7] = ord("N")
    tb[int(to[4242]) + 3] = ord("-")
    ctx.set_option("chunk_pairs", 4096)            # several waves, so the host path pipelines its downloads
    try:
        _check_packed_vs_oracle(ctx, qb, qo, tb, to, 0, 1, -1, -1, 97)
        qb2, qo2, tb2, to2 = seqgen.short_pairs(5151, 16384)
        _check_packed_vs_oracle(ctx, qb2, qo2, tb2, to2, 0, 1, -1, -1, 97)
    finally:
        ctx.set_option("chunk_pairs", 0)


def test_short_kernel_ragged_lengths_and_fallback(ctx):
    rng = np.random.default_rng(77)
    n = 12000
    qs, ts = [], []
    for k in range(n):
        T = int(rng.integers(0, 260))
        t = seqgen.random_dna(rng, T)
        if k % 3 == 0:
            q = seqgen.random_dna(rng, int(rng.integers(0, 260)))
        else:
            q = seqgen.mutate(rng, t, sub=0.05, ins=0.04, dele=0.04)
        if k % 97 == 0 and len(q) > 3:      # not pure ACGT: must fall back to the generic kernel
            q = q.copy(); q[len(q) // 2] = ord("N")
        if k % 389 == 0 and T > 3:
            t = t.copy(); t[1] = ord("-")
        qs.append(q); ts.append(t)
    qb, qo = seqgen.pack_arrays(qs)
    tb, to = seqgen.pack_arrays(ts)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, 0, 1, -1, -1, 11)
    # same plan shape, now all-ACGT content: the fallback patch must be undone
    qs2 = [np.where(np.isin(q, seqgen.ACGT), q, ord("A")).astype(np.uint8) for q in qs]
    ts2 = [np.where(np.isin(t, seqgen.ACGT), t, ord("C")).astype(np.uint8) for t in ts]
    qb2, qo2 = seqgen.pack_arrays(qs2)
    tb2, to2 = seqgen.pack_arrays(ts2)
    _check_packed_vs_oracle(ctx, qb2, qo2, tb2, to2, 0, 3, -2, -4, 11)


def test_cpp_dropin_wrappers():
    """tests/cpp/dropin_test.cpp: a caller written against the reference headers, linked to our library."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cpp", "dropin_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "dropin_test: OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_generic_kernel_forced_on_plain_dna(ctx, typ):
    """The int32 byte-compare kernel is the fallback of every fast path: keep it covered on ACGT input."""
    rng = np.random.default_rng(31 + typ)
    qs, ts = [], []
    for n in (40, 300, 700, 1100):
        t = seqgen.random_dna(rng, n)
        qs.append(seqgen.mutate(rng, t, sub=0.04, ins=0.04, dele=0.04).tobytes())
        ts.append(t.tobytes())
    ctx.set_option("force_generic", 1)
    try:
        _check_batch(ctx, qs, ts, typ)
    finally:
        ctx.set_option("force_generic", 0)
    _check_batch(ctx, qs, ts, typ)   # and the fast path on the same input


# ---- K3: packed int16x2 long-pair kernel (2048-row stripes, per-block bases) --------------------

@pytest.mark.parametrize("typ", [0, 1, 2])
def test_long16_stripe_edges_and_score_range(ctx, typ):
    """Lengths straddle the 32/64-row blocks and the 2048-row stripes; score sets run from the unit scores to
    the largest |4(s-gap)+1| the byte tables take, where the 16-bit spread bound is tightest."""
    rng = np.random.default_rng(500 + typ)
    qs, ts = [], []
    for n in (31, 33, 63, 64, 65, 127, 129, 2047, 2048, 2049, 2111, 4100, 6500):
        t = seqgen.random_dna(rng, n)
        q = seqgen.mutate(rng, t, sub=0.03, ins=0.05, dele=0.05)
        qs.append(q.tobytes()); ts.append(t.tobytes())
        qs.append(t.tobytes()[: max(1, n // 3)]); ts.append(q.tobytes())
    qs += [b"A" * 2500, b"ACGT" * 600, b"G" * 70]
    ts += [b"A" * 2300, b"A" * 900, b"G" * 3000]
    _check_batch(ctx, qs, ts, typ)
    for m, x, g in ((2, -3, -2), (10, -10, -10), (15, -16, -16), (25, -6, -6), (1, -1, 1), (0, 0, 0), (-2, 3, -1)):
        _check_batch(ctx, qs[:12] + qs[-3:], ts[:12] + ts[-3:], typ, m, x, g)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_long16_large_scores_leave_int16(ctx, typ):
    """Identical 9 kb sequences: H reaches 9000 (4H > 32767), so only the per-block bases keep the halves in
    range, and the local clamp sits far below the representable window."""
    rng = np.random.default_rng(600 + typ)
    t = seqgen.random_dna(rng, 9000).tobytes()
    q2 = seqgen.mutate(rng, np.frombuffer(t, dtype=np.uint8), sub=0.02, ins=0.01, dele=0.01).tobytes()
    _check_batch(ctx, [t, q2, t[:5000]], [t, t, q2], typ)
    _check_batch(ctx, [t, q2], [t, t], typ, 3, -2, -4)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_long_pairs_positive_gap_scores(ctx, typ):
    """A positive gap score makes cells next to a zero border grow by `gap` per row: the 16-bit kernel's spread
    bound only holds for global alignments there, the other two types must be planned on the int32 kernel
    (found by tools/fuzz_gpu.py)."""
    rng = np.random.default_rng(700 + typ)
    qs, ts = [], []
    for n in (2402, 5088):
        t = seqgen.random_dna(rng, n)
        qs.append(seqgen.mutate(rng, t, sub=0.05, ins=0.06, dele=0.06).tobytes()); ts.append(t.tobytes())
    qs.append(seqgen.random_dna(rng, 2402).tobytes()); ts.append(seqgen.random_dna(rng, 4839).tobytes())
    for m, x, g in ((12, -6, 2), (-3, 1, 2), (1, -1, 1)):
        _check_batch(ctx, qs, ts, typ, m, x, g)


@pytest.mark.parametrize("typ", [0, 1, 2])
def test_long32_kernel_kept_covered(ctx, typ):
    """The int32 stripe kernel serves score sets the 16-bit bound rejects: keep it covered on the same inputs."""
    qs, ts = seqgen.ont_like_pairs(40 + typ, 3, mean_len=2500, min_len=1800, max_len=3200)
    qs = [q.tobytes() for q in qs] + [b"A" * 700, b"ACGT" * 300]
    ts = [t.tobytes() for t in ts] + [b"ACGT" * 400, b"A" * 900]
    ctx.set_option("long16", 0)
    try:
        _check_batch(ctx, qs, ts, typ)
    finally:
        ctx.set_option("long16", 1)
    # 4(s-gap)+1 fits the byte tables, but one vertical step moves Y by 803: the 16-bit bound rejects it
    _check_batch(ctx, qs, ts, typ, -170, -200, -200)


@pytest.mark.parametrize("typ", [1, 2])
def test_short_kernel_semi_global_and_local(ctx, typ):
    """K1 on the other two alignment types: end-cell rules (last column before last row, first maximum in
    row-major order) inside the thread-per-pair kernel, uniform and ragged batches, with fallbacks."""
    qb, qo, tb, to = seqgen.short_pairs(23 + typ, 16384)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, 1, -1, -1, 5)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, 2, -3, -2, 37)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, 0, -1, -1, 37)      # zero match score: ties everywhere
    rng = np.random.default_rng(80 + typ)
    qs, ts = [], []
    for k in range(12000):
        T = int(rng.integers(0, 260))
        t = seqgen.random_dna(rng, T)
        if k % 3 == 0:
            q = seqgen.random_dna(rng, int(rng.integers(0, 260)))
        elif k % 3 == 1:
            q = seqgen.mutate(rng, t, sub=0.05, ins=0.04, dele=0.04)
        else:   # overhangs on either side: what semi-global and local are for
            a, b = sorted(int(x) for x in rng.integers(0, T + 1, size=2))
            q = np.concatenate([seqgen.random_dna(rng, int(rng.integers(0, 40))), t[a:b], seqgen.random_dna(rng, int(rng.integers(0, 40)))])
        if k % 97 == 0 and len(q) > 3:
            q = q.copy(); q[len(q) // 2] = ord("N")
        qs.append(q); ts.append(t)
    qb, qo = seqgen.pack_arrays(qs)
    tb, to = seqgen.pack_arrays(ts)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, 1, -1, -1, 7)
    _check_packed_vs_oracle(ctx, qb, qo, tb, to, typ, 3, -2, -4, 13)


# ---- full-size properties (BASELINE.json sizes; the oracle only sees a sample) ---------------------

def test_config2_full_size_two_implementations_agree(ctx):
    """1 048 576 pairs of 150 x 150 (BASELINE config 2): the int16x2 thread-per-pair kernel and the int32
    byte-compare warp kernel are independent implementations; every score, target_begin and CIGAR byte of the
    full batch must agree, and a sample is checked against the oracle."""
    n = 1 << 20
    qb, qo, tb, to = seqgen.short_pairs(1000, n)
    fast = ctx.align_packed(qb, qo, tb, to, 0)
    ctx.set_option("force_generic", 1)
    try:
        slow = ctx.align_packed(qb, qo, tb, to, 0)
    finally:
        ctx.set_option("force_generic", 0)
    assert np.array_equal(fast[0], slow[0]) and np.array_equal(fast[1], slow[1])
    assert np.array_equal(fast[3], slow[3])
    total = int(fast[3][-1])
    assert np.array_equal(fast[2][:total], slow[2][:total])
    for k in range(0, n, 9973):
        q = qb[int(qo[k]):int(qo[k + 1])].tobytes(); t = tb[int(to[k]):int(to[k + 1])].tobytes()
        assert (int(fast[0][k]), int(fast[1][k]), fast[2][int(fast[3][k]):int(fast[3][k + 1])].tobytes()) == ORACLE.align(q, t, 0)


@pytest.mark.parametrize("typ", [1, 2])
def test_ont_like_long_pairs_two_implementations_agree(ctx, typ):
    """ONT-like pairs of BASELINE configs 4/5 shape (8 kb mean, 12 % indel-heavy error; 10 kb local): the packed
    int16x2 stripe kernel against the int32 stripe kernel on every pair, the oracle on two of them."""
    qs, ts = seqgen.ont_like_pairs(900 + typ, 48, mean_len=8000) if typ == 2 else seqgen.ont_like_pairs(900 + typ, 24, fixed=10000)
    qb, qo = seqgen.pack_arrays(qs)
    tb, to = seqgen.pack_arrays(ts)
    a = ctx.align_packed(qb, qo, tb, to, typ)
    ctx.set_option("long16", 0)
    try:
        b = ctx.align_packed(qb, qo, tb, to, typ)
    finally:
        ctx.set_option("long16", 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    total = int(a[3][-1])
    assert np.array_equal(a[2][:total], b[2][:total])
    for k in (0, len(qs) - 1):
        exp = ORACLE.align(qs[k].tobytes(), ts[k].tobytes(), typ)
        assert (int(a[0][k]), int(a[1][k]), a[2][int(a[3][k]):int(a[3][k + 1])].tobytes()) == exp
