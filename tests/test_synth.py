"""CPU tests of the seeded input generator (tests/synth.c) and of the checkers' batch entry points."""
import hashlib

import numpy as np

from cpu_checkers import load_oracle, load_ref
import synth


def _h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_generator_output_is_pinned():
    """The bytes are a function of the seeds alone (digests committed here), whatever the thread count or slice."""
    ref = synth.dna(1, 100_000)
    b, o = synth.ont_reads(2, ref, n=5, mean=2000.0, lo=500, hi=5000)
    q, qo, t, to = synth.short_pairs(7, 4)
    lq, _, lt, _ = synth.long_pairs(5, 1, 1000)
    assert [_h(x) for x in (ref, b, o, q, t, lq, lt)] == [
        "f2a0942f7bb35000", "a851a767c62f76d0", "95befabdd8a9879c", "172526ca45614fc2", "5ecba9d577565cf9",
        "db930aa79f1779d0", "6dafe4a9c8ce3869"]
    assert list(synth.ont_spans(2, 6)) == [3833, 4035, 4943, 4935, 6439, 5332]


def test_any_subset_can_be_generated_alone():
    ref = synth.dna(1, 50_000)
    buf, off = synth.ont_reads(3, ref, n=300, mean=1500.0, lo=300, hi=4000)
    pick = [0, 7, 8, 9, 150, 299]
    b2, o2 = synth.ont_reads(3, ref, idx=pick, mean=1500.0, lo=300, hi=4000)
    for k, i in enumerate(pick):
        assert bytes(buf[int(off[i]):int(off[i + 1])]) == bytes(b2[int(o2[k]):int(o2[k + 1])])
    q, qo, t, to = synth.short_pairs(11, 1000)
    q2, _, t2, _ = synth.short_pairs(11, 10, first=500)
    assert bytes(q[500 * 150:510 * 150]) == bytes(q2[:1500]) and bytes(t[500 * 150:510 * 150]) == bytes(t2[:1500])
    assert bytes(synth.dna(1, 50_000)[777:999]) == bytes(ref[777:999])


def test_reads_are_erroneous_copies_of_the_reference():
    ref = synth.dna(1, 200_000)
    buf, off, start, strand, span = synth.ont_reads(2, ref, n=50, with_truth=True)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    assert lens.min() >= 800 and lens.max() <= 45_000
    assert 0.35 < strand.mean() < 0.65
    # the oracle's semi-global alignment of a forward read onto its source window is mostly matches
    orc = load_oracle()
    i = int(np.nonzero(strand)[0][0])
    q = bytes(buf[int(off[i]):int(off[i + 1])])[:1500]
    t = bytes(ref[int(start[i]):int(start[i]) + 1700])
    score, _, _ = orc.align(q, t, 2, 1, -1, -1, False)
    assert score > 0.55 * len(q)


def test_checker_batch_forms_agree_with_the_single_calls():
    orc, ref = load_oracle(), load_ref()
    q, qo, t, to = synth.short_pairs(21, 64, 90)
    for chk in (c for c in (orc, ref) if c is not None):
        for typ in (0, 1, 2):
            s, tb, cig, coff = chk.align_batch_full(q, qo, t, to, typ, 2, -3, -2, threads=3)
            for i in (0, 1, 31, 63):
                exp = chk.align(bytes(q[int(qo[i]):int(qo[i + 1])]), bytes(t[int(to[i]):int(to[i + 1])]), typ, 2, -3, -2, True)
                assert (int(s[i]), int(tb[i]), bytes(cig[int(coff[i]):int(coff[i + 1])])) == exp
    rf = synth.dna(1, 30_000)
    buf, off = synth.ont_reads(4, rf, n=6, mean=900.0, lo=200, hi=2000)
    for chk in (c for c in (orc, ref) if c is not None):
        h, p, f, oo = chk.minimize_batch(buf, off, 15, 5, True)
        for i in range(6):
            eh, ep, ef = chk.minimize(bytes(buf[int(off[i]):int(off[i + 1])]), 15, 5, True)
            a, b = int(oo[i]), int(oo[i + 1])
            assert np.array_equal(h[a:b], eh) and np.array_equal(p[a:b], ep) and np.array_equal(f[a:b], ef)
    if ref is not None:   # the port and the reference agree on the batch forms too
        a = orc.align_batch_full(q, qo, t, to, 1, 1, -1, -1, threads=2)
        b = ref.align_batch_full(q, qo, t, to, 1, 1, -1, -1, threads=2)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
