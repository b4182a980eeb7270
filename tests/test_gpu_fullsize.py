"""GPU parity at BASELINE.md section 4's sample sizes, against the UNMODIFIED reference (oracle/_ref, prebuilt;
the C port when that is absent), all through the C ABI:

  config 2   ALL 1 048 576 pairs 150 x 150, global: score, target_begin, every CIGAR byte
  config 3   2 000 ONT-like reads: every minimizer tuple (k = 15, w = 5)
  config 4   64 ONT-like reads against the 4.6 Mbp reference: PAF of b200_mapper vs the reference mapper, byte for byte
  config 5   8 local pairs 10 kb x 10 kb: score, target_begin, full CIGAR

Inputs come from tests/synth.c (seeded splitmix64 streams). The CPU side takes ~2 minutes on the box's host cores.
"""
import os
import subprocess

import numpy as np
import pytest

from cpu_checkers import ROOT, load_oracle, load_ref
import synth

pytestmark = pytest.mark.gpu
REF = load_ref()
CHK = REF or load_oracle()
THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def ctx():
    from bioinfo1_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _same_alignments(got, exp, what):
    gs, gt, gc, go = got
    es, et, ec, eo = exp
    n = len(es)
    bad = np.nonzero(gs[:n] != es)[0]
    assert bad.size == 0, f"{what}: {bad.size} scores differ, first at pair {bad[0]}: {gs[bad[0]]} vs {es[bad[0]]}"
    bad = np.nonzero(gt[:n] != et)[0]
    assert bad.size == 0, f"{what}: {bad.size} target_begin differ, first at pair {bad[0]}"
    bad = np.nonzero(go[:n + 1] != eo)[0]
    assert bad.size == 0, f"{what}: CIGAR lengths differ from pair {bad[0] - 1} on"
    total = int(eo[n])
    assert np.array_equal(gc[:total], ec[:total]), f"{what}: CIGAR bytes differ"


def test_config2_every_pair_against_the_reference(ctx):
    n = 1 << 20
    qb, qo, tb, to = synth.short_pairs(1000, n, 150)
    got = ctx.align_packed(qb, qo, tb, to, 0, 1, -1, -1, True, cigar_cap=64 * n)
    exp = CHK.align_batch_full(qb, qo, tb, to, 0, 1, -1, -1, threads=THREADS)
    _same_alignments(got, exp, "config 2")
    # the other two alignment types of the same kernel on a slice of the batch
    m = 1 << 16
    for typ in (1, 2):
        got = ctx.align_packed(qb, qo[:m + 1], tb, to[:m + 1], typ, 1, -1, -1, True, cigar_cap=64 * m)
        exp = CHK.align_batch_full(qb, qo[:m + 1], tb, to[:m + 1], typ, 1, -1, -1, threads=THREADS)
        _same_alignments(got, exp, f"config 2 shape, type {typ}")


def test_config3_minimizer_tuples_of_2000_reads(ctx):
    ref = synth.dna(1, 4_600_000)
    buf, off = synth.ont_reads(2, ref, n=2000)
    h, p, f, oo = ctx.minimize_packed(buf, off, 15, 5)
    eh, ep, ef, eo = CHK.minimize_batch(buf, off, 15, 5, True)
    assert np.array_equal(oo, eo)
    assert np.array_equal(h, eh) and np.array_equal(p, ep) and np.array_equal(f, ef)
    # the reference itself, both strands, through the same entry point (flags true / false)
    sl = 200_000
    piece = np.concatenate([ref[:sl], ref[:sl][::-1], np.zeros(1, np.uint8)])
    off2 = np.array([0, sl, 2 * sl], dtype=np.uint64)
    h, p, f, oo = ctx.minimize_packed(piece, off2, 15, 5, is_fwd=[1, 0])
    for s, fw in ((0, True), (1, False)):
        eh, ep, ef, _ = CHK.minimize_batch(piece, off2[s:s + 2], 15, 5, fw)
        a, b = int(oo[s]), int(oo[s + 1])
        assert np.array_equal(h[a:b], eh) and np.array_equal(p[a:b], ep) and np.array_equal(f[a:b], ef)


def _write_fasta(path, names, buf, off, fastq=False):
    with open(path, "wb") as fh:
        for i, nm in enumerate(names):
            seq = bytes(buf[int(off[i]):int(off[i + 1])])
            fh.write((b"@" if fastq else b">") + nm.encode() + b"\n" + seq + b"\n")
            if fastq:
                fh.write(b"+\n" + b"I" * len(seq) + b"\n")


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_mapper")), reason="prebuilt reference mapper absent")
def test_config4_paf_of_64_reads_matches_the_reference_mapper(tmp_path):
    ref = synth.dna(1, 4_600_000)
    buf, off = synth.ont_reads(2, ref, n=64)
    # FASTQ input: the reference then looks both strands up in full (team_mapper.cpp:710-790), so every read is a
    # full-length semi-global alignment (FASTA input is covered by the golden PAFs of test_gpu_mapper.py)
    refp, readp = str(tmp_path / "ref.fa"), str(tmp_path / "reads.fq")
    _write_fasta(refp, ["synthetic_ecoli_sized"], ref, np.array([0, 4_600_000], dtype=np.uint64))
    _write_fasta(readp, [f"read{i}" for i in range(64)], buf, off, fastq=True)
    argv = ["-a", "semiGlobal", "-c", "-f", "0", "-k", "15", "-w", "5", refp, readp]
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    mine = subprocess.run([exe] + argv, capture_output=True, timeout=600)
    assert mine.returncode == 0, mine.stderr.decode(errors="replace")
    theirs = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_mapper")] + argv, capture_output=True, timeout=1500)
    assert theirs.returncode == 0, theirs.stderr.decode(errors="replace")
    # the reference prints its statistics on stdout too; PAF lines are the ones with 12+ tab-separated fields
    paf = b"".join(ln + b"\n" for ln in theirs.stdout.split(b"\n") if ln.count(b"\t") >= 11)
    assert mine.stdout.count(b"\n") >= 60          # ONT-like reads of a random reference map
    assert mine.stdout == paf


def test_config5_eight_local_pairs_10kb_full_cigar(ctx):
    qb, qo, tb, to = synth.long_pairs(5, 8, 10000)
    got = ctx.align_packed(qb, qo, tb, to, 1, 1, -1, -1, True)
    exp = CHK.align_batch_full(qb, qo, tb, to, 1, 1, -1, -1, threads=min(4, THREADS))
    _same_alignments(got, exp, "config 5")
    # and the semi-global form of the same pairs (config 4's Align step at its full length)
    got = ctx.align_packed(qb, qo[:5], tb, to[:5], 2, 1, -1, -1, True)
    exp = CHK.align_batch_full(qb, qo[:5], tb, to[:5], 2, 1, -1, -1, threads=min(4, THREADS))
    _same_alignments(got, exp, "config 5 shape, semiGlobal")
