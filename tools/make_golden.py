#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref/libref.so,
built by oracle/Makefile from /root/reference/team_alignment/team_alignment.cpp and
/root/reference/team_minimizers/team_minimizers.cpp).

Run in the authoring container only (the reference tree is not on the GPU box):
    make -C oracle && python tools/make_golden.py
The vectors are committed so that the CPU and GPU suites can check parity without the
reference being present.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpu_checkers import load_ref, TYPE_NAMES  # noqa: E402


def rand_seq(rng, n, alphabet):
    return bytes(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, err, alphabet=b"ACGT"):
    out = bytearray()
    for c in s:
        r = rng.random()
        if r < err * 0.4:
            continue                      # deletion
        if r < err * 0.8:
            out.append(rng.choice(alphabet))  # insertion before
            out.append(c)
        elif r < err:
            out.append(rng.choice(alphabet))  # substitution
        else:
            out.append(c)
    return bytes(out)


def main():
    ref = load_ref()
    if ref is None:
        sys.exit("oracle/_ref/libref.so missing: run `make -C oracle` where /root/reference exists")
    rng = random.Random(20261018)
    align_cases = []

    def add(q, t, m=1, x=-1, g=-1, tag=""):
        for typ in (0, 1, 2):
            s, tb, cg = ref.align(q, t, typ, m, x, g, True)
            s2, tb2, _ = ref.align(q, t, typ, m, x, g, False)
            assert (s, tb) == (s2, tb2)
            align_cases.append(dict(tag=tag, q=q.hex(), t=t.hex(), type=typ, match=m, mismatch=x, gap=g,
                                    score=s, target_begin=tb, cigar=cg.hex()))

    # the bundled example pairs (BASELINE.json config 1) and the SURVEY.md section-4 table
    toys = [(b"GTACC", b"GATACGTTA"), (b"ACGTACGTAA", b"ACGTCGTTAA"), (b"GTACC", b"TTCACGTTA"),
            (b"GATCATATT", b"TCGTAGCG"), (b"TGACGTACATGGACA", b"CGTACATGGA"),
            (b"CGTACATGGA", b"TGACGTACATGGACA"), (b"", b""), (b"", b"ACG"), (b"ACG", b""),
            (b"AAAA", b"CCCC"), (b"A-CG", b"ACG"), (b"acgt", b"ACGT"), (b"AA", b"A"), (b"A", b"AA")]
    for q, t in toys:
        add(q, t, tag="toy")
    add(b"ACGTACGTAA", b"ACGTCGTTAA", 2, -3, -2, tag="toy-scores")
    add(b"GTACGT", b"ACGTACGAC", 2, -1, 2, tag="deck-positive-gap")

    # random small cases over awkward alphabets and score signs
    alphabets = [b"ACGT", b"AC", b"ACGTN", b"ACGT-", b"ACgt-N", b"A"]
    for _ in range(400):
        ab = rng.choice(alphabets)
        q = rand_seq(rng, rng.randint(0, 24), ab)
        t = rand_seq(rng, rng.randint(0, 24), ab)
        m, x, g = rng.randint(-3, 4), rng.randint(-4, 3), rng.randint(-4, 3)
        add(q, t, m, x, g, tag="rand-small")
    # related sequences at realistic error, sizes crossing the 16/32/512-row tiling edges
    for n in (15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 150, 255, 256, 257, 300, 511, 512, 513, 700, 1030):
        t = rand_seq(rng, n, b"ACGT")
        q = mutate(rng, t, 0.12)
        add(q, t, tag="related")
        add(t, q, 2, -3, -2, tag="related-swapped")
    # unequal shapes
    for (a, b) in ((1, 600), (600, 1), (40, 900), (900, 40), (530, 70), (70, 530)):
        q = rand_seq(rng, a, b"ACGT")
        t = rand_seq(rng, b, b"ACGT")
        add(q, t, tag="skinny")
    with open(os.path.join(ROOT, "tests", "golden", "align_golden.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py", source="reference team::Align via oracle/_ref/libref.so",
                       types=TYPE_NAMES, cases=align_cases), f, separators=(",", ":"))

    min_cases = []

    def addm(seq, k, w, fwd=True, tag=""):
        h, p, fl = ref.minimize(seq, k, w, fwd)
        min_cases.append(dict(tag=tag, seq=seq.hex(), k=k, w=w, fwd=fwd, hash=[int(v) for v in h],
                              pos=[int(v) for v in p], flag=[int(v) for v in fl]))

    ref84 = (b"AATCGTGACGTACATGGACAGCTTACGGTACATGGAGGCGTACATGGACAAGCTTGACGTACATGGACATTTGGCGTACATGGA")
    addm(b"TGACGTACATGGACA", 3, 3, True, "toy")
    addm(b"TGACGTACATGGACA", 3, 4, True, "toy")
    addm(b"TGACGTACATGGACA", 3, 4, False, "toy")
    addm(b"ACGTACGAC", 3, 3, True, "toy")
    addm(b"ACGTACGAC", 3, 1, True, "toy")
    addm(b"ACG", 3, 3, True, "short-nul")
    addm(b"G" * 20, 16, 3, True, "sentinel")
    addm(b"ACGTACGTACGTACGTACGT", 17, 2, True, "k17")
    addm(b"ACNNacgtACGT", 3, 2, True, "non-acgt")
    addm(ref84, 15, 5, True, "reference.fasta")
    addm(ref84, 15, 5, False, "reference.fasta")
    addm(ref84, 5, 3, True, "reference.fasta")
    for _ in range(300):
        k = rng.randint(1, 20)
        w = rng.randint(1, 9)
        L = rng.randint(max(0, k + w - 3), 90)   # parity is defined for L >= k+w-3
        ab = rng.choice([b"ACGT", b"ACGTN", b"G", b"GGGT", b"ACgt"])
        addm(rand_seq(rng, L, ab), k, w, rng.random() < 0.5, "rand")
    for L in (1000, 4099):
        addm(rand_seq(rng, L, b"ACGT"), 15, 5, True, "long")
        addm(rand_seq(rng, L, b"ACGT"), 19, 10, False, "long")
    with open(os.path.join(ROOT, "tests", "golden", "minimize_golden.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py",
                       source="reference team::KMER::Minimize via oracle/_ref/libref.so", cases=min_cases),
                  f, separators=(",", ":"))
    print(len(align_cases), "align cases,", len(min_cases), "minimize cases")


if __name__ == "__main__":
    main()
