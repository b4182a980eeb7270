#!/usr/bin/env python3
"""Randomised differential run of MinimizeBatch against the CPU oracle. Usage: python tools/fuzz_gpu_minimize.py [seconds] [seed]"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioinfo1_b200 import capi
from cpu_checkers import load_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
O = load_oracle(); ctx = capi.Context(0)
pr = random.Random(seed)
t_end = time.time() + budget
rounds = cases = 0
while time.time() < t_end:
    rounds += 1
    k = pr.choice([1, 2, 3, 5, 8, 11, 14, 15, 16, 17, 19, 24, 31])
    w = pr.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 20])
    ab = pr.choice([b"ACGT", b"ACGT", b"ACGTN", b"GGGGGGGT", b"ACGTacgt-", b"AC"])
    seqs, fw = [], []
    for _ in range(pr.randint(1, 30)):
        hi = pr.choice([40, 300, 3000, 20000])
        L = pr.randint(max(0, k + w - 3), max(k + w - 3, hi))
        seqs.append(bytes(pr.choice(ab) for _ in range(L)) if L < 4000 else bytes(np.frombuffer(ab, dtype=np.uint8)[np.random.default_rng(pr.randrange(1 << 30)).integers(0, len(ab), size=L)]))
        fw.append(pr.randint(0, 1))
    got = ctx.minimize(seqs, k, w, fw)
    for s, f, g in zip(seqs, fw, got):
        if len(s) < k + w - 3 and len(s) >= k:
            continue
        e = O.minimize(s, k, w, bool(f))
        cases += 1
        if not all(np.array_equal(a, b) for a, b in zip(g, e)):
            print("MISMATCH", dict(seed=seed, round=rounds, k=k, w=w, L=len(s), fwd=f, alphabet=ab))
            sys.exit(1)
print(f"minimize fuzz ok: {rounds} rounds, {cases} sequences checked against the oracle, seed {seed}")
