#!/usr/bin/env python3
"""Randomised differential run of the mapper pipeline (index, de-dup, seeds, chains, region, Align) against the
CPU mapper oracle. Usage: python tools/fuzz_gpu_mapper.py [seconds] [seed]"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import seqgen
import mapper_oracle
from bioinfo1_b200 import capi
from cpu_checkers import load_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
O = load_oracle(); ctx = capi.Context(0)
pr = random.Random(seed); rng = np.random.default_rng(seed)
comp = bytes.maketrans(b"ACGT", b"TGCA")
t_end = time.time() + budget
rounds = cases = mapped = 0
while time.time() < t_end:
    rounds += 1
    k = pr.choice([7, 10, 12, 15, 15, 16]); w = pr.choice([2, 3, 5, 5, 8])
    f = pr.choice([0.0, 0.0, 0.001, 0.01])
    typ = pr.randrange(3); m, x, g = pr.choice([(1, -1, -1), (2, -3, -2), (3, -1, -4)])
    n_ref = pr.choice([3000, 20000, 60000])
    ref = seqgen.random_dna(rng, n_ref)
    if pr.random() < 0.5:      # a repeat, so that some minimizers occur more than once
        a = pr.randrange(n_ref // 2); ln = pr.randrange(50, 600); b = pr.randrange(n_ref // 2, n_ref - ln)
        ref[b:b + ln] = ref[a:a + ln]
    ref = ref.tobytes()
    reads = []
    for i in range(pr.randint(1, 12)):
        L = pr.randint(30, min(2500, n_ref - 1)); s = pr.randrange(0, n_ref - L)
        q = seqgen.mutate(rng, np.frombuffer(ref[s:s + L], dtype=np.uint8), sub=0.03, ins=0.03, dele=0.03).tobytes()
        if pr.random() < 0.15:
            q = seqgen.random_dna(rng, L).tobytes()
        reads.append(q.translate(comp)[::-1] if pr.random() < 0.4 else q)
    fastq = pr.random() < 0.5
    idx = capi.Index(ctx, ref, k, w, f)
    oidx = mapper_oracle.Index(O, ref, k, w, f)
    try:
        res, cigs = idx.map_batch(reads, fastq, typ, m, x, g, True)
    finally:
        idx.close()
    for i, rd in enumerate(reads):
        exp = mapper_oracle.map_read(O, oidx, rd, k, w, typ, m, x, g, True, fasta_path=not fastq)
        cases += 1
        if exp is None:
            ok = not res[i]["mapped"]
        else:
            mapped += 1
            got = dict(q_begin=int(res[i]["q_begin"]), q_end=int(res[i]["q_end"]), fwd=bool(res[i]["strand_fwd"]),
                       t_begin=int(res[i]["t_begin"]), t_end=int(res[i]["t_end"]), score=int(res[i]["score"]), cigar=cigs[i])
            ok = bool(res[i]["mapped"]) and got == exp
        if not ok:
            print("MISMATCH", dict(seed=seed, round=rounds, k=k, w=w, f=f, typ=typ, scores=(m, x, g), fastq=fastq, read=i, L=len(rd)))
            sys.exit(1)
print(f"mapper fuzz ok: {rounds} rounds, {cases} reads ({mapped} mapped) checked against the CPU mapper oracle, seed {seed}")
