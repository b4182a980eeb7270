#!/usr/bin/env python3
"""Randomised differential run of every alignment kernel against the CPU oracle (test infrastructure, like tests/):
random lengths, alphabets, score sets and kernel routes. Usage: python tools/fuzz_gpu.py [seconds] [seed]"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import seqgen
from bioinfo1_b200 import capi
from cpu_checkers import load_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
O = load_oracle(); ctx = capi.Context(0)
pr = random.Random(seed); rng = np.random.default_rng(seed)
t_end = time.time() + budget
rounds = cases = 0
while time.time() < t_end:
    rounds += 1
    typ = pr.randrange(3)
    m = pr.randint(-3, 12); x = pr.randint(-12, 3); g = pr.randint(-12, 2)
    route = pr.choice(["default", "default", "long32", "generic"])
    shape = pr.choice(["tiny", "mid", "long", "k1", "k1u"])
    # kernel variants and host-pipeline schedules (round 2): every combination must give the reference's bytes
    opts = {"subst_lds": pr.choice([0, 1, 2, 3]), "fill_pipe": pr.choice([0, 1]), "stream_fill": pr.choice([0, 1]),
            "taper_tail": pr.choice([0, 1, 2]), "chunk_pairs": pr.choice([0, 0, 1024, 2048])}
    for k_, v_ in opts.items():
        ctx.set_option(k_, v_)
    qs, ts = [], []
    if shape == "k1u":     # a UNIFORM batch: the device-built plan, several waves, pipelined downloads, streaming fill
        n = 8192 + pr.randrange(9000)
        Lq, Lt = pr.randint(1, 160), pr.randint(1, 160)
        for k in range(n):
            t = seqgen.random_dna(rng, Lt)
            q = seqgen.fixed_len(rng, seqgen.mutate(rng, t, sub=0.06, ins=0.03, dele=0.03), Lq) if k % 3 else seqgen.random_dna(rng, Lq)
            if pr.random() < 0.001:
                q = q.copy(); q[pr.randrange(len(q))] = ord(pr.choice("N-a"))
            qs.append(q.tobytes()); ts.append(t.tobytes())
    elif shape == "k1":      # >= 8192 short pairs: the thread-per-pair kernel (if the scores allow)
        n = 8192 + pr.randrange(700)
        for k in range(n):
            T = int(rng.integers(0, 120)); t = seqgen.random_dna(rng, T)
            q = seqgen.mutate(rng, t, sub=0.06, ins=0.05, dele=0.05) if k % 2 else seqgen.random_dna(rng, int(rng.integers(0, 120)))
            qs.append(q.tobytes()); ts.append(t.tobytes())
    else:
        n = {"tiny": 40, "mid": 12, "long": 3}[shape]
        hi = {"tiny": 70, "mid": 900, "long": 5200}[shape]
        for k in range(n):
            T = int(rng.integers(0, hi)); t = seqgen.random_dna(rng, T)
            if pr.random() < 0.7:
                q = seqgen.mutate(rng, t, sub=0.05, ins=0.06, dele=0.06)
                if pr.random() < 0.3 and len(q) > 10:
                    a = pr.randrange(len(q) // 2); q = q[a:a + max(1, len(q) // 2)]
            else:
                q = seqgen.random_dna(rng, int(rng.integers(0, hi)))
            if pr.random() < 0.1 and len(q) > 2:
                q = q.copy(); q[pr.randrange(len(q))] = ord(pr.choice("N-acgt"))
            qs.append(q.tobytes()); ts.append(t.tobytes())
    if route == "long32": ctx.set_option("long16", 0)
    if route == "generic": ctx.set_option("force_generic", 1)
    try:
        got = ctx.align(qs, ts, typ, m, x, g, True)
    finally:
        ctx.set_option("long16", 1); ctx.set_option("force_generic", 0)
    step = 1 if shape not in ("k1", "k1u") else 37
    for k in range(0, len(qs), step):
        exp = O.align(qs[k], ts[k], typ, m, x, g, True)
        cases += 1
        if got[k] != exp:
            print("MISMATCH", dict(seed=seed, round=rounds, typ=typ, scores=(m, x, g), route=route, shape=shape, opts=opts, k=k,
                                   Q=len(qs[k]), T=len(ts[k]), got=got[k][:2], exp=exp[:2]))
            sys.exit(1)
print(f"fuzz ok: {rounds} rounds, {cases} pairs checked against the oracle, seed {seed}")
