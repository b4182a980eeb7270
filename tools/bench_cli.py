#!/usr/bin/env python3
"""Wall-clock of the b200_mapper executable on a synthetic data set of BASELINE config 4's shape (files on disk in,
PAF on stdout): python tools/bench_cli.py [reads] [gpus]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import seqgen

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(1)
ref = seqgen.random_dna(rng, 4_600_000)
comp = bytes.maketrans(b"ACGT", b"TGCA")
base = []
for i in range(512):
    L = int(np.clip(rng.lognormal(np.log(8000) - 0.125, 0.5), 1000, 40000))
    s0 = int(rng.integers(0, len(ref) - L))
    q = seqgen.mutate(rng, ref[s0:s0 + L], sub=0.024, ins=0.048, dele=0.048).tobytes()
    base.append(q.translate(comp)[::-1] if i % 2 else q)
with tempfile.TemporaryDirectory() as td:
    with open(os.path.join(td, "ref.fa"), "wb") as f:
        f.write(b">ref\n" + ref.tobytes() + b"\n")
    nb = 0
    with open(os.path.join(td, "reads.fq"), "wb") as f:
        for i in range(n_reads):
            sq = base[i % 512]
            nb += len(sq)
            f.write(b"@r%d\n" % i + sq + b"\n+\n" + b"I" * len(sq) + b"\n")
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    for argv in (["-a", "semiGlobal", "-c"], ["-a", "semiGlobal"]):
        t0 = time.perf_counter()
        r = subprocess.run([exe] + argv + ["--gpus", str(gpus), "ref.fa", "reads.fq"], cwd=td, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, B200_TRACE="1"))
        t = time.perf_counter() - t0
        print(json.dumps({"argv": argv, "gpus": gpus, "reads": n_reads, "bases": nb, "rc": r.returncode, "wall_s": t,
                          "reads_per_s": n_reads / t, "paf_lines": r.stdout.count(b"\n"), "paf_bytes": len(r.stdout),
                          "trace": [l for l in r.stderr.decode(errors="replace").splitlines() if "b200_mapper trace" in l]}))
