#!/usr/bin/env python3
"""Wall-clock of the b200_mapper executable on BASELINE config 4's data set (tests/synth.c: 4.6 Mbp reference, distinct
ONT-like reads; files in, PAF on stdout): python tools/bench_cli.py [reads] [gpus]"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ref = synth.dna(1, 4_600_000)
buf, off = synth.ont_reads(2, ref, n=n_reads)
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
with tempfile.TemporaryDirectory(dir=base) as td:
    with open(os.path.join(td, "ref.fa"), "wb") as f:
        f.write(b">ref\n" + ref[:4_600_000].tobytes() + b"\n")
    nb = int(off[-1])
    with open(os.path.join(td, "reads.fq"), "wb") as f:
        for i in range(n_reads):
            sq = buf[int(off[i]):int(off[i + 1])].tobytes()
            f.write(b"@r%d\n" % i + sq + b"\n+\n" + b"I" * len(sq) + b"\n")
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    for argv in (["-a", "semiGlobal", "-c"], ["-a", "semiGlobal"]):
        for rep in range(2):   # the second run has the files in the page cache and the driver warm
            # stdout goes to a file, as a user's `> out.paf` would (a Python pipe reader drains 300 MB far slower than the mapper writes)
            with open(os.path.join(td, "out.paf"), "wb") as outf:
                t0 = time.perf_counter()
                r = subprocess.run([exe] + argv + ["--gpus", str(gpus), "ref.fa", "reads.fq"], cwd=td, stdout=outf, stderr=subprocess.PIPE, env=dict(os.environ, B200_TRACE="1"))
                t = time.perf_counter() - t0
        paf = open(os.path.join(td, "out.paf"), "rb").read()
        print(json.dumps({"argv": argv, "gpus": gpus, "reads": n_reads, "bases": nb, "rc": r.returncode, "wall_s": t,
                          "reads_per_s": n_reads / t, "paf_lines": paf.count(b"\n"), "paf_bytes": len(paf),
                          "paf_md5": hashlib.md5(paf).hexdigest(),
                          "trace": [l for l in r.stderr.decode(errors="replace").splitlines() if "b200_mapper trace" in l][-6:]}), flush=True)
