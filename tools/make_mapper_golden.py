#!/usr/bin/env python3
"""Generate tests/golden/mapper/: synthetic inputs + the PAF the UNMODIFIED reference mapper prints
for them (oracle/_ref/ref_mapper = /root/reference/team_mapper.cpp built against oracle/bioparser_shim
by oracle/Makefile). Run in the authoring container only:  make -C oracle && python tools/make_mapper_golden.py
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import seqgen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "mapper")
REF_MAPPER = os.path.join(ROOT, "oracle", "_ref", "ref_mapper")
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def write_fasta(path, recs):
    with open(path, "w") as f:
        for name, seq in recs:
            f.write(f">{name}\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70].decode() + "\n")


def write_fastq(path, recs):
    with open(path, "w") as f:
        for name, seq in recs:
            f.write(f"@{name}\n{seq.decode()}\n+\n{'I' * len(seq)}\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    ref = seqgen.random_dna(rng, 60_000)
    # a repeated segment so that some minimizers have several reference positions
    ref[30_000:30_800] = ref[5_000:5_800]
    reads = []
    for i in range(48):
        L = int(rng.integers(500, 2600))
        s = int(rng.integers(0, len(ref) - L))
        frag = ref[s:s + L]
        q = seqgen.mutate(rng, frag, sub=0.02, ins=0.04, dele=0.04)
        b = q.tobytes()
        if i % 2:
            b = b.translate(COMP)[::-1]
        reads.append((f"read{i}", b))
    reads.append(("tiny", b"ACGTACGTAC"))                      # shorter than k: unmapped
    reads.append(("junk", seqgen.random_dna(rng, 900).tobytes()))  # unrelated: usually unmapped
    write_fasta(os.path.join(OUT, "ref.fa"), [("synthref", ref.tobytes())])
    write_fasta(os.path.join(OUT, "reads.fa"), reads)
    write_fastq(os.path.join(OUT, "reads.fq"), reads)
    # the reference's own toy files, reproduced verbatim as test inputs (data, not code)
    toy = {"toy_ref.fa": ">ref\nACGTACGAC\n",
           "toy_seq.fa": ">seq1\nGTACGT\n>seq2\nTACGATG\n>seq3\nACGTAC\n>seq4\nATTACAC\n>seq5\nTCGTAAGA\n>seq6\nTTACAC\n",
           "toy_reference84.fa": ">ref\nAATCGTGACGTACATGGACAGCTTACGGTACATGGAGGCGTACATGGACAAGCTTGACGTACATGGACATTTGGCGTACATGGA\n",
           "toy_doc.fa": ">seq1\nTGACGTACATGGACA\n>seq2\nCGTACATGGA\n"}
    for fn, txt in toy.items():
        with open(os.path.join(OUT, fn), "w") as f:
            f.write(txt)

    cases = [
        ("synth_fq_semi_c", ["-a", "semiGlobal", "-c", "-f", "0", "ref.fa", "reads.fq"]),
        ("synth_fa_semi_c", ["-a", "semiGlobal", "-c", "-f", "0", "ref.fa", "reads.fa"]),
        ("synth_fq_global", ["-a", "global", "-f", "0", "ref.fa", "reads.fq"]),
        ("synth_fq_local_c_k12w4", ["-a", "local", "-c", "-f", "0", "-k", "12", "-w", "4", "ref.fa", "reads.fq"]),
        ("synth_fq_semi_scores", ["-a", "semiGlobal", "-c", "-f", "0", "-m", "2", "-n", "-3", "-g", "-2", "ref.fa", "reads.fq"]),
        ("toy_semi_k3w2", ["-a", "semiGlobal", "-k", "3", "-w", "2", "-c", "toy_ref.fa", "toy_seq.fa"]),
        ("toy_global_k3w2", ["-a", "global", "-k", "3", "-w", "2", "-c", "toy_ref.fa", "toy_seq.fa"]),
        ("toy_local_k3w2", ["-a", "local", "-k", "3", "-w", "2", "-c", "toy_ref.fa", "toy_seq.fa"]),
        ("toy_deck_positive_gap", ["-a", "local", "-m", "2", "-n", "-1", "-g", "2", "-k", "3", "-w", "2", "-c", "toy_ref.fa", "toy_seq.fa"]),
        ("toy_doc_k5w3", ["-a", "semiGlobal", "-k", "5", "-w", "3", "-c", "toy_reference84.fa", "toy_doc.fa"]),
        ("toy_default_kw_prints_nothing", ["toy_ref.fa", "toy_seq.fa"]),
    ]
    manifest = []
    for name, argv in cases:
        r = subprocess.run([REF_MAPPER] + argv, cwd=OUT, capture_output=True, text=True, check=True)
        with open(os.path.join(OUT, name + ".paf"), "w") as f:
            f.write(r.stdout)
        manifest.append({"name": name, "argv": argv, "lines": r.stdout.count("\n")})
        print(name, r.stdout.count("\n"), "lines")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump({"generator": "tools/make_mapper_golden.py",
                   "source": "unmodified /root/reference/team_mapper.cpp via oracle/_ref/ref_mapper", "cases": manifest}, f, indent=1)


if __name__ == "__main__":
    main()
