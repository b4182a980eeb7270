"""GPU parity of the mapper pipeline ("next" rows: index, de-dup, seed lookup, chaining, region, Align)
against the PAF the UNMODIFIED reference mapper printed (tests/golden/mapper, -f 0 => fully
deterministic), and against the CPU mapper oracle on fresh synthetic reads."""
import json
import os
import sys

import numpy as np
import pytest

from cpu_checkers import ROOT, load_oracle
import seqgen

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mapper_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "mapper")
with open(os.path.join(GOLD, "manifest.json")) as f:
    CASES = json.load(f)["cases"]


@pytest.fixture(scope="module")
def ctx():
    from bioinfo1_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _gpu_paf(ctx, ref_path, reads_path, typ, m, x, g, k, w, f, want_cigar):
    from bioinfo1_b200 import capi
    ref_name, ref = mapper_oracle.read_fasta(ref_path)[0]
    reads, is_fastq = mapper_oracle.load_reads(reads_path)
    idx = capi.Index(ctx, ref, k, w, f)
    try:
        res, cigs = idx.map_batch([s for _, s in reads], is_fastq, typ, m, x, g, want_cigar)
    finally:
        idx.close()
    lines = []
    for i, (name, seq) in enumerate(reads):
        r = res[i]
        if not r["mapped"]:
            continue
        d = dict(q_begin=int(r["q_begin"]), q_end=int(r["q_end"]), fwd=bool(r["strand_fwd"]), t_begin=int(r["t_begin"]),
                 t_end=int(r["t_end"]), score=int(r["score"]), cigar=cigs[i] if want_cigar else None)
        lines.append(mapper_oracle.paf_line(name, len(seq), ref_name, len(ref), d, want_cigar))
    return lines


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_mapper_reproduces_reference_paf(ctx, case):
    opts, files = mapper_oracle.parse_mapper_argv(case["argv"])
    got = _gpu_paf(ctx, os.path.join(GOLD, files[0]), os.path.join(GOLD, files[1]), **opts)
    with open(os.path.join(GOLD, case["name"] + ".paf")) as f:
        exp = f.read().splitlines()
    assert got == exp


def test_gpu_mapper_vs_cpu_oracle_fresh_reads(ctx):
    """New synthetic reads (both strands, with repeats in the reference), k=15 w=5, f=0 and default f."""
    from bioinfo1_b200 import capi
    oracle = load_oracle()
    rng = np.random.default_rng(99)
    ref = seqgen.random_dna(rng, 40_000)
    ref[20_000:20_600] = ref[1_000:1_600]
    ref = ref.tobytes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for i in range(30):
        L = int(rng.integers(300, 1800)); s = int(rng.integers(0, len(ref) - L))
        q = seqgen.mutate(rng, np.frombuffer(ref[s:s + L], dtype=np.uint8), sub=0.03, ins=0.03, dele=0.03).tobytes()
        reads.append(q.translate(comp)[::-1] if i % 3 == 0 else q)
    reads += [b"", b"ACGT", b"G" * 300]
    for f in (0.0, 0.001):
        idx = capi.Index(ctx, ref, 15, 5, f)
        oidx = mapper_oracle.Index(oracle, ref, 15, 5, f)
        for fastq in (True, False):
            res, cigs = idx.map_batch(reads, fastq, 2, 1, -1, -1, True)
            for i, rd in enumerate(reads):
                exp = mapper_oracle.map_read(oracle, oidx, rd, 15, 5, 2, 1, -1, -1, True, fasta_path=not fastq)
                if exp is None:
                    assert not res[i]["mapped"], i
                    continue
                got = dict(q_begin=int(res[i]["q_begin"]), q_end=int(res[i]["q_end"]), fwd=bool(res[i]["strand_fwd"]),
                           t_begin=int(res[i]["t_begin"]), t_end=int(res[i]["t_end"]), score=int(res[i]["score"]), cigar=cigs[i])
                assert got == exp, (i, f, fastq)
        idx.close()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cli_stdout_is_byte_identical_to_the_reference_mapper(case):
    """bioinfo1_b200/b200_mapper with the reference's own command lines: stdout must match byte for byte."""
    import subprocess
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    r = subprocess.run([exe] + case["argv"], cwd=GOLD, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    with open(os.path.join(GOLD, case["name"] + ".paf"), "rb") as f:
        assert r.stdout == f.read()


def test_cli_reads_gzip_input(tmp_path):
    """The reference links zlib (through bioparser) and takes .gz files; so does the CLI: same bytes as for the plain files."""
    import gzip
    import shutil
    import subprocess
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    for name in ("ref.fa", "reads.fq"):
        with open(os.path.join(GOLD, name), "rb") as src, gzip.open(tmp_path / (name + ".gz"), "wb") as dst:
            shutil.copyfileobj(src, dst)
    argv = ["-a", "semiGlobal", "-c", "-f", "0", str(tmp_path / "ref.fa.gz"), str(tmp_path / "reads.fq.gz")]
    r = subprocess.run([exe] + argv, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    with open(os.path.join(GOLD, "synth_fq_semi_c.paf"), "rb") as f:
        assert r.stdout == f.read()


def test_cli_statistics_go_to_stderr_and_two_gpus_option_is_accepted():
    import subprocess
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    argv = ["-a", "semiGlobal", "-c", "-f", "0", "-s", "--gpus", "2", "ref.fa", "reads.fq"]
    r = subprocess.run([exe] + argv, cwd=GOLD, capture_output=True, timeout=300)
    assert r.returncode == 0
    with open(os.path.join(GOLD, "synth_fq_semi_c.paf"), "rb") as f:
        assert r.stdout == f.read()          # stdout carries PAF only
    assert b"N50 length" in r.stderr and b"Number of distinct minimizers" in r.stderr


def test_cli_output_does_not_depend_on_the_number_of_devices():
    """--gpus 2 (two devices taking batches of 3 reads from one queue) prints the bytes --gpus 1 prints. Needs two GPUs."""
    import subprocess
    from bioinfo1_b200 import capi
    if capi.lib().b200_device_count() < 2:
        pytest.skip("one device visible")
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    outs = []
    for g in ("1", "2"):
        env = dict(os.environ, B200_MAPPER_BATCH_READS="3", B200_TRACE="1")
        r = subprocess.run([exe, "-a", "semiGlobal", "-c", "-f", "0", "--gpus", g, "ref.fa", "reads.fq"], cwd=GOLD, capture_output=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr.decode(errors="replace")
        outs.append(r.stdout)
        if g == "2":
            assert b"gpu 1 worker 0: batch of" in r.stderr      # the second device really took batches
    assert outs[0] == outs[1]
    with open(os.path.join(GOLD, "synth_fq_semi_c.paf"), "rb") as f:
        assert outs[0] == f.read()


def test_cli_two_workers_over_small_batches_keep_the_output(monkeypatch):
    """Batches of 3 reads taken in turn by two contexts / host threads: same bytes, same (input) order."""
    import subprocess
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    argv = ["-a", "semiGlobal", "-c", "-f", "0", "ref.fa", "reads.fq"]
    with open(os.path.join(GOLD, "synth_fq_semi_c.paf"), "rb") as f:
        exp = f.read()
    for workers in ("2", "1"):
        env = dict(os.environ, B200_MAPPER_BATCH_READS="3", B200_MAPPER_WORKERS=workers, B200_TRACE="1")
        r = subprocess.run([exe] + argv, cwd=GOLD, capture_output=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr.decode(errors="replace")
        assert r.stdout == exp
        if workers == "2":
            assert b"worker 1: batch of" in r.stderr      # the second context really took batches
