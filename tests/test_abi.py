"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol that
include/b200map.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from bioinfo1_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200map.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build.lib_path())
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200map.h but not exported"


def test_version_and_count_formula():
    L = capi.lib()
    assert L.b200_version() == 1
    # (w-1) + max(0, n-w+1) + min(w-1, n), n = len-k+1  (team_minimizers.cpp:140-222)
    assert L.b200_minimize_count(4_600_000, 15, 5) == 4_599_990
    assert L.b200_minimize_count(15, 3, 3) == 15
    assert L.b200_minimize_count(3, 3, 3) == 3
    assert L.b200_minimize_count(2, 3, 3) == 0
    assert L.b200_minimize_count(9, 3, 0) == 0
    assert L.b200_minimize_count(9, 3, 1) == 7


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.B200Error) as e:
        capi.Context(0)
    assert e.value.code == capi.E_NOGPU
    with pytest.raises(capi.B200Error) as e:
        capi.align_batch_pointers(0, [b"ACGT"], [b"ACGT"], capi.GLOBAL)
    assert e.value.code == capi.E_NOGPU


def test_unknown_alignment_type_is_rejected_before_any_device_work():
    with pytest.raises(capi.B200Error) as e:
        capi.align_batch_pointers(0, [b"ACGT"], [b"ACGT"], 7)
    assert e.value.code == capi.E_TYPE


def test_product_does_not_reference_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bioinfo1_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) and f != "smoke.py":
                if re.search(r"oracle[/_.]|liboracle|libref", open(os.path.join(dirpath, f), errors="replace").read()):
                    bad.append(f)
    assert not bad, bad


def test_cpp_dropin_library_exports_the_reference_symbols():
    """Itanium-ABI names a caller compiled against the reference headers links to."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", build.lib_path("libteam_b200.so")], capture_output=True,
                         text=True, check=True).stdout
    for sym in ("_ZN4team5AlignEPKcjS1_jNS_13AlignmentTypeEiiiPNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEPj",
                "_ZN4team4KMERC1Eb", "_ZN4team4KMER8MinimizeEPKcjjj", "_ZN4team4KMER23GetMinimizerFrequenciesEv",
                "_ZN4team4KMER19GetUniqueMinimizersEv", "_ZN4team4KMER19SetFrequenciesCountEb",
                "_ZN4team4KMER23MappSeqCharPointerToBitEPKcj", "_ZN4team4KMER17ReverseComplementERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEE"):
        assert sym in out, sym


def test_cli_help_version_and_usage_errors_without_a_device():
    import subprocess
    exe = os.path.join(ROOT, "bioinfo1_b200", "b200_mapper")
    r = subprocess.run([exe, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "toolForGenomeAllignment v3.1.0"
    r = subprocess.run([exe, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage" in r.stdout and "-k KMER" in r.stdout
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Not enough arguments" in r.stderr
    r = subprocess.run([exe, "only_one_file"], capture_output=True, text=True)
    assert r.returncode == 1 and "two input files" in r.stderr.lower()
    r = subprocess.run([exe, "-a", "bogus", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Expected Alignment type" in r.stderr


def test_host_side_2bit_packer_matches_the_device_layout():
    """bioinfo1_b200/csrc/host_pack.hpp (what the pointer-array gather uses) against a byte-at-a-time restatement of
    pack_kernel's layout and flag rules: 200 000 random sequences, with and without foreign bytes (CPU only)."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cpp", "host_pack_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_pack ok" in r.stdout, r.stdout + r.stderr
