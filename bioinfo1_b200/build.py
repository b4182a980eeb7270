"""In-tree native build (no setuptools, no JIT cache): nvcc for the CUDA library, g++ for the
C++ drop-in wrappers. Outputs land next to this file so they travel with gpurun snapshots.

    libb200map.so   CUDA kernels + the C ABI of include/b200map.h          (csrc/*.cu, one object each)
    libteam_b200.so team::Align / team::KMER drop-in wrappers over the ABI (csrc/team_*.cpp)
    b200_mapper     the <team>_mapper command line (csrc/b200_mapper.cpp)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC,-O2,-Wall,-pthread", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"]
OBJ_DIR = os.path.join(HERE, "build")


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose, log_path=None, append=False):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log_path:
        with open(log_path, "a" if append else "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd))


def lib_path(name="libb200map.so"):
    return os.path.join(HERE, name)


def build_all(verbose=False, force=False):
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cuda_hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".hpp"))]
    hdrs = [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    out = lib_path()
    os.makedirs(OBJ_DIR, exist_ok=True)
    log = os.path.join(HERE, "build_capi.log")
    # every translation unit includes internal.hpp and some kernel headers: a header change rebuilds all of them
    deps = cuda_hdrs + hdrs + [os.path.abspath(__file__)]
    objs = [os.path.join(OBJ_DIR, os.path.basename(u)[:-3] + ".o") for u in units]
    stale = [(u, o) for u, o in zip(units, objs) if force or _newer(o, [u] + deps)]
    if stale:
        if os.path.exists(log):
            os.remove(log)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(stale), os.cpu_count() or 1)) as ex:
            jobs = [ex.submit(_run, [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-c", "-o", o, u], verbose, log, True)
                    for u, o in stale]
            for j in jobs:
                j.result()
    if stale or force or _newer(out, objs):
        _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC,-pthread",
              "-o", out] + objs, verbose)
    cpp_src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.startswith("team_") and f.endswith(".cpp")]
    if cpp_src:
        out2 = lib_path("libteam_b200.so")
        if force or _newer(out2, cpp_src + hdrs + [out]):
            _run(["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-I", INCLUDE, "-o", out2] + cpp_src +
                 ["-L", HERE, "-lb200map", "-Wl,-rpath,$ORIGIN"], verbose)
    cli_src = os.path.join(CSRC, "b200_mapper.cpp")
    if os.path.exists(cli_src):
        exe = os.path.join(HERE, "b200_mapper")
        if force or _newer(exe, [cli_src, out] + hdrs):
            _run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I", INCLUDE, "-o", exe, cli_src, "-L", HERE, "-lb200map", "-lz",
                  "-Wl,-rpath,$ORIGIN"], verbose)
    test_src = os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp")
    if cpp_src and os.path.exists(test_src):
        exe = os.path.join(ROOT, "tests", "cpp", "dropin_test")
        if force or _newer(exe, [test_src, lib_path("libteam_b200.so")] + hdrs):
            _run(["g++", "-O1", "-std=c++17", "-Wall", "-I", INCLUDE, "-o", exe, test_src, "-L", HERE, "-lteam_b200",
                  "-lb200map", "-Wl,-rpath," + HERE], verbose)
    hp_src = os.path.join(ROOT, "tests", "cpp", "host_pack_test.cpp")
    if os.path.exists(hp_src):
        exe = os.path.join(ROOT, "tests", "cpp", "host_pack_test")
        if force or _newer(exe, [hp_src, os.path.join(CSRC, "host_pack.hpp")]):
            _run(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, hp_src], verbose)
    return out


if __name__ == "__main__":
    import sys
    build_all(verbose=True, force="--force" in sys.argv)
