"""Build libhs2b200.so in-tree with nvcc for sm_100a (no other target).

    python -m heatsim2_b200.build [--force]

The shared library is a plain C-ABI CUDA library (no torch / Python types in
it); the built file is git-ignored but travels to the GPU box with the tree.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhs2b200.so")
OBJ = os.path.join(PKG, "_obj")          # object files (git-ignored, not sent to the GPU box)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
         "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
         "--fmad=true", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG, "..", "include", "hs2_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    """Compile every csrc/*.cu to an object file (one nvcc per file, in parallel: the
    files share no device code) and link them into libhs2b200.so."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    compile_flags = [f for f in FLAGS if f != "-shared"]
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        jobs.append((obj, [NVCC] + compile_flags + list(extra) + ["-c", src, "-o", obj]))
    if verbose:
        for _, cmd in jobs:
            print(" ".join(cmd))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
        list(pool.map(lambda j: subprocess.check_call(j[1]), jobs))
    link = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + [j[0] for j in jobs] + ["-o", LIB]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a not in ("--force", "-v")]
    print(build(force="--force" in sys.argv, verbose=True, extra=extra))
