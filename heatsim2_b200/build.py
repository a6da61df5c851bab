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
    if not force and not _stale():
        return LIB
    cmd = [NVCC] + FLAGS + list(extra) + sources() + ["-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a not in ("--force", "-v")]
    print(build(force="--force" in sys.argv, verbose=True, extra=extra))
