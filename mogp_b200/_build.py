"""Build the C-ABI library in-tree: nvcc cross-compiles sm_100a without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("MOGP_B200_LIB") or os.path.join(HERE, "libmogp_b200.so")   # (override: sanitizer builds)
SOURCES = [os.path.join(HERE, "csrc", f) for f in
           ("engine.cu", "chain.inl", "refit.inl", "lbfgsb.hpp", "tile.cuh", "factor.cuh", "score.cuh", "update.cuh", "kmeans.cuh")]
HEADER = os.path.join(os.path.dirname(HERE), "include", "mogp_b200.h")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES + [HEADER])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SOURCES[0]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
