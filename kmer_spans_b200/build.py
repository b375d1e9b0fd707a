"""Builds the sm_100a shared library IN-TREE (kmer_spans_b200/csrc/libkspans_cuda.so).

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libkspans_cuda.so")
SOURCES = ["ks_api.cu"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
    os.path.join("..", "..", "include", "kspans.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("KS_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libkspans_cuda.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
