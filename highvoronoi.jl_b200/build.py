"""Builds libhvb200.so in-tree with nvcc for sm_100a (explicit command, no torch involved)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "hvb_api.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("hvb_api.cu", "hvb_core.cuh", "hvb_kernels.cuh", "hvb_coop.cuh", "hvb_geometry.cuh", "hvb_host.hpp")] + \
       [os.path.join(HERE, "..", "include", "hvb200.h")]
OUT = os.path.join(HERE, "lib", "libhvb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def build(force=False, verbose=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
