"""Builds libhvb200.so in-tree with nvcc for sm_100a (explicit commands, no torch involved).

The search is compiled once per dimension (csrc/hvb_dim.cu with -DHVB_DIM=2..6) plus the C ABI (csrc/hvb_api.cu) and the
in-library multi-GPU layer (csrc/hvb_multi.cu); the translation units build in parallel and are linked into one shared
object.  Objects are cached under lib/obj/ and rebuilt when a source or header is newer."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADERS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".hpp"))] + \
          [os.path.join(HERE, "..", "include", "hvb200.h")]
LIBDIR = os.path.join(HERE, "lib")
OUT = os.environ.get("HVB_OUT") or os.path.join(LIBDIR, "libhvb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
EXTRA = os.environ.get("HVB_NVCC_FLAGS", "").split()            # tuning builds: e.g. -DHVB_COOP_MINB=5
DIMS = (2, 3, 4, 5, 6)


def units(tag):
    objdir = os.path.join(LIBDIR, "obj" + tag)
    u = [(os.path.join(CSRC, "hvb_dim.cu"), os.path.join(objdir, "hvb_dim%d.o" % d), ["-DHVB_DIM=%d" % d]) for d in DIMS]
    for name in ("hvb_api", "hvb_multi"):
        src = os.path.join(CSRC, name + ".cu")
        if os.path.exists(src):
            u.append((src, os.path.join(objdir, name + ".o"), []))
    return objdir, u


def build(force=False, verbose=False):
    tag = ("_" + str(abs(hash(" ".join(EXTRA))) % 10 ** 8)) if EXTRA else ""
    objdir, us = units(tag)
    os.makedirs(objdir, exist_ok=True)
    newest_hdr = max(os.path.getmtime(h) for h in HEADERS + [os.path.abspath(__file__)])
    todo = [(s, o, f) for s, o, f in us
            if force or not os.path.exists(o) or os.path.getmtime(o) < max(newest_hdr, os.path.getmtime(s))]

    def compile_one(u):
        s, o, f = u
        cmd = [NVCC] + FLAGS + EXTRA + f + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", o, s]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stderr))
        return r.stderr

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for log in ex.map(compile_one, todo):
                if verbose and log:
                    sys.stderr.write(log)
    objs = [o for _, o, _ in us]
    if todo or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-ldl", "-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
