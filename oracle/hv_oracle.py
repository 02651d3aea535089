"""ctypes front-end of the CPU restatement (oracle/hv_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhvoracle.so")
_lib = None

STAT_NAMES = ("raycasts", "nn_calls", "inrange_calls", "points_visited", "descents",
              "corrections", "degenerate", "duplicates", "rejected", "search_us", "build_us")


def build(force=False):
    """Compile the restatement with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "hv_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        L.hvo_run.restype = ctypes.c_void_p
        L.hvo_run.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64]
        L.hvo_set_method.restype = ctypes.c_int
        L.hvo_set_method.argtypes = [ctypes.c_int]
        L.hvo_error.restype = ctypes.c_char_p
        L.hvo_error.argtypes = [ctypes.c_void_p]
        for name in ("hvo_counts", "hvo_fetch_vertices", "hvo_fetch_rays", "hvo_fetch_neighbors", "hvo_stats", "hvo_free"):
            getattr(L, name).restype = None
        L.hvo_counts.argtypes = [ctypes.c_void_p] * 4
        L.hvo_fetch_vertices.argtypes = [ctypes.c_void_p] * 3
        L.hvo_fetch_rays.argtypes = [ctypes.c_void_p] * 5
        L.hvo_fetch_neighbors.argtypes = [ctypes.c_void_p] * 3
        L.hvo_stats.argtypes = [ctypes.c_void_p] * 2
        L.hvo_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


METHODS = {"RCNonGeneral": 0, "RCNonGeneralHP": 0, "RCStandard": 0, "RCOriginal": 1, "RCCombined": 2, "RCNonGeneralFast": 3}


def run(xs, plane_base=None, plane_normal=None, nthreads=1, seed=0, method=0):
    """voronoi(xs; searcher=Raycast(xs; domain, options=RaycastParameter(method=...))) restated on the CPU.

    method: 0 / "RCNonGeneral" (the default, raycast.jl:794-970), 1 / "RCOriginal" (:972-1012), 2 / "RCCombined" (:504-528),
    3 / "RCNonGeneralFast" (:542-631) -- the four methods of test/rcmethods.jl.

    Returns dict(sig[V,d+1] int64 1-based sorted rows in lexicographic order, r[V,d], ray_edge, ray_base,
    ray_dir, ray_node, nb_off[n+1], nb_ids, stats)."""
    L = _load()
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if plane_base is None:
        plane_base = np.zeros((0, d))
        plane_normal = np.zeros((0, d))
    pb = np.ascontiguousarray(plane_base, dtype=np.float64).reshape(-1, d)
    pn = np.ascontiguousarray(plane_normal, dtype=np.float64).reshape(-1, d)
    if L.hvo_set_method(int(METHODS.get(method, method))) != 0:
        raise ValueError("the restatement has methods 0 (RCNonGeneral), 1 (RCOriginal), 2 (RCCombined) and 3 (RCNonGeneralFast)")
    try:
        h = L.hvo_run(d, n, _p(xs), pb.shape[0], _p(pb), _p(pn), int(nthreads), int(seed))
    finally:
        L.hvo_set_method(0)
    try:
        err = L.hvo_error(h)
        if err:
            raise RuntimeError(err.decode())
        nv, nr, nn = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        L.hvo_counts(h, ctypes.byref(nv), ctypes.byref(nr), ctypes.byref(nn))
        sig = np.empty((nv.value, d + 1), dtype=np.int64)
        r = np.empty((nv.value, d), dtype=np.float64)
        L.hvo_fetch_vertices(h, _p(sig), _p(r))
        re = np.empty((nr.value, d), dtype=np.int64)
        rb = np.empty((nr.value, d))
        rd = np.empty((nr.value, d))
        rn = np.empty((nr.value,), dtype=np.int64)
        L.hvo_fetch_rays(h, _p(re), _p(rb), _p(rd), _p(rn))
        off = np.empty((n + 1,), dtype=np.int64)
        ids = np.empty((nn.value,), dtype=np.int64)
        L.hvo_fetch_neighbors(h, _p(off), _p(ids))
        st = np.zeros(len(STAT_NAMES), dtype=np.int64)
        L.hvo_stats(h, _p(st))
    finally:
        L.hvo_free(h)
    return dict(sig=sig, r=r, ray_edge=re, ray_base=rb, ray_dir=rd, ray_node=rn, nb_off=off, nb_ids=ids,
                stats=dict(zip(STAT_NAMES, st.tolist())))
