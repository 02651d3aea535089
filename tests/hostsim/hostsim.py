"""ctypes wrapper of the CPU debugging harness (tests/hostsim/hostsim.cpp).  Test infrastructure only."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostsim.so")
CTR = ("raycasts", "dup_hits", "closed_skips", "cand32", "cand64", "rows", "stages", "seeds", "degenerate", "seed_fail", "dead", "rounds")


def build():
    src = os.path.join(_HERE, "hostsim.cpp")
    core = os.path.join(_HERE, "..", "..", "highvoronoi.jl_b200", "csrc", "hvb_core.cuh")
    host = os.path.join(_HERE, "..", "..", "highvoronoi.jl_b200", "csrc", "hvb_host.hpp")
    geom = os.path.join(_HERE, "..", "..", "highvoronoi.jl_b200", "csrc", "hvb_geometry.cuh")
    hull = os.path.join(_HERE, "..", "..", "highvoronoi.jl_b200", "csrc", "hvb_hull.cuh")
    wrap = os.path.join(_HERE, "..", "..", "highvoronoi.jl_b200", "csrc", "hvb_wrap.cuh")
    nong = os.path.join(_HERE, "..", "..", "highvoronoi.jl_b200", "csrc", "hvb_nongeneral.hpp")
    newest = max(os.path.getmtime(f) for f in (src, core, host, geom, hull, wrap, nong))
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", _SO, src])
    return _SO


def run(xs, base=None, normal=None, ppc=0, probe_scale=0.0, fp32=1, seed_stride=0, trace=False, cells=None):
    """cells: 1-based ids of the cells to explore (Iter / the slab of a rank); None = all"""
    L = ctypes.CDLL(build())
    L.hostsim_set_trace(1 if trace else 0)
    L.hostsim_set_active.restype = None
    L.hostsim_set_active.argtypes = [ctypes.c_int64, ctypes.c_void_p]
    if cells is None:
        L.hostsim_set_active(0, None)
    else:
        act = np.zeros(len(xs), dtype=np.uint8)
        act[np.asarray(cells, dtype=np.int64) - 1] = 1
        L.hostsim_set_active(len(act), act.ctypes.data_as(ctypes.c_void_p))
    L.hostsim_run.restype = ctypes.c_void_p
    L.hostsim_run.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    for f in ("hostsim_counts", "hostsim_fetch", "hostsim_free"):
        getattr(L, f).restype = None
    L.hostsim_counts.argtypes = [ctypes.c_void_p] * 4
    L.hostsim_fetch.argtypes = [ctypes.c_void_p] * 4
    L.hostsim_free.argtypes = [ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if base is None:
        base = np.zeros((0, d)); normal = np.zeros((0, d))
    base = np.ascontiguousarray(base, dtype=np.float64); normal = np.ascontiguousarray(normal, dtype=np.float64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    h = L.hostsim_run(d, n, P(xs), base.shape[0], P(base), P(normal), ppc, probe_scale, fp32, seed_stride)
    nv, nr = ctypes.c_int64(), ctypes.c_int64()
    ctr = np.zeros(12, dtype=np.int64)
    L.hostsim_counts(h, ctypes.byref(nv), ctypes.byref(nr), P(ctr))
    sig = np.empty((nv.value, d + 1), dtype=np.int64); r = np.empty((nv.value, d)); re = np.empty((nr.value, d), dtype=np.int64)
    L.hostsim_fetch(h, P(sig), P(r), P(re))
    out = dict(sig=sig, r=r, ray_edge=re, stats=dict(zip(CTR, ctr.tolist())))
    if trace:
        L.hostsim_trace_size.restype = ctypes.c_int64
        L.hostsim_trace_size.argtypes = [ctypes.c_void_p]
        L.hostsim_trace_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        m = L.hostsim_trace_size(h)
        tr = np.zeros(m, dtype=np.int32)
        if m:
            L.hostsim_trace_fetch(h, P(tr))
        out["trace"] = tr.reshape(-1, 2)
    L.hostsim_free(h)
    L.hostsim_set_trace(0)
    L.hostsim_set_active(0, None)
    return out


def volumes(xs, sig, base=None, normal=None):
    """cell volumes from vertex rows with the product's own formula (vertex_flag_sum, hvb_geometry.cuh) on the host"""
    L = ctypes.CDLL(build())
    L.hostsim_volumes.restype = None
    L.hostsim_volumes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if base is None:
        base = np.zeros((0, d)); normal = np.zeros((0, d))
    base = np.ascontiguousarray(base, dtype=np.float64); normal = np.ascontiguousarray(normal, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.int64)
    vol = np.zeros(n)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.hostsim_volumes(d, n, P(xs), base.shape[0], P(base), P(normal), sig.shape[0], P(sig), P(vol))
    return vol


def areas(xs, sig, off, ids, base=None, normal=None):
    """interface areas aligned with the CSR neighbour lists, with the product's own formula on the host"""
    L = ctypes.CDLL(build())
    L.hostsim_areas.restype = None
    L.hostsim_areas.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if base is None:
        base = np.zeros((0, d)); normal = np.zeros((0, d))
    base = np.ascontiguousarray(base, dtype=np.float64); normal = np.ascontiguousarray(normal, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.int64); off = np.ascontiguousarray(off, dtype=np.int64); ids = np.ascontiguousarray(ids, dtype=np.int64)
    area = np.zeros(int(off[n]))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.hostsim_areas(d, n, P(xs), base.shape[0], P(base), P(normal), sig.shape[0], P(sig), P(off), P(ids), P(area))
    return area


def hull(xs, ppc=0):
    """convex hull by the facet walk of hvb_hull.cuh on the host: (facets [F, d] sorted 1-based ids, normals [F, d], stats)"""
    L = ctypes.CDLL(build())
    L.hostsim_hull.restype = ctypes.c_void_p
    L.hostsim_hull.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
    for f in ("hostsim_hull_counts", "hostsim_hull_fetch", "hostsim_hull_free"):
        getattr(L, f).restype = None
    L.hostsim_hull_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.hostsim_hull_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.hostsim_hull_free.argtypes = [ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    h = L.hostsim_hull(d, n, P(xs), ppc)
    c = np.zeros(5, dtype=np.int64)
    L.hostsim_hull_counts(h, P(c))
    facets = np.empty((c[0], d), dtype=np.int64); normals = np.empty((c[0], d))
    L.hostsim_hull_fetch(h, P(facets), P(normals))
    L.hostsim_hull_free(h)
    return facets, normals, dict(facets=int(c[0]), raycasts=int(c[1]), records=int(c[2]), rounds=int(c[3]), degenerate=int(c[4]))


def wrap(xs, slots=1, fp32=1):
    """convex hull by gift wrapping (hvb_wrap.cuh) on the host: (facets [F, d] sorted 1-based ids, normals [F, d], centres [F, d], stats);
    `records` counts the FP64 evaluations"""
    L = ctypes.CDLL(build())
    L.hostsim_wrap.restype = ctypes.c_void_p
    L.hostsim_wrap.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    for f in ("hostsim_hull_counts", "hostsim_hull_fetch", "hostsim_hull_free", "hostsim_wrap_centres"):
        getattr(L, f).restype = None
    L.hostsim_hull_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.hostsim_hull_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.hostsim_wrap_centres.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.hostsim_hull_free.argtypes = [ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    h = L.hostsim_wrap(d, n, P(xs), slots, fp32)
    c = np.zeros(5, dtype=np.int64)
    L.hostsim_hull_counts(h, P(c))
    facets = np.empty((c[0], d), dtype=np.int64); normals = np.empty((c[0], d)); centres = np.empty((c[0], d))
    L.hostsim_hull_fetch(h, P(facets), P(normals))
    L.hostsim_wrap_centres(h, P(centres))
    L.hostsim_hull_free(h)
    return facets, normals, centres, dict(facets=int(c[0]), raycasts=int(c[1]), fp64=int(c[2]), rounds=int(c[3]), degenerate=int(c[4]))


def moments(xs, sig, base=None, normal=None):
    """integrals of 1, x_a and x_a x_b (a <= b, row by row) over every cell from vertex rows, with the product's own formula
    (vertex_flag_moments, hvb_geometry.cuh) on the host: [n, 1 + d + d (d + 1) / 2]"""
    L = ctypes.CDLL(build())
    L.hostsim_moments.restype = None
    L.hostsim_moments.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if base is None:
        base = np.zeros((0, d)); normal = np.zeros((0, d))
    base = np.ascontiguousarray(base, dtype=np.float64); normal = np.ascontiguousarray(normal, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.int64)
    out = np.zeros((n, 1 + d + d * (d + 1) // 2))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.hostsim_moments(d, n, P(xs), base.shape[0], P(base), P(normal), sig.shape[0], P(sig), P(out))
    return out


def resolve(xs, base=None, normal=None):
    """a cloud in non-general position resolved as the library does it (perturbation + merge, csrc/hvb_ctx.cuh resolve_degenerate):
    the search on the host build, the merge and the neighbour lists with the library's own host functions (hvb_nongeneral.hpp).
    -> dict(off, ids, r, nb_off, nb_ids, max_siglen, simplicial)"""
    L = ctypes.CDLL(build())
    L.hostsim_resolve.restype = ctypes.c_void_p
    L.hostsim_resolve.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    for f in ("hostsim_resolve_counts", "hostsim_resolve_fetch", "hostsim_resolve_free"):
        getattr(L, f).restype = None
    L.hostsim_resolve_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.hostsim_resolve_fetch.argtypes = [ctypes.c_void_p] * 6
    L.hostsim_resolve_free.argtypes = [ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if base is None:
        base = np.zeros((0, d)); normal = np.zeros((0, d))
    base = np.ascontiguousarray(base, dtype=np.float64); normal = np.ascontiguousarray(normal, dtype=np.float64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    h = L.hostsim_resolve(d, n, P(xs), base.shape[0], P(base), P(normal))
    c = np.zeros(5, dtype=np.int64)
    L.hostsim_resolve_counts(h, P(c))
    off = np.empty(c[0] + 1, dtype=np.int64); ids = np.empty(c[1], dtype=np.int64); r = np.empty((c[0], d))
    nb_off = np.empty(n + 1, dtype=np.int64); nb_ids = np.empty(c[2], dtype=np.int64)
    L.hostsim_resolve_fetch(h, P(off), P(ids), P(r), P(nb_off), P(nb_ids))
    L.hostsim_resolve_free(h)
    return dict(off=off, ids=ids, r=r, nb_off=nb_off, nb_ids=nb_ids, max_siglen=int(c[3]), simplicial=int(c[4]))


def area_moments(xs, sig, off, ids, base=None, normal=None):
    """area and first moment (global coordinates) of every interface aligned with the CSR neighbour lists, with the product's
    own formula on the host: [off[n], 1 + d]"""
    L = ctypes.CDLL(build())
    L.hostsim_area_moments.restype = None
    L.hostsim_area_moments.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n, d = xs.shape
    if base is None:
        base = np.zeros((0, d)); normal = np.zeros((0, d))
    base = np.ascontiguousarray(base, dtype=np.float64); normal = np.ascontiguousarray(normal, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.int64); off = np.ascontiguousarray(off, dtype=np.int64); ids = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.zeros((int(off[n]), 1 + d))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.hostsim_area_moments(d, n, P(xs), base.shape[0], P(base), P(normal), sig.shape[0], P(sig), P(off), P(ids), P(out))
    return out
