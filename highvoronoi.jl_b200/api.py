"""Host-side mirror of the reference interface for the raycast vertex-search path.

Julia is not available in the build environment, so this Python layer plays the role of the Julia shim
(julia/HighVoronoiB200.jl): same names, argument meaning and error behaviour as the reference for this path --
VoronoiNodes (voronoinodes.jl:14), Boundary / cuboid (boundary.jl:22-29, 510-534), RaycastParameter
(raycast-types.jl:312-324), Raycast (raycast.jl:26), voronoi (sysvoronoi.jl:21), VoronoiGeometry (geometry.jl:139),
VoronoiData (voronoidata.jl:621).  All computation happens in libhvb200.so on the GPU.
"""
import ctypes

import numpy as np

from . import _abi
from ._abi import HVBError  # noqa: F401


def VoronoiNodes(x):
    """VoronoiNodes(x::Matrix) (voronoinodes.jl:14-20): columns are points.  Accepts (d, N) like the reference;
    an (N, d) array is accepted when it cannot be confused (N > 6 >= d)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("VoronoiNodes expects a matrix")
    if x.shape[0] <= 6 < x.shape[1]:
        x = x.T
    return np.ascontiguousarray(x)


class Boundary:
    """Boundary(planes...) (boundary.jl:78-85): convex domain {y : normal_p . (y - base_p) <= 0}."""

    def __init__(self, base=None, normal=None, periodic=()):
        self.base = np.zeros((0, 0)) if base is None else np.ascontiguousarray(base, dtype=np.float64)
        self.normal = np.zeros((0, 0)) if normal is None else np.ascontiguousarray(normal, dtype=np.float64)
        self.periodic = tuple(periodic)

    def __len__(self):
        return self.base.shape[0]

    def plane_bc(self):
        """Plane.BC per plane (boundary.jl:15-29): 1-based partner index for periodic planes, 0 otherwise.  `periodic`
        lists 1-based axes of a cuboid: plane 2i-1 and plane 2i are partners (boundary.jl:517-519)."""
        bc = np.zeros(len(self), dtype=np.int32)
        for i in self.periodic:
            bc[2 * i - 2] = 2 * i
            bc[2 * i - 1] = 2 * i - 1
        return bc


def cuboid(dim, dimensions=None, periodic=None, neumann=(), offset=None):
    """cuboid(dim; dimensions, periodic, neumann, offset) (boundary.jl:510-534).  Plane 2i-1 is the upper face of
    axis i, plane 2i the lower one.  NOTE the reference's default is periodic=1:dim; the search itself treats every
    plane as a mirror (geometry.jl:156), periodisation is host orchestration outside this path."""
    dimensions = np.ones(dim) if dimensions is None else np.asarray(dimensions, dtype=np.float64)
    offset = np.zeros(dim) if offset is None else np.asarray(offset, dtype=np.float64)
    periodic = tuple(range(1, dim + 1)) if periodic is None else tuple(periodic)
    base = np.zeros((2 * dim, dim))
    normal = np.zeros((2 * dim, dim))
    for i in range(dim):
        base[2 * i] = offset
        base[2 * i, i] += dimensions[i]
        normal[2 * i, i] = 1.0
        base[2 * i + 1] = offset
        normal[2 * i + 1, i] = -1.0
    return Boundary(base, normal, periodic)


# method / threading singletons (raycast-types.jl:244-284, HighVoronoi.jl:60-81)
RCStandard = RCNonGeneral = RCNonGeneralHP = 0
RCOriginal = 1
RCCombined = 2
RCNonGeneralFast = 3
RCOriginalSafety, RCNonGeneralSkip, RCOriginalHP, RCNonGeneralCutoff = 4, 5, 6, 7


class SingleThread:
    pass


class B200Thread:
    """The new `threading` singleton the Julia shim adds.

    B200Thread(ngpus=N[, devices=[...]]): ONE process drives N GPUs (hvb_create_multi) -- the analogue of the reference's
    MultiThread(N, 1) (sysvoronoi.jl:50-82).  B200Thread(device, rank, world): this process is slab `rank` of `world`
    processes, one per GPU (torchrun / MPI style); see hvb200.multigpu for the communicator."""

    def __init__(self, device=0, rank=0, world=1, ngpus=1, devices=None):
        self.device, self.rank, self.world = device, rank, world
        self.ngpus, self.devices = int(ngpus), (None if devices is None else [int(v) for v in devices])


def RaycastParameter(variance_tol=1e-15, break_tol=1e-5, b_nodes_tol=1e-7, plane_tolerance=1e-12, ray_tol=1e-12,
                     method=RCStandard, threading=None, **backend):
    """RaycastParameter{Float64}(; ...) (raycast-types.jl:312-324) plus backend knobs (fp32_filter, on_degenerate,
    points_per_cell, seed_stride, sort_output, vertex_capacity, probe_scale)."""
    p = _abi.hvb_params()
    _abi.lib().hvb_default_params(ctypes.byref(p))
    p.variance_tol, p.break_tol, p.b_nodes_tol, p.plane_tolerance, p.ray_tol = variance_tol, break_tol, b_nodes_tol, plane_tolerance, ray_tol
    p.method = int(method)
    if threading is not None and isinstance(threading, B200Thread):
        p.device, p.rank, p.world = threading.device, threading.rank, threading.world
        if threading.ngpus > 1 or threading.devices is not None:
            p._multi = (threading.ngpus if threading.devices is None else len(threading.devices), threading.devices)
    for k, v in backend.items():
        if not hasattr(p, k):
            raise TypeError("unknown search setting %r" % k)
        setattr(p, k, v)
    return p


class Raycast:
    """Raycast(xs; domain=Boundary(), options=RaycastParameter()) (raycast.jl:26): owns the device context
    (generators + spatial index)."""

    def __init__(self, xs, domain=None, options=None, periodic=False):
        """periodic=True: the periodic planes of `domain` are honoured by the backend (halo generators + certificate,
        hvb_create_periodic) -- what VoronoiGeometry does for such a domain.  The default treats every plane as a
        mirror, like the reference's own voronoi() call (geometry.jl:156)."""
        self.xs = VoronoiNodes(xs) if not (isinstance(xs, np.ndarray) and xs.flags.c_contiguous and xs.dtype == np.float64 and xs.ndim == 2 and xs.shape[0] > xs.shape[1]) else xs
        self.domain = domain if domain is not None else Boundary()
        self.parameters = options if options is not None else RaycastParameter()
        n, d = self.xs.shape
        self.n, self.dim = n, d
        L = _abi.lib()
        self._ctx = ctypes.c_void_p()
        P = len(self.domain)
        base = self.domain.base.ctypes.data_as(ctypes.c_void_p) if P else None
        normal = self.domain.normal.ctypes.data_as(ctypes.c_void_p) if P else None
        self.periodic = bool(periodic) and len(self.domain.periodic) > 0
        self.multi = getattr(self.parameters, "_multi", None)
        if self.multi is not None:
            ngpus, devices = self.multi
            self._bc = self.domain.plane_bc() if self.periodic else None
            dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
            rc = L.hvb_create_multi(ctypes.byref(self._ctx), d, n, self.xs.ctypes.data_as(ctypes.c_void_p), P, base, normal,
                                    self._bc.ctypes.data_as(ctypes.c_void_p) if self.periodic else None,
                                    ctypes.byref(self.parameters), ngpus, None if dev is None else dev.ctypes.data_as(ctypes.c_void_p))
        elif self.periodic:
            self._bc = self.domain.plane_bc()
            rc = L.hvb_create_periodic(ctypes.byref(self._ctx), d, n, self.xs.ctypes.data_as(ctypes.c_void_p), P, base, normal,
                                       self._bc.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.parameters))
        else:
            rc = L.hvb_create(ctypes.byref(self._ctx), d, n, self.xs.ctypes.data_as(ctypes.c_void_p), P, base, normal,
                              ctypes.byref(self.parameters))
        if rc != _abi.HVB_OK:
            self._ctx = None
            _abi.check(rc, None)

    def halo(self):
        """(origin, mult, xs, margin) of a periodic searcher: halo generator n+1+i copies caller generator origin[i]
        shifted by sum_k mult[i,k] periods of periodic pair k (references / reference_shifts, domain.jl:338-390)."""
        L, ctx = _abi.lib(), self._ctx
        nh, npairs, margin = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_double()
        _abi.check(L.hvb_halo_count(ctx, ctypes.byref(nh), ctypes.byref(npairs), ctypes.byref(margin)), ctx)
        origin = np.empty((nh.value,), dtype=np.int64)
        mult = np.empty((nh.value, max(npairs.value, 0)), dtype=np.int32)
        xs = np.empty((nh.value, self.dim), dtype=np.float64)
        if nh.value:
            P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            _abi.check(L.hvb_fetch_halo(ctx, P(origin), P(mult), P(xs)), ctx)
        return origin, mult, xs, margin.value

    def set_points(self, xs):
        """re-target this searcher to another generator set (same dimension and domain), re-using the device context"""
        xs = VoronoiNodes(xs) if not (isinstance(xs, np.ndarray) and xs.flags.c_contiguous and xs.dtype == np.float64 and xs.ndim == 2 and xs.shape[0] > xs.shape[1]) else xs
        if xs.shape[1] != self.dim:
            raise ValueError("dimension mismatch")
        _abi.check(_abi.lib().hvb_set_points(self._ctx, xs.shape[0], xs.ctypes.data_as(ctypes.c_void_p)), self._ctx)
        self.xs, self.n = xs, xs.shape[0]

    def close(self):
        if getattr(self, "_ctx", None):
            _abi.lib().hvb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def owned(self):
        """bool mask over the caller's cells: True where this searcher's slab owns the cell (all True unless world > 1)"""
        m = np.empty((self.n,), dtype=np.uint8)
        _abi.check(_abi.lib().hvb_fetch_owned(self._ctx, m.ctypes.data_as(ctypes.c_void_p)), self._ctx)
        return m.astype(bool)

    def stats(self):
        s = _abi.hvb_stats_t()
        _abi.check(_abi.lib().hvb_stats(self._ctx, ctypes.byref(s)), self._ctx)
        return s.as_dict()


class _Owned(np.ndarray):
    """ndarray over page-locked memory of a device context; keeps the owning searcher alive"""
    _owner = None


def _wrap(ptr, shape, ctype, owner):
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=np.dtype(ctype))
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,)).reshape(shape).view(_Owned)
    arr._owner = owner
    return arr


class VoronoiMesh:
    """Result of voronoi(): the vertex database in the reference's external numbering (1-based ids, plane p = n+p).

    With copy=False (default) `sig`, `r` and the neighbour arrays are zero-copy views of the context's page-locked
    staging memory: they stay valid until the next search / set_points on the same searcher.  copy=True returns
    private arrays (hvb_fetch_*)."""

    def __init__(self, searcher, copy=False):
        self.searcher = searcher
        copy = bool(copy) or getattr(searcher, "multi", None) is not None     # shards of several GPUs have no common staging buffer
        self.copy = copy
        self.n, self.dim = searcher.n, searcher.dim
        L, ctx = _abi.lib(), searcher._ctx
        nv, nr, ml = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _abi.check(L.hvb_counts(ctx, ctypes.byref(nv), ctypes.byref(nr), ctypes.byref(ml)), ctx)
        d = self.dim
        self.max_siglen = ml.value
        self.sig_off = self.sig_ids = None
        if ml.value > d + 1:
            # non-general position resolved by the backend (on_degenerate = 2): vertices with more than d + 1 generators, the
            # reference's variable-length sig vectors (raycast.jl:926-949).  `sig` does not exist; use sigs() / sig_off, sig_ids
            P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            self.copy = copy = True
            self.sig_off = np.empty((nv.value + 1,), dtype=np.int64)
            _abi.check(L.hvb_fetch_vertices_var(ctx, P(self.sig_off), None, None), ctx)
            self.sig_ids = np.empty((int(self.sig_off[-1]),), dtype=np.int64)
            self.r = np.empty((nv.value, d), dtype=np.float64)
            _abi.check(L.hvb_fetch_vertices_var(ctx, P(self.sig_off), P(self.sig_ids), P(self.r)), ctx)
            self.sig = None
        elif copy:
            self.sig = np.empty((nv.value, d + 1), dtype=np.int64)
            self.r = np.empty((nv.value, d), dtype=np.float64)
            _abi.check(L.hvb_fetch_vertices(ctx, self.sig.ctypes.data_as(ctypes.c_void_p), self.r.ctypes.data_as(ctypes.c_void_p)), ctx)
        else:
            ps, pr, cnt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
            self.wire32 = bool(searcher.parameters.wire32)
            if self.wire32:                 # compact wire format: int32 ids (hvb_view_vertices32)
                _abi.check(L.hvb_view_vertices32(ctx, ctypes.byref(ps), ctypes.byref(pr), ctypes.byref(cnt)), ctx)
                self.sig = _wrap(ps, (cnt.value, d + 1), ctypes.c_int32, searcher)
            else:
                _abi.check(L.hvb_view_vertices(ctx, ctypes.byref(ps), ctypes.byref(pr), ctypes.byref(cnt)), ctx)
                self.sig = _wrap(ps, (cnt.value, d + 1), ctypes.c_int64, searcher)
            self.r = _wrap(pr, (cnt.value, d), ctypes.c_double, searcher)
        self.ray_edge = np.empty((nr.value, d), dtype=np.int64)
        self.ray_base = np.empty((nr.value, d))
        self.ray_dir = np.empty((nr.value, d))
        self.ray_node = np.empty((nr.value,), dtype=np.int64)
        if nr.value:
            P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            _abi.check(L.hvb_fetch_rays(ctx, P(self.ray_edge), P(self.ray_base), P(self.ray_dir), P(self.ray_node)), ctx)
        self._nb = None
        self._cell_index = None
        self.n_halo = 0
        if getattr(searcher, "periodic", False):
            # extended numbering: caller generators 1..n, halo n+1..n+n_halo, planes behind
            self.halo_origin, self.halo_mult, self.halo_xs, self.margin = searcher.halo()
            self.n_halo = self.halo_origin.shape[0]
            self.n_user = self.n
            self.n = self.n + self.n_halo
            self.canonical = np.empty((nv.value,), dtype=np.uint8)
            if nv.value:
                _abi.check(L.hvb_fetch_vertex_flags(ctx, self.canonical.ctypes.data_as(ctypes.c_void_p)), ctx)
            self.canonical = self.canonical.astype(bool)

    def sigs(self):
        """the signatures as a list of sorted 1-based id arrays, whatever their lengths (general position: d + 1 each)"""
        if self.sig_off is None:
            return [row for row in np.asarray(self.sig)]
        return [self.sig_ids[self.sig_off[v]:self.sig_off[v + 1]] for v in range(len(self.sig_off) - 1)]

    def origin_of(self, ids):
        """folds extended ids back to caller ids (halo -> the generator it copies; planes -> n_user + p)"""
        ids = np.asarray(ids)
        if not self.n_halo:
            return ids
        out = ids.copy()
        h = (ids > self.n_user) & (ids <= self.n)
        out[h] = self.halo_origin[ids[h] - self.n_user - 1]
        pl = ids > self.n
        out[pl] = ids[pl] - self.n_halo
        return out

    def volumes(self):
        """volumes of the cells of the caller's generators (VoronoiData(...).volume, voronoidata.jl; computed on the device
        from the vertex rows, hvb_cell_volumes); +inf for cells with an unbounded edge"""
        L, ctx = _abi.lib(), self.searcher._ctx
        vol = np.empty((getattr(self, "n_user", self.n),), dtype=np.float64)
        _abi.check(L.hvb_cell_volumes(ctx, vol.ctypes.data_as(ctypes.c_void_p)), ctx)
        return vol

    def moments(self):
        """(vol [n], first [n, d], second [n, d, d]): the integrals of 1, x_a and x_a x_b over every cell, exact
        (VoronoiData(...).bulk_integral for polynomial integrands up to degree two; hvb_cell_moments)"""
        L, ctx = _abi.lib(), self.searcher._ctx
        n, d = getattr(self, "n_user", self.n), self.dim
        vol = np.empty((n,)); first = np.empty((n, d)); second = np.empty((n, d, d))
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _abi.check(L.hvb_cell_moments(ctx, P(vol), P(first), P(second)), ctx)
        return vol, first, second

    def centroids(self):
        """centres of mass of the cells (first moment / volume): the update of a Lloyd step"""
        vol, first, _ = self.moments()
        return first / vol[:, None]

    def areas(self):
        """interface areas aligned with the ids of neighbors() (VoronoiData(...).area; hvb_cell_areas); +inf for facets
        that hold an unbounded edge"""
        L, ctx = _abi.lib(), self.searcher._ctx
        off, ids = self.neighbors()
        area = np.empty((ids.shape[0],), dtype=np.float64)
        _abi.check(L.hvb_cell_areas(ctx, area.ctypes.data_as(ctypes.c_void_p)), ctx)
        return area

    def area_moments(self):
        """(area [m], first [m, d]) aligned with the ids of neighbors(): the integrals of 1 and x_a over every interface
        (VoronoiData(...).interface_integral for integrands up to degree one; hvb_cell_area_moments)"""
        L, ctx = _abi.lib(), self.searcher._ctx
        off, ids = self.neighbors()
        area = np.empty((ids.shape[0],)); first = np.empty((ids.shape[0], self.dim))
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _abi.check(L.hvb_cell_area_moments(ctx, P(area), P(first)), ctx)
        return area, first

    def neighbors(self):
        """CSR (offsets[n+1], ids) of neighbors_of_cell for every cell (neighbors.jl:214-262)."""
        if self._nb is None:
            L, ctx = _abi.lib(), self.searcher._ctx
            if self.copy:
                tot = ctypes.c_int64()
                _abi.check(L.hvb_neighbor_count(ctx, ctypes.byref(tot)), ctx)
                off = np.empty((self.n + 1,), dtype=np.int64)
                ids = np.empty((tot.value,), dtype=np.int64)
                _abi.check(L.hvb_fetch_neighbors(ctx, off.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p)), ctx)
            else:
                po, pi, tot = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
                if getattr(self, "wire32", False):
                    _abi.check(L.hvb_view_neighbors32(ctx, ctypes.byref(po), ctypes.byref(pi), ctypes.byref(tot)), ctx)
                    ids = _wrap(pi, (tot.value,), ctypes.c_int32, self.searcher)
                else:
                    _abi.check(L.hvb_view_neighbors(ctx, ctypes.byref(po), ctypes.byref(pi), ctypes.byref(tot)), ctx)
                    ids = _wrap(pi, (tot.value,), ctypes.c_int64, self.searcher)
                off = _wrap(po, (self.n + 1,), ctypes.c_int64, self.searcher)
            self._nb = (off, ids)
        return self._nb

    def neighbors_of_cell(self, i):
        off, ids = self.neighbors()
        return ids[off[i - 1]:off[i]]

    def vertices_iterator(self, i):
        """all (sig, r) of cell i (1-based), like vertices_iterator(mesh, i) (abstractmesh.jl:179)."""
        if self.sig_off is not None:
            for sg, rr in zip(self.sigs(), self.r):
                if i in sg:
                    yield sg, rr
            return
        if self._cell_index is None:
            real = self.sig <= self.n
            rows = np.repeat(np.arange(self.sig.shape[0]), self.dim + 1)[real.ravel()]
            cells = self.sig.ravel()[real.ravel()]
            order = np.argsort(cells, kind="stable")
            self._cell_index = (np.searchsorted(cells[order], np.arange(1, self.n + 2)), rows[order])
        starts, rows = self._cell_index
        for v in rows[starts[i - 1]:starts[i]]:
            yield self.sig[v], self.r[v]

    def number_of_vertices(self):
        return self.r.shape[0]


def voronoi(xs, searcher=None, Iter=None, copy=True, known=None, **_ignored):
    """voronoi(xs; searcher=Raycast(xs), Iter=1:length(xs)) (sysvoronoi.jl:7-39) -> (mesh, searcher).

    copy=True (default): the mesh owns its arrays.  copy=False returns zero-copy views of the searcher's page-locked
    staging memory, valid only until the next search / set_points / close on that searcher.

    known=(sig, r): vertices the mesh already holds (the reference passes a non-empty mesh in refinement,
    meshrefine.jl:199-215): the walk continues from them and only NEW vertices are returned."""
    if searcher is None:
        searcher = Raycast(xs)
    L, ctx = _abi.lib(), searcher._ctx
    cells_p, ncells = None, 0
    if Iter is not None:
        cells = np.ascontiguousarray(np.asarray(list(Iter), dtype=np.int64))
        cells_p, ncells = cells.ctypes.data_as(ctypes.c_void_p), cells.shape[0]
    if known is None:
        rc = L.hvb_search(ctx, cells_p, ncells, None, None, 0, 0)
    else:
        ksig = np.ascontiguousarray(known[0], dtype=np.int64)
        kr = np.ascontiguousarray(known[1], dtype=np.float64)
        rc = L.hvb_search(ctx, cells_p, ncells, ksig.ctypes.data_as(ctypes.c_void_p), kr.ctypes.data_as(ctypes.c_void_p),
                          ksig.shape[0], ksig.shape[1])
    _abi.check(rc, ctx)
    return VoronoiMesh(searcher, copy=copy), searcher


def refine(searcher, new_xs, old_xs, old_sig, old_r):
    """systematic_refine! (meshrefine.jl:183-216) on the device: new_xs are PREPENDED to the generators as the reference does
    (new ids 1..m, old ids shift by m, plane p = n + p), the old mesh (old_sig, old_r in the OLD numbering) loses the
    vertices whose ball a new node invades (hvb_clean_affected), the new cells are explored (hvb_search with Iter = 1:m).
    Returns (xs_all, sig, r, affected): the vertices of the refined mesh, sorted, and the 1-based ids of the affected cells."""
    new_xs, old_xs = VoronoiNodes(new_xs), VoronoiNodes(old_xs)
    m, n0 = new_xs.shape[0], old_xs.shape[0]
    xs_all = np.ascontiguousarray(np.vstack([new_xs, old_xs]))
    searcher.set_points(xs_all)
    sig = np.ascontiguousarray(np.asarray(old_sig, dtype=np.int64) + m)          # generators and planes shift alike
    r = np.ascontiguousarray(old_r, dtype=np.float64)
    keep = np.zeros(sig.shape[0], dtype=np.uint8)
    affected = np.zeros(n0 + m, dtype=np.uint8)
    L, ctx = _abi.lib(), searcher._ctx
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _abi.check(L.hvb_clean_affected(ctx, P(sig), P(r), sig.shape[0], sig.shape[1], 1, m, P(keep), P(affected)), ctx)
    mesh, _ = voronoi(xs_all, searcher=searcher, Iter=range(1, m + 1), copy=True)
    all_sig = np.vstack([sig[keep.astype(bool)], mesh.sig])
    all_r = np.vstack([r[keep.astype(bool)], mesh.r])
    order = np.lexsort(all_sig.T[::-1])
    return xs_all, all_sig[order], all_r[order], np.nonzero(affected)[0] + 1


class ConvexHull:
    """ConvexHull(xs) (chull.jl:213-238, docs/src/man/convexhull.md): `len(cv)` surface elements, `cv[i] = (sig, r, u)` with
    sig the d generating nodes (1-based, sorted), r a point of the facet's plane and u its outer unit normal.

    General position only.  The hull is wrapped facet by facet on the device (hvb_convex_hull, csrc/hvb_wrap.cuh: one query per
    open ridge, a query streams all generators through shared memory) -- the interior of the tessellation is never computed.
    via="walk": the first device version (around the unbounded 2-faces of the Voronoi diagram with ordinary min-t queries,
    csrc/hvb_hull.cuh); via="search": round 1's way (a complete search on the unbounded domain, facets read off its unbounded
    edges).  Same result all three ways; the tests use the other two as cross-checks."""

    def __init__(self, xs, intro="", nthreads=None, method=None, options=None, via="wrap", searcher=None):
        xs = VoronoiNodes(xs)
        # searcher: a Raycast(xs, domain=Boundary()) to reuse (a warm context allocates nothing); it stays open
        s = searcher if searcher is not None else Raycast(xs, domain=Boundary(), options=options or RaycastParameter())
        try:
            if via == "search":
                mesh, _ = voronoi(xs, searcher=s, copy=True)
                edge, udir, base = mesh.ray_edge, mesh.ray_dir, mesh.ray_base
            else:
                L, ctx = _abi.lib(), s._ctx
                _abi.check(L.hvb_convex_hull_via(ctx, {"wrap": 0, "walk": 1}[via]), ctx)
                nv, nr = ctypes.c_int64(), ctypes.c_int64()
                _abi.check(L.hvb_counts(ctx, ctypes.byref(nv), ctypes.byref(nr), None), ctx)
                d = xs.shape[1]
                edge = np.empty((nr.value, d), dtype=np.int64); base = np.empty((nr.value, d)); udir = np.empty((nr.value, d))
                node = np.empty((nr.value,), dtype=np.int64)
                if nr.value:
                    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
                    _abi.check(L.hvb_fetch_rays(ctx, P(edge), P(base), P(udir), P(node)), ctx)
            self.stats = s.stats()
            order = np.lexsort(edge.T[::-1]) if len(edge) else np.zeros(0, dtype=np.int64)
            self.sig = edge[order]
            self.u = udir[order]
            base = base[order]
        finally:
            if searcher is None:
                s.close()
        self.xs = xs
        # the reference projects the stored point onto the facet's plane (chull.jl:224-232)
        self.r = base + self.u * ((xs[self.sig[:, 0] - 1] - base) * self.u).sum(axis=1)[:, None] if len(self.sig) else base

    def __len__(self):
        return self.sig.shape[0]

    def __getitem__(self, i):
        return self.sig[i].copy(), self.r[i].copy(), self.u[i].copy()

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class VoronoiGeometry:
    """VoronoiGeometry(xs, b; search_settings=(...)) (geometry.jl:139-201), restricted to what this path produces:
    the vertex database and the neighbour lists (integrate=false)."""

    def __init__(self, xs, b=None, search_settings=None, **_ignored):
        xs = VoronoiNodes(xs)
        b = b if b is not None else Boundary()
        opts = RaycastParameter(**(search_settings or {}))
        # a domain with periodic planes is periodised by the backend (the reference: Create_Discrete_Domain, domain.jl:175)
        self.searcher = Raycast(xs, domain=b, options=opts, periodic=len(b.periodic) > 0)
        self.mesh, _ = voronoi(xs, searcher=self.searcher)
        self.nodes = xs
        self.domain = b


def _split(a, off, n):
    return [a[off[i]:off[i + 1]] for i in range(n)]


def orientations_of(xs, off, ids, n_user, halo_xs, base, normal):
    """`orientations` of VoronoiData (voronoidata.jl:593-594): for every neighbour ids[k] of cell i the vector from generator i
    to that neighbour -- the generator itself, the periodic IMAGE the cell really touches (halo ids n_user+1..n_user+n_halo: the
    case the reference's docstring calls tricky by hand), or the mirror image of x_i behind a boundary plane (reflect,
    boundary.jl:206-211) -- so that x_i + orientation / 2 lies on the interface (voronoidata.jl:788).  [m, d], aligned with ids."""
    xs = np.asarray(xs)
    ids = np.asarray(ids, dtype=np.int64)
    off = np.asarray(off, dtype=np.int64)
    d = xs.shape[1]
    n_halo = 0 if halo_xs is None else len(halo_xs)
    cell = np.repeat(np.arange(n_user), np.diff(off[:n_user + 1]))
    ids = ids[off[0]:off[n_user]]
    x0 = xs[cell]
    out = np.empty((ids.shape[0], d))
    g = ids <= n_user
    out[g] = xs[ids[g] - 1] - x0[g]
    h = (ids > n_user) & (ids <= n_user + n_halo)
    if h.any():
        out[h] = np.asarray(halo_xs)[ids[h] - n_user - 1] - x0[h]
    pl = ids > n_user + n_halo
    if pl.any():
        p = ids[pl] - n_user - n_halo - 1
        nrm = np.asarray(normal)[p]
        nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        dist = ((np.asarray(base)[p] - x0[pl]) * nrm).sum(1)
        out[pl] = 2.0 * dist[:, None] * nrm
    return out


def boundary_nodes_of(xs, off, ids, n_user, n_halo, base, normal, onboundary=False):
    """`boundary_nodes` of VoronoiData (voronoidata.jl:596-598): {i: {p: mirrored generator i}} for every cell i (1-based) that
    touches boundary plane p (1-based); onboundary=True gives the projection of x_i onto the plane instead."""
    ori = orientations_of(xs, off, ids, n_user, np.zeros((n_halo, np.asarray(xs).shape[1])), base, normal)
    out = {}
    ids = np.asarray(ids, dtype=np.int64)
    for i in range(n_user):
        for k in range(int(off[i]), int(off[i + 1])):
            if ids[k] > n_user + n_halo:
                out.setdefault(i + 1, {})[int(ids[k] - n_user - n_halo)] = np.asarray(xs)[i] + (0.5 if onboundary else 1.0) * ori[k - int(off[0])]
    return out


class VoronoiData:
    """VoronoiData(VG; getFIELD...) (voronoidata.jl:545-703, docstring :571-620): the fields of the reference's data view that
    this path and its geometry products fill.  Every field is a hard copy (the reference's `getFIELD=true`).

    nodes, vertices, boundary_vertices (edge => (base, direction, node), voronoidata.jl:580-582), neighbors, orientations, volume,
    area, boundary_nodes, bulk_integral, interface_integral, references / reference_shifts / offset (reduce_to_periodic=False).

    Integrals: the reference integrates a Julia closure; across the C ABI the integrands are the monomials up to degree two
    (bulk: 1, x_a, x_a x_b -> bulk_integral[i] of length 1 + d + d*d) and up to degree one on interfaces (1, x_a ->
    interface_integral[i][k] of length 1 + d), exact (hvb_cell_moments, hvb_cell_area_moments).

    Periodic domains, reduce_to_periodic=True (default, as in the reference): neighbours are folded back to the caller's ids, a
    node may appear several times (voronoidata.jl:589), sorted=True orders them with their areas / integrals / orientations.
    reduce_to_periodic=False shows the halo: ids n+1..n+offset are the periodic copies (the reference numbers them 1..offset IN
    FRONT of the official nodes; here they stay behind them, as the backend numbers them), references[k] is the official node
    halo node k copies and reference_shifts[k] the shift: node[n + k] = node[references[k]] + reference_shifts[k]."""

    def __init__(self, VG, getvertices=False, getneighbors=False, getvolume=False, getarea=False, getorientations=False,
                 getboundary_vertices=False, getboundary_nodes=False, getbulk_integral=False, getinterface_integral=False,
                 getreferences=False, getreference_shifts=False, copyall=False, reduce_to_periodic=True, onboundary=False,
                 sorted=False, **_ignored):
        self.nodes = VG.nodes
        self.geometry = VG
        m = VG.mesh
        nu = getattr(m, "n_user", m.n)
        n_halo = getattr(m, "n_halo", 0)
        dom = VG.domain
        want = lambda f: bool(f) or bool(copyall)
        self.boundary = dom
        self.boundary_nodes_on_boundary = bool(onboundary)
        self.offset = 0 if reduce_to_periodic else n_halo
        if want(getvolume):
            self.volume = m.volumes()
        if want(getvertices):
            self.vertices = [list(m.vertices_iterator(i)) for i in range(1, nu + 1)]
        if want(getboundary_vertices):
            self.boundary_vertices = {tuple(int(g) for g in e): (np.array(b), np.array(u), int(nd))
                                      for e, b, u, nd in zip(m.ray_edge, m.ray_base, m.ray_dir, m.ray_node)}
        if n_halo and (want(getreferences) or want(getreference_shifts)):
            self.references = np.array(m.halo_origin)
            self.reference_shifts = np.asarray(m.halo_xs) - np.asarray(VG.nodes)[np.asarray(m.halo_origin) - 1]
        if want(getbulk_integral):
            vol, first, second = m.moments()
            self.bulk_integral = np.concatenate([vol[:, None], first, second.reshape(len(vol), -1)], axis=1)
            if want(getvolume):
                self.volume = vol
        per_nb = [want(getneighbors), want(getarea), want(getorientations), want(getinterface_integral), want(getboundary_nodes)]
        if not any(per_nb):
            return
        off, ids = m.neighbors()
        off = np.asarray(off, dtype=np.int64); ids = np.asarray(ids, dtype=np.int64)
        lo, hi = int(off[0]), int(off[nu])
        shown = ids[lo:hi]
        if reduce_to_periodic and n_halo:
            # periodic copies fold back to the node they copy, planes follow the official nodes (voronoidata.jl:623)
            shown = m.origin_of(shown)
        order = None
        if sorted and reduce_to_periodic and n_halo:
            cell = np.repeat(np.arange(nu), np.diff(off[:nu + 1]))
            order = np.lexsort((shown, cell))
            shown = shown[order]
        pick = (lambda a: a[order]) if order is not None else (lambda a: a)
        loc = off[:nu + 1] - lo
        if want(getneighbors):
            self.neighbors = _split(shown, loc, nu)
        if want(getarea) and not want(getinterface_integral):
            self.area = _split(pick(np.asarray(m.areas())[lo:hi]), loc, nu)
        if want(getinterface_integral):
            a, first = m.area_moments()
            if want(getarea):
                self.area = _split(pick(np.asarray(a)[lo:hi]), loc, nu)
            self.interface_integral = _split(pick(np.concatenate([np.asarray(a)[lo:hi, None], np.asarray(first)[lo:hi]], axis=1)), loc, nu)
        halo_xs = getattr(m, "halo_xs", None) if n_halo else None
        if want(getorientations):
            self.orientations = _split(pick(orientations_of(VG.nodes, off, ids, nu, halo_xs, dom.base, dom.normal)), loc, nu)
        if want(getboundary_nodes):
            self.boundary_nodes = boundary_nodes_of(VG.nodes, off, ids, nu, n_halo, dom.base, dom.normal, onboundary)
