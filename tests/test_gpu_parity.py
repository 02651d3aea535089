"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the oracle on the same seeded
inputs, against the committed golden vectors, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Bar: bit-exact signatures and neighbour lists, coordinates within 1e-10 relative."""
import glob
import os

import numpy as np
import pytest

import qhull_oracle
from util import assert_same_mesh, empty_ball_violations, neighbors_from_sig, points

pytestmark = pytest.mark.gpu
COORD_TOL = 1e-10          # north_star: "Vertex coordinates must agree within 1e-10 relative"
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def run_gpu(hvb, xs, bounded=True, **settings):
    d = xs.shape[1]
    dom = hvb.cuboid(d, periodic=[]) if bounded else hvb.Boundary()
    s = hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(**settings))
    mesh, _ = hvb.voronoi(xs, searcher=s)
    return mesh, s


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors(hvb, path):
    g = np.load(path)
    dom = hvb.Boundary(g["base"], g["normal"]) if g["base"].shape[0] else hvb.Boundary()
    s = hvb.Raycast(g["xs"], domain=dom)
    mesh, _ = hvb.voronoi(g["xs"], searcher=s)
    assert_same_mesh(mesh.sig, mesh.r, g["sig"], g["r"], g["xs"], COORD_TOL)
    assert sorted(map(tuple, mesh.ray_edge.tolist())) == sorted(map(tuple, g["ray_edge"].tolist()))
    off, ids = mesh.neighbors()
    assert np.array_equal(off, g["nb_off"]) and np.array_equal(ids, g["nb_ids"])


# configs[0] of BASELINE.json (C1) and reduced sizes of C3/C4/C5 that the oracle finishes in seconds
@pytest.mark.parametrize("d,n,bounded", [(3, 1000, True), (3, 1000, False), (2, 20000, True), (2, 5000, False),
                                          (4, 2000, True), (4, 500, False), (5, 1000, True), (5, 300, False),
                                          (6, 300, True), (6, 120, False), (3, 20000, True)])
def test_matches_oracle(hvb, oracle, d, n, bounded):
    xs = points(n, d, 1000 + 10 * d + int(bounded))
    if bounded:
        base, normal = qhull_oracle.cuboid(d)
        o = oracle.run(xs, base, normal)
    else:
        o = oracle.run(xs)
    mesh, s = run_gpu(hvb, xs, bounded)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)
    assert sorted(map(tuple, mesh.ray_edge.tolist())) == sorted(map(tuple, o["ray_edge"].tolist()))
    off, ids = mesh.neighbors()
    assert np.array_equal(off, o["nb_off"]) and np.array_equal(ids, o["nb_ids"])
    st = s.stats()
    assert st["degenerate"] == 0 and st["vertices"] == len(o["sig"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c1_seeds(hvb, oracle, seed):
    xs = points(1000, 3, seed)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    mesh, _ = run_gpu(hvb, xs, True)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)


def test_matches_qhull_directly(hvb):
    """independent of the restatement: Delaunay circumcentres + per-cell half-space intersection"""
    xs = points(600, 3, 77)
    base, normal = qhull_oracle.cuboid(3)
    q = qhull_oracle.bounded(xs, base, normal)
    mesh, _ = run_gpu(hvb, xs, True)
    got = [tuple(s) for s in mesh.sig.tolist()]
    assert set(got) == set(q) and len(got) == len(q)
    assert max(np.abs(mesh.r[k] - q[s]).max() for k, s in enumerate(got)) < 1e-11


@pytest.mark.parametrize("key", ["xs_1161", "xs_1282"])
def test_walks_from_vertices_1e7_diameters_away(hvb, key):
    """unbounded clouds whose density varies by ten orders of magnitude along one axis: flat hull simplices put vertices
    1e7 ... 1e10 cloud diameters away (tests/test_hostsim.py has the story of the regression); against Qhull"""
    xs = np.load(os.path.join(os.path.dirname(__file__), "golden", "clouds", "far_vertices.npz"))[key]
    truth, rays = qhull_oracle.unbounded(xs)
    mesh, s = run_gpu(hvb, xs, False)
    assert {tuple(r) for r in mesh.sig.tolist()} == set(truth) and len(mesh.sig) == len(truth)
    assert {tuple(r) for r in mesh.ray_edge.tolist()} == rays


def test_fp32_filter_equals_fp64_only(hvb):
    xs = points(30000, 3, 3)
    a, sa = run_gpu(hvb, xs, True, fp32_filter=1)
    b, sb = run_gpu(hvb, xs, True, fp32_filter=0)
    assert np.array_equal(a.sig, b.sig) and np.array_equal(a.r, b.r)
    assert sa.stats()["candidates_fp64"] < sb.stats()["candidates_fp64"]


def test_deterministic_and_knob_independent(hvb):
    xs = points(20000, 3, 4)
    a, _ = run_gpu(hvb, xs, True)
    b, _ = run_gpu(hvb, xs, True)
    c, _ = run_gpu(hvb, xs, True, points_per_cell=6, seed_stride=3, probe_scale=2.0)
    assert np.array_equal(a.sig, b.sig) and np.array_equal(a.r, b.r)       # bitwise, run to run
    assert np.array_equal(a.sig, c.sig) and np.array_equal(a.r, c.r)       # canonical coordinates


def test_iter_subset_finds_all_vertices_of_the_cells(hvb, oracle):
    xs = points(3000, 3, 8)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    cells = np.arange(1, 301)
    s = hvb.Raycast(xs, domain=hvb.cuboid(3, periodic=[]))
    mesh, _ = hvb.voronoi(xs, searcher=s, Iter=cells)
    want = {tuple(r) for r in o["sig"].tolist() if any(1 <= g <= 300 for g in r)}
    got = {tuple(r) for r in mesh.sig.tolist()}
    assert want <= got <= {tuple(r) for r in o["sig"].tolist()}


@pytest.mark.parametrize("nb_in_search,decomposition,world", [(1, 1, 4), (0, 1, 4), (1, 0, 4), (1, 1, 8), (1, 1, 6)])
def test_slab_union_equals_full(hvb, oracle, nb_in_search, decomposition, world):
    """the multi-GPU decomposition on one device: the union of the slab searches is the full vertex set, and every rank
    holds the complete neighbour lists of the cells it owns (empty lists elsewhere)"""
    xs = points(6000, 3, 10)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    rows, total = set(), 0
    owned_by = np.zeros(6000, dtype=int)
    for rank in range(world):
        s = hvb.Raycast(xs, domain=hvb.cuboid(3, periodic=[]),
                        options=hvb.RaycastParameter(threading=hvb.B200Thread(0, rank, world), neighbors=nb_in_search, decomposition=decomposition))
        mesh, _ = hvb.voronoi(xs, searcher=s)
        assert (np.diff(mesh.sig[:, 0]) >= 0).all()                      # each shard is sorted
        rows |= {tuple(r) for r in mesh.sig.tolist()}
        total += mesh.sig.shape[0]
        assert s.stats()["raycasts"] < 0.9 * len(o["sig"])               # a rank walks its part (plus a layer around it) only
        own = s.owned()
        assert 0.6 * 6000 / world <= own.sum() <= 1.6 * 6000 / world      # equal counts per part (quantile cuts / equal slabs)
        owned_by += own
        off, ids = mesh.neighbors()
        for c in range(6000):
            mine, ref = ids[off[c]:off[c + 1]], o["nb_ids"][o["nb_off"][c]:o["nb_off"][c + 1]]
            if own[c]:
                assert np.array_equal(mine, ref), (rank, c)
            else:
                assert len(mine) == 0, (rank, c)
    assert (owned_by == 1).all()                                         # the slabs partition the cells
    assert rows == {tuple(r) for r in o["sig"].tolist()}
    assert total == len(o["sig"])                                        # ownership rule: the shards are disjoint


def test_capacity_retry_path(hvb, oracle):
    """a far too small vertex capacity must be grown transparently (tables re-allocated, search restarted)"""
    xs = points(3000, 3, 12)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    for persistent in (4, 3, 2, 1, 0):
        mesh, s = run_gpu(hvb, xs, True, vertex_capacity=700, persistent=persistent)
        assert np.array_equal(mesh.sig, o["sig"])
        assert s.stats()["capacity_retries"] >= 3


@pytest.mark.parametrize("persistent", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("tile", [1, 2, 4, 8, 16, 32])
def test_every_tile_size_and_both_walk_modes(hvb, oracle, tile, persistent):
    xs = points(4000, 3, 13)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    mesh, _ = run_gpu(hvb, xs, True, tile_size=tile, persistent=persistent)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)


def test_int32_wire_format_equals_int64(hvb):
    """wire32: signatures and neighbour ids staged as int32 (hvb_view_vertices32 / hvb_view_neighbors32), same content;
    the int64 calls keep working on the same context"""
    xs = points(8000, 3, 17)
    dom = hvb.cuboid(3, periodic=[])
    a, _ = hvb.voronoi(xs, searcher=hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(neighbors=1)))
    s = hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(wire32=1, neighbors=1))
    b, _ = hvb.voronoi(xs, searcher=s, copy=False)
    assert b.sig.dtype == np.int32 and np.array_equal(b.sig, a.sig) and np.array_equal(b.r, a.r)
    off, ids = b.neighbors()
    off0, ids0 = a.neighbors()
    assert ids.dtype == np.int32 and np.array_equal(off, off0) and np.array_equal(ids, ids0)
    c = hvb.VoronoiMesh(s, copy=True)                     # int64 fetch on the wire32 context
    assert c.sig.dtype == np.int64 and np.array_equal(c.sig, a.sig)
    off2, ids2 = c.neighbors()
    assert np.array_equal(ids2, ids0)


def test_context_reuse_set_points(hvb, oracle):
    """hvb_set_points: one context, several clouds of different sizes, results independent of the history"""
    base, normal = qhull_oracle.cuboid(3)
    s = hvb.Raycast(points(2000, 3, 14), domain=hvb.cuboid(3, periodic=[]))
    for n, seed in [(2000, 14), (9000, 15), (500, 16), (9000, 15)]:
        xs = points(n, 3, seed)
        s.set_points(xs)
        mesh, _ = hvb.voronoi(xs, searcher=s)
        o = oracle.run(xs, base, normal)
        assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)
        off, ids = mesh.neighbors()
        assert np.array_equal(off, o["nb_off"]) and np.array_equal(ids, o["nb_ids"])


@pytest.mark.parametrize("d,n", [(3, 4000), (2, 6000), (4, 800)])
def test_known_vertices_are_continued_not_returned(hvb, oracle, d, n):
    """hvb_search with seed vertices (a mesh that already holds vertices, meshrefine.jl:199-215): the walk continues
    from them, returns exactly the missing ones, and the neighbour lists cover old and new vertices"""
    xs = points(n, d, 40 + d)
    base, normal = qhull_oracle.cuboid(d)
    o = oracle.run(xs, base, normal)
    rng = np.random.default_rng(1)
    keep = rng.random(len(o["sig"])) < 0.4
    keep[(o["sig"] <= n // 3).any(axis=1)] = True            # a whole region is already meshed
    s = hvb.Raycast(xs, domain=hvb.cuboid(d, periodic=[]))
    mesh, _ = hvb.voronoi(xs, searcher=s, known=(o["sig"][keep], o["r"][keep]))
    assert np.array_equal(mesh.sig, o["sig"][~keep])
    assert_same_mesh(mesh.sig, mesh.r, o["sig"][~keep], o["r"][~keep], xs, COORD_TOL)
    off, ids = mesh.neighbors()
    assert np.array_equal(off, o["nb_off"]) and np.array_equal(ids, o["nb_ids"])
    # unused trailing entries (0) and a wider stride are accepted, malformed rows are refused
    wide = np.zeros((int(keep.sum()), d + 3), dtype=np.int64)
    wide[:, :d + 1] = o["sig"][keep]
    mesh2, _ = hvb.voronoi(xs, searcher=s, known=(wide, o["r"][keep]))
    assert np.array_equal(mesh2.sig, o["sig"][~keep])
    bad = o["sig"][keep].copy()
    bad[0, 1] = bad[0, 0]
    with pytest.raises(hvb.HVBError):
        hvb.voronoi(xs, searcher=s, known=(bad, o["r"][keep]))


def test_points_outside_domain_are_rejected(hvb):
    xs = points(100, 3, 0)
    xs[5, 1] = 1.5
    with pytest.raises(hvb.HVBError) as e:
        hvb.Raycast(xs, domain=hvb.cuboid(3, periodic=[]))
    assert e.value.code == hvb._abi.HVB_EINVAL and "does not lie in the domain" in str(e.value)


def test_degenerate_input_is_reported(hvb):
    g = np.stack(np.meshgrid(*[np.arange(6.0)] * 3, indexing="ij"), -1).reshape(-1, 3) / 6 + 1 / 12
    with pytest.raises(hvb.HVBError) as e:
        run_gpu(hvb, g, True, on_degenerate=0)
    assert e.value.code == hvb._abi.HVB_EDEGENERATE


# ---- BASELINE.json full sizes against the oracle itself (multi-threaded restatement, ~10 s each) -------------------
@pytest.mark.parametrize("d,n", [(3, 100000), (2, 300000), (5, 5000), (4, 20000)])
def test_full_size_matches_oracle(hvb, oracle, d, n):
    xs = points(n, d, 0)
    base, normal = qhull_oracle.cuboid(d)
    o = oracle.run(xs, base, normal, nthreads=min(8, os.cpu_count() or 1))
    mesh, s = run_gpu(hvb, xs, True)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)
    off, ids = mesh.neighbors()
    assert np.array_equal(off, o["nb_off"]) and np.array_equal(ids, o["nb_ids"])


# ---- every BASELINE.json config at its OWN size, exact parity against the 8-thread restatement ----------------------
# C3 (1 000 000 points, d = 2), C4 (50 000 points, d = 5) and d = 6 at 2 000 points (C5's dimension; its periodic form is
# covered in test_gpu_periodic.py).  The oracle needs 17 s / 110 s / 20 s on 8 host threads.
@pytest.mark.parametrize("d,n", [(2, 1000000), (5, 50000), (6, 2000)])
def test_config_size_matches_oracle(hvb, oracle, d, n):
    xs = points(n, d, 0)
    base, normal = qhull_oracle.cuboid(d)
    o = oracle.run(xs, base, normal, nthreads=min(8, os.cpu_count() or 1))
    mesh, s = run_gpu(hvb, xs, True)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)
    if d <= 5:
        off, ids = mesh.neighbors()
        assert np.array_equal(off, o["nb_off"]) and np.array_equal(ids, o["nb_ids"])
    st = s.stats()
    assert st["degenerate"] == 0 and st["rejected"] == 0 and st["vertices"] == len(o["sig"])


# ---- the search_settings raycast methods (test/rcmethods.jl:4-13: d = 4, 1000 points, cuboid, |sum(volume) - 1| < 1e-3) ----
@pytest.mark.parametrize("method", ["RCCombined", "RCOriginal", "RCNonGeneralFast", "RCNonGeneral", "RCOriginalHP", "RCNonGeneralSkip"])
def test_rcmethods(hvb, oracle, method):
    xs = points(1000, 4, 99)
    base, normal = qhull_oracle.cuboid(4)
    # against the same method of the restated reference where it restates it (the four of test/rcmethods.jl), else the default
    o = oracle.run(xs, base, normal, method=method if method in oracle.METHODS else 0)
    mesh, s = run_gpu(hvb, xs, True, method=getattr(hvb, method))
    assert s.parameters.method == getattr(hvb, method)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"], o["r"], xs, COORD_TOL)     # the winner does not depend on the procedure
    vol = mesh.volumes()
    assert abs(vol.sum() - 1.0) < 1e-3                                       # the reference's own bar
    assert abs(vol.sum() - 1.0) < 1e-10


def test_unknown_method_and_bad_tolerances_are_refused(hvb):
    xs = points(100, 3, 0)
    for kw in (dict(method=8), dict(method=-1), dict(break_tol=0.0), dict(variance_tol=-1.0), dict(ray_tol=float("nan"))):
        with pytest.raises(hvb.HVBError) as e:
            run_gpu(hvb, xs, True, **kw)
        assert e.value.code == hvb._abi.HVB_EINVAL


def test_break_tol_and_variance_tol_are_honoured(hvb):
    """walkray_correct_vertex (raycast.jl:257-279): vertices whose squared radii vary by more than break_tol (relative
    variance) are dropped and counted, those above variance_tol are kept and counted"""
    xs = points(4000, 3, 21)
    a, sa = run_gpu(hvb, xs, True)
    assert sa.stats()["rejected"] == 0 and sa.stats()["suboptimal"] == 0
    # canonical coordinates leave a relative variance of ~1e-31 .. 1e-28: thresholds inside that range split the rows
    b, sb = run_gpu(hvb, xs, True, variance_tol=1e-40, break_tol=1e-5)
    assert sb.stats()["rejected"] == 0 and sb.stats()["suboptimal"] > 0
    assert np.array_equal(a.sig, b.sig)
    c, sc = run_gpu(hvb, xs, True, variance_tol=1e-40, break_tol=1e-31)
    rej = sc.stats()["rejected"]
    assert 0 < rej < len(a.sig) and len(c.sig) == len(a.sig) - rej
    assert {tuple(r) for r in c.sig.tolist()} <= {tuple(r) for r in a.sig.tolist()}


@pytest.mark.parametrize("eps,degenerate", [(0.0, True), (1e-14, True), (1e-7, False)])
def test_near_degenerate_input_is_reported_like_the_reference(hvb, eps, degenerate):
    """five generators on a common sphere: the reference appends every candidate inside its tie window to the signature
    (raycast.jl:902,926-949, a vertex with d + 2 generators); this backend reports exactly those inputs as non-general
    position and treats a 1e-7 perturbation as general, as the reference does"""
    rng = np.random.default_rng(5)
    u = rng.normal(size=(5, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    cos = 0.5 + 0.25 * u                                   # on the sphere of radius 0.25 around the centre of the cube
    cos[4] = 0.5 + 0.25 * (1.0 + eps) * u[4]
    far = 0.5 + 0.49 * np.sign(rng.normal(size=(40, 3))) * (0.8 + 0.2 * rng.random((40, 3)))   # keep the ball empty
    xs = np.vstack([cos, far])
    if degenerate:
        with pytest.raises(hvb.HVBError) as e:
            run_gpu(hvb, xs, True, on_degenerate=0)
        assert e.value.code == hvb._abi.HVB_EDEGENERATE
    else:
        mesh, s = run_gpu(hvb, xs, True)
        assert s.stats()["degenerate"] == 0 and len(mesh.sig) > 0


# ---- BASELINE.json full sizes: properties that need no oracle run --------------------------------------------
@pytest.mark.parametrize("d,n,vpp_lo,vpp_hi", [(3, 100000, 6.4, 7.1), (2, 1000000, 1.95, 2.05), (5, 50000, 60, 190)])
def test_full_size_properties(hvb, d, n, vpp_lo, vpp_hi):
    xs = points(n, d, 0)
    mesh, s = run_gpu(hvb, xs, True)
    V = mesh.sig.shape[0]
    assert vpp_lo <= V / n <= vpp_hi, V / n
    # rows sorted, unique, every row sorted and holding at least one real generator
    assert (np.diff(mesh.sig, axis=1) > 0).all() and (mesh.sig[:, 0] <= n).all()
    key = mesh.sig[1:] != mesh.sig[:-1]
    first = key.argmax(axis=1)
    assert key.any(axis=1).all()
    assert (np.take_along_axis(mesh.sig[1:], first[:, None], 1) > np.take_along_axis(mesh.sig[:-1], first[:, None], 1)).all()
    # verify_vertex (raycast.jl:477-502) on a sample: empty ball, equidistance; all vertices inside the domain
    assert empty_ball_violations(mesh.sig, mesh.r, xs, sample=1500) == 0
    assert mesh.r.min() >= -1e-12 and mesh.r.max() <= 1 + 1e-12
    # every cell has vertices, and the neighbour lists equal the host recomputation
    assert len(np.unique(mesh.sig[mesh.sig <= n])) == n
    if d <= 3:
        off, ids = mesh.neighbors()
        off2, ids2 = neighbors_from_sig(mesh.sig, n)
        assert np.array_equal(off, off2) and np.array_equal(ids, ids2)
    st = s.stats()
    assert st["degenerate"] == 0 and st["capacity_retries"] == 0
