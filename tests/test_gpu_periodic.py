"""Periodic domains on the GPU (hvb_create_periodic): parity with the CPU restatement on the explicit halo problem the
library itself built, with Qhull on the 3^k replication, and size-independent invariants at larger sizes."""
import ctypes

import numpy as np
import pytest

import periodic_oracle as po
from util import assert_same_mesh, empty_ball_violations, points

pytestmark = pytest.mark.gpu
COORD_TOL = 1e-10


def run_periodic(hvb, xs, axes, **settings):
    d = xs.shape[1]
    s = hvb.Raycast(xs, domain=hvb.cuboid(d, periodic=list(axes)), options=hvb.RaycastParameter(**settings), periodic=True)
    mesh, _ = hvb.voronoi(xs, searcher=s)
    return mesh, s


@pytest.mark.parametrize("d,n,axes", [(2, 2000, (1, 2)), (2, 1500, (2,)), (3, 800, (1, 2, 3)), (3, 800, (1, 3)), (4, 600, (1, 2, 3, 4)),
                                      (5, 400, (1, 2)), (5, 1500, (2, 5)), (6, 2000, (1,))])
def test_periodic_matches_oracle_on_the_halo_problem(hvb, oracle, d, n, axes):
    xs = points(n, d, 300 + d)
    mesh, s = run_periodic(hvb, xs, axes)
    origin, mult, hxs, margin = mesh.halo_origin, mesh.halo_mult, mesh.halo_xs, mesh.margin
    # the halo the library built is the documented one
    o2, m2, h2 = po.halo(xs, axes, margin)
    assert np.array_equal(origin, o2) and np.array_equal(mult, m2) and np.array_equal(hxs, h2)
    # the reference algorithm on caller + halo generators inside the pushed planes, rows that touch caller generators
    ext = np.vstack([xs, hxs])
    base, normal = po.pushed_cuboid(d, axes, margin)
    o = oracle.run(ext, base, normal, nthreads=8)
    touch = (o["sig"] <= n).any(axis=1)
    assert_same_mesh(mesh.sig, mesh.r, o["sig"][touch], o["r"][touch], ext, COORD_TOL)
    # neighbour lists of the caller cells
    off, ids = mesh.neighbors()
    assert off.shape[0] == ext.shape[0] + 1
    assert np.array_equal(off[:n + 1], o["nb_off"][:n + 1]) and np.array_equal(ids[:off[n]], o["nb_ids"][:o["nb_off"][n]])
    st = s.stats()
    assert st["halo_nodes"] == len(origin) and st["vertices"] == int(touch.sum()) and st["degenerate"] == 0
    assert st["unique_vertices"] == int(mesh.canonical.sum())
    if len(axes) == d:
        assert (mesh.sig <= ext.shape[0]).all()                    # no plane takes part in a fully periodic tessellation


@pytest.mark.parametrize("d,n", [(2, 3000), (3, 1000), (4, 500)])
def test_periodic_matches_qhull_torus(hvb, d, n):
    xs = points(n, d, 310 + d)
    axes = tuple(range(1, d + 1))
    mesh, s = run_periodic(hvb, xs, axes)
    got = po.fold_rows(mesh.sig, n, mesh.halo_origin, mesh.halo_mult)
    want = po.torus_simplices(xs, axes)
    assert got == want
    assert int(mesh.canonical.sum()) == po.canonical_classes(want)
    if d == 2:
        assert int(mesh.canonical.sum()) == 2 * n                   # Euler: a triangulated torus has 2n triangles


def test_margin_retry_path(hvb):
    """a first margin that is far too small must be corrected by the certificate, with the same final result"""
    xs = points(3000, 2, 320)
    a, sa = run_periodic(hvb, xs, (1, 2))
    b, sb = run_periodic(hvb, xs, (1, 2), periodic_margin=0.004)
    assert sa.stats()["periodic_retries"] == 0 and sb.stats()["periodic_retries"] >= 1
    fa = po.fold_rows(a.sig, 3000, a.halo_origin, a.halo_mult)
    fb = po.fold_rows(b.sig, 3000, b.halo_origin, b.halo_mult)
    assert fa == fb and int(b.canonical.sum()) == 6000


@pytest.mark.parametrize("d,n,vpp", [(2, 400000, 2.0), (3, 60000, 6.768), (5, 6000, None), (6, 2000, None)])
def test_periodic_invariants_at_size(hvb, d, n, vpp):
    xs = points(n, d, 0)
    axes = tuple(range(1, d + 1))
    mesh, s = run_periodic(hvb, xs, axes)
    st = s.stats()
    ne = n + mesh.n_halo
    assert (mesh.sig <= ne).all() and (np.diff(mesh.sig, axis=1) > 0).all()
    assert (mesh.sig <= n).any(axis=1).all()                        # every row touches a caller generator
    uniq = int(mesh.canonical.sum())
    assert uniq == st["unique_vertices"]
    if d == 2:
        assert uniq == 2 * n
    elif vpp:
        assert abs(uniq / n - vpp) < 0.03 * vpp, uniq / n           # Poisson-Delaunay mean, no boundary effects on a torus
    # each caller generator: the canonical images counted with multiplicity d+1 cover all cells equally
    ext = np.vstack([xs, mesh.halo_xs])
    assert empty_ball_violations(mesh.sig, mesh.r, ext, sample=1200) == 0
    # every image class has exactly one canonical member: folding canonical rows gives `uniq` distinct origin sets
    orig = np.sort(mesh.origin_of(mesh.sig[mesh.canonical]), axis=1)
    assert len(np.unique(orig, axis=0)) >= 0.999 * uniq
    # sum over caller cells of (vertices of the cell) = (d+1) * unique vertices
    assert int((mesh.sig <= n).sum()) == (d + 1) * uniq
    # folded neighbour relation is symmetric
    off, ids = mesh.neighbors()
    deg = np.diff(off[:n + 1])
    a = np.repeat(np.arange(1, n + 1), deg)
    b = mesh.origin_of(ids[:off[n]])
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    back = np.unique(np.stack([pairs[:, 1], pairs[:, 0]], 1), axis=0)
    assert np.array_equal(pairs, back)


def test_periodic_argument_errors(hvb):
    L = hvb._abi.lib()
    xs = points(50, 2, 1)
    b = hvb.cuboid(2, periodic=[])
    ctx = ctypes.c_void_p()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    bad = np.array([2, 0, 0, 0], dtype=np.int32)                    # plane 1 names plane 2, plane 2 does not answer
    assert L.hvb_create_periodic(ctypes.byref(ctx), 2, 50, P(xs), 4, P(b.base), P(b.normal), P(bad), None) == hvb._abi.HVB_EINVAL
    bad = np.array([3, 0, 1, 0], dtype=np.int32)                    # partners that are not parallel
    assert L.hvb_create_periodic(ctypes.byref(ctx), 2, 50, P(xs), 4, P(b.base), P(b.normal), P(bad), None) == hvb._abi.HVB_EINVAL
    # too few generators for the period: a cell would neighbour its own image (three clustered points: their cells
    # are strips that wrap around the torus; three well spread points would NOT do, their torus triangulation is proper)
    with pytest.raises(hvb.HVBError) as e:
        run_periodic(hvb, 0.4 + 0.05 * points(3, 2, 2), (1, 2))
    assert e.value.code in (hvb._abi.HVB_EINCOMPLETE, hvb._abi.HVB_EINVAL)
    # periodic contexts do not take seed vertices
    s = hvb.Raycast(points(500, 2, 3), domain=hvb.cuboid(2), periodic=True)
    with pytest.raises(hvb.HVBError):
        hvb.voronoi(s.xs, searcher=s, known=(np.array([[1, 2, 3]]), np.zeros((1, 2))))


def test_voronoi_geometry_periodic_front_end(hvb):
    xs = points(1200, 3, 5)
    VG = hvb.VoronoiGeometry(xs, hvb.cuboid(3))                     # the reference's default: every axis periodic
    vd = hvb.VoronoiData(VG, getneighbors=True)
    assert len(vd.neighbors) == 1200 and all(len(nb) >= 4 for nb in vd.neighbors)
    assert max(int(nb.max()) for nb in vd.neighbors) <= 1200       # folded back to caller ids
