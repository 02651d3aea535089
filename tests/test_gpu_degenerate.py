"""Non-general position: clouds with more than d + 1 cospherical generators.  The reference returns ONE vertex per
cospherical set whose signature lists all its generators (raycast.jl:870-969) and enumerates its edges with FastEdgeIterator
(edgeiterate.jl:82-780).  The backend gets the same result by perturbation + merge (csrc/hvb_ctx.cuh, resolve_degenerate); the
truth here is Qhull's Voronoi diagram of the same cloud, which merges cospherical facets (oracle/qhull_oracle.py,
voronoi_nongeneral) -- independent of the search."""
import numpy as np
import pytest

import qhull_oracle
from util import points

pytestmark = pytest.mark.gpu


def grid(m, d):
    return (np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), -1).reshape(-1, d) + 0.5) / m


def run(hvb, xs, bounded=True, **settings):
    d = xs.shape[1]
    dom = hvb.cuboid(d, periodic=[]) if bounded else hvb.Boundary()
    s = hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(**settings))
    mesh, _ = hvb.voronoi(xs, searcher=s)
    return mesh, s


def check_against_qhull(mesh, xs, bounded=True):
    d = xs.shape[1]
    base, normal = qhull_oracle.cuboid(d) if bounded else (None, None)
    want = qhull_oracle.voronoi_nongeneral(xs, base, normal)
    got = {frozenset(int(i) for i in sg): r for sg, r in zip(mesh.sigs(), mesh.r)}
    assert len(got) == mesh.number_of_vertices()
    assert set(got) == set(want)
    assert max(np.abs(got[k] - want[k]).max() for k in want) < 1e-10
    return want


@pytest.mark.parametrize("d,m", [(2, 12), (3, 6), (4, 4), (5, 2), (6, 2)])
def test_cubic_grid_matches_qhull(hvb, d, m):
    """test/fraud.jl / test/periodicgrids.jl style input: every interior vertex has 2^d generators"""
    xs = grid(m, d)
    mesh, s = run(hvb, xs)
    assert mesh.max_siglen == 2 ** d and mesh.sig is None
    want = check_against_qhull(mesh, xs)
    assert len(want) == (m + 1) ** d                        # every corner of the lattice of cubes
    st = s.stats()
    assert st["vertices"] == len(want) and st["degenerate"] == sum(1 for k in want if len(k) > d + 1)
    # rows are sorted lexicographically
    rows = [tuple(int(i) for i in sg) for sg in mesh.sigs()]
    assert rows == sorted(rows)


@pytest.mark.parametrize("d,m,nrand", [(2, 10, 300), (3, 5, 200), (4, 3, 100)])
def test_grid_inside_a_random_cloud(hvb, d, m, nrand):
    """a lattice patch (non-general) surrounded by random generators (general): both kinds of vertices in one mesh"""
    g = 0.3 + 0.4 * grid(m, d)
    rnd = points(nrand, d, 50 + d)
    keep = np.any((rnd < 0.28) | (rnd > 0.72), axis=1)
    xs = np.vstack([g, rnd[keep]])
    mesh, s = run(hvb, xs)
    want = check_against_qhull(mesh, xs)
    assert max(len(k) for k in want) == 2 ** d and min(len(k) for k in want) == d + 1


def test_unbounded_grid_is_reported(hvb):
    """without boundary planes the faces of a lattice are coplanar hull facets: not resolved, reported as before"""
    xs = grid(7, 3)
    with pytest.raises(hvb.HVBError) as e:
        run(hvb, xs, bounded=False)
    assert e.value.code == hvb._abi.HVB_EDEGENERATE


def test_neighbours_of_a_grid_are_the_face_neighbours(hvb):
    """neighbors_of_cell (neighbors.jl:205-212): a neighbour shares a FULL interface; the 2^d - 1 - d cells that meet a cube
    cell only at an edge or a corner are not neighbours although they share vertices"""
    m, d = 5, 3
    xs = grid(m, d)
    mesh, s = run(hvb, xs)
    off, ids = mesh.neighbors()
    n = len(xs)
    cell = lambda i, j, k: (i * m + j) * m + k + 1
    for i in range(m):
        for j in range(m):
            for k in range(m):
                want = set()
                for axis, c in enumerate((i, j, k)):
                    for step in (-1, 1):
                        q = [i, j, k]; q[axis] += step
                        if 0 <= q[axis] < m:
                            want.add(cell(*q))
                        else:
                            want.add(n + 2 * axis + (1 if step < 0 else 2))      # plane ids: checked as a count below
                got = set(int(v) for v in ids[off[cell(i, j, k) - 1]:off[cell(i, j, k)]])
                assert {v for v in got if v <= n} == {v for v in want if v <= n}
                assert len({v for v in got if v > n}) == len({v for v in want if v > n})


def test_cospherical_generators_become_one_vertex(hvb):
    """five generators on a common sphere (the case test_near_degenerate_input_is_reported_like_the_reference constructs):
    one vertex with five generators, also when the fifth is off the sphere by 1e-14"""
    rng = np.random.default_rng(5)
    u = rng.normal(size=(5, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    far = 0.5 + 0.49 * np.sign(rng.normal(size=(40, 3))) * (0.8 + 0.2 * rng.random((40, 3)))
    for eps in (0.0, 1e-14):
        cos = 0.5 + 0.25 * u
        cos[4] = 0.5 + 0.25 * (1.0 + eps) * u[4]
        xs = np.vstack([cos, far])
        mesh, s = run(hvb, xs)
        big = [sg for sg in mesh.sigs() if len(sg) > 4]
        assert len(big) == 1 and list(big[0]) == [1, 2, 3, 4, 5]
        v = [i for i, sg in enumerate(mesh.sigs()) if len(sg) > 4][0]
        assert np.abs(mesh.r[v] - 0.5).max() < 1e-10


def test_reporting_is_still_available_and_fixed_width_calls_refuse(hvb):
    xs = grid(5, 3)
    with pytest.raises(hvb.HVBError) as e:
        run(hvb, xs, on_degenerate=0)
    assert e.value.code == hvb._abi.HVB_EDEGENERATE
    mesh, s = run(hvb, xs)
    L, ctx = hvb._abi.lib(), s._ctx
    sig = np.empty((mesh.number_of_vertices(), 4), dtype=np.int64)
    assert L.hvb_fetch_vertices(ctx, sig.ctypes.data_as(hvb._abi.ctypes.c_void_p), None) == hvb._abi.HVB_ESTATE
    area = np.empty(10 * len(xs))
    assert L.hvb_cell_areas(ctx, area.ctypes.data_as(hvb._abi.ctypes.c_void_p)) == hvb._abi.HVB_ESTATE
    # a second search on the same context goes straight to the resolved form; new points reset it
    mesh2, _ = hvb.voronoi(xs, searcher=s)
    assert [tuple(a) for a in mesh2.sigs()] == [tuple(a) for a in mesh.sigs()]
    # an Iter subset would lose generators of a cospherical set: reported, not resolved
    with pytest.raises(hvb.HVBError) as e:
        hvb.voronoi(xs, searcher=s, Iter=range(1, 11))
    assert e.value.code == hvb._abi.HVB_EDEGENERATE


def test_general_position_is_untouched(hvb):
    """on_degenerate = 2 (the default) costs nothing and changes nothing for a cloud in general position"""
    xs = points(3000, 3, 8)
    a, sa = run(hvb, xs, on_degenerate=0)
    b, sb = run(hvb, xs, on_degenerate=2)
    assert b.max_siglen == 4 and np.array_equal(a.sig, b.sig) and np.array_equal(a.r, b.r)
    assert sb.stats()["degenerate"] == 0 and sb.stats()["kernel_launches"] == sa.stats()["kernel_launches"]


def test_large_grid(hvb):
    """30^3 lattice: 31^3 vertices, 29^3 of them with eight generators"""
    m = 30
    xs = grid(m, 3)
    mesh, s = run(hvb, xs)
    lens = np.diff(mesh.sig_off)
    assert mesh.number_of_vertices() == (m + 1) ** 3 and int((lens == 8).sum()) == (m - 1) ** 3
    # every vertex is a corner of the lattice of cubes
    assert np.abs(mesh.r * m - np.round(mesh.r * m)).max() < 1e-9


@pytest.mark.parametrize("d,m", [(2, 10), (3, 6), (4, 3)])
def test_volumes_of_a_grid(hvb, d, m):
    """the reference's known-answer test of this path (test/periodicgrids.jl, test/rcmethods.jl:8): the cell volumes add up to the
    domain -- on a lattice every cell is a cube of volume m^-d.  Computed on the perturbed diagram: exact to ~1e-9"""
    xs = grid(m, d)
    mesh, s = run(hvb, xs)
    vol = mesh.volumes()
    assert abs(vol.sum() - 1.0) < 1e-7
    assert np.abs(vol * m ** d - 1.0).max() < 1e-6


@pytest.mark.parametrize("d,m", [(2, 6), (3, 4)])
def test_periodic_lattice_through_the_host_side_halo(hvb, d, m):
    """test/periodicgrids.jl style input the way the reference itself periodises (Create_Discrete_Domain, domain.jl:175-213: halo
    generators on the host, then voronoi() on generators + halo with a plain Boundary -- the seam the Julia shim replaces): a lattice
    on the unit torus, its periodic copies within a margin, mirror planes pushed out by that margin.  Every vertex in [0, 1)^d is a
    corner of the lattice of cubes with 2^d generators, one per class of the torus: m^d of them."""
    import itertools
    g = grid(m, d)
    margin = 1.5 / m
    pts = [g]
    for shift in itertools.product((-1, 0, 1), repeat=d):
        if any(shift):
            c = g + np.array(shift, dtype=float)
            pts.append(c[np.all((c > -margin) & (c < 1 + margin), axis=1)])
    xs = np.vstack(pts)
    dom = hvb.cuboid(d, dimensions=np.full(d, 1 + 2 * margin), periodic=[], offset=np.full(d, -margin))
    s = hvb.Raycast(xs, domain=dom)
    mesh, _ = hvb.voronoi(xs, searcher=s)
    inside = np.all((mesh.r > -1e-9) & (mesh.r < 1 - 1e-9), axis=1)
    lens = np.diff(mesh.sig_off)
    assert int(inside.sum()) == m ** d and np.all(lens[inside] == 2 ** d)
    assert np.abs(mesh.r[inside] * m - np.round(mesh.r[inside] * m)).max() < 1e-9
    # folded to the torus every such vertex names 2^d DIFFERENT caller generators (m >= 3)
    origin = np.concatenate([np.arange(len(g))] + [np.flatnonzero(np.all((g + np.array(sh, dtype=float) > -margin) & (g + np.array(sh, dtype=float) < 1 + margin), axis=1))
                                                   for sh in itertools.product((-1, 0, 1), repeat=d) if any(sh)])
    for v in np.flatnonzero(inside):
        ids = mesh.sig_ids[mesh.sig_off[v]:mesh.sig_off[v + 1]]
        assert ids.max() <= len(xs) and len(set(origin[ids - 1].tolist())) == 2 ** d
