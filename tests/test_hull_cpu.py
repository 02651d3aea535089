"""The two hull walks of the device compiled as host C++ (tests/hostsim): gift wrapping (csrc/hvb_wrap.cuh, behind
hvb_convex_hull) and the walk around the unbounded 2-faces (csrc/hvb_hull.cuh, hvb_convex_hull_via(ctx, 1)): identical facets
and normals to Qhull.  Runs without a GPU; the same code runs on the device (tests/test_gpu_convexhull.py)."""
import numpy as np
import pytest

import hostsim
from util import points


@pytest.mark.parametrize("d,n", [(2, 500), (3, 2000), (4, 1500), (5, 600), (6, 150)])
def test_facet_walk_matches_qhull_on_the_host(d, n):
    from scipy.spatial import ConvexHull as QHull
    xs = points(n, d, 900 + d)
    F, N, st = hostsim.hull(xs)
    q = QHull(xs)
    want = {tuple(sorted(int(v) + 1 for v in f)): eq[:d] for f, eq in zip(q.simplices, q.equations)}
    got = [tuple(f) for f in F.tolist()]
    assert len(got) == len(set(got)) == len(want) and set(got) == set(want)
    assert max(np.abs(N[i] - want[got[i]]).max() for i in range(len(got))) < 1e-12
    assert st["degenerate"] == 0
    # a walk around a ridge is a handful of edge walks: far fewer raycasts than a tessellation has vertices
    assert st["raycasts"] < 12 * len(got) * d


@pytest.mark.parametrize("d,n,slots", [(2, 5000, 1), (3, 20000, 7), (4, 3000, 3), (5, 1200, 2), (6, 200, 5)])
def test_gift_wrapping_matches_qhull_on_the_host(d, n, slots):
    from scipy.spatial import ConvexHull as QHull
    xs = points(n, d, 930 + d)
    F, N, C, st = hostsim.wrap(xs, slots=slots)
    q = QHull(xs)
    want = {tuple(sorted(int(v) + 1 for v in f)): eq[:d] for f, eq in zip(q.simplices, q.equations)}
    got = [tuple(f) for f in F.tolist()]
    assert len(got) == len(set(got)) == len(want) and set(got) == set(want)
    assert max(np.abs(N[i] - want[got[i]]).max() for i in range(len(got))) < 1e-12
    assert st["degenerate"] == 0
    # the reported point is the circumcentre of the facet's generators inside the facet's hyperplane (chull.jl:224-232)
    for i, f in enumerate(got):
        P = xs[np.array(f) - 1]
        rad = np.linalg.norm(P - C[i], axis=1)
        assert rad.max() - rad.min() < 1e-9 * (1 + rad.max()) and abs((P[0] - C[i]) @ N[i]) < 1e-10
    # at most one query per ridge (d / 2 per facet) plus the d - 1 seed steps of the 2 d first facets
    assert st["raycasts"] <= len(got) * d / 2 + 2 * d * (d - 1)
    # the FP32 filter leaves about one FP64 evaluation per query and piece of the stream
    assert st["fp64"] <= 4 * slots * st["raycasts"]


def test_gift_wrapping_filter_is_sound():
    """with the FP32 filter off every generator is evaluated in FP64: same facets, normals and centres bitwise"""
    xs = points(1500, 4, 941)
    a = hostsim.wrap(xs, slots=2)
    b = hostsim.wrap(xs, slots=5, fp32=0)
    assert np.array_equal(a[0], b[0]) or sorted(map(tuple, a[0].tolist())) == sorted(map(tuple, b[0].tolist()))
    ka = {tuple(f): (tuple(nv), tuple(c)) for f, nv, c in zip(a[0].tolist(), a[1].tolist(), a[2].tolist())}
    kb = {tuple(f): (tuple(nv), tuple(c)) for f, nv, c in zip(b[0].tolist(), b[1].tolist(), b[2].tolist())}
    assert ka == kb
    assert b[3]["fp64"] > 100 * a[3]["fp64"]


def test_gift_wrapping_reports_coplanar_generators():
    g = np.stack(np.meshgrid(*[np.arange(4.0)] * 3, indexing="ij"), -1).reshape(-1, 3)          # cube grid: square facets
    assert hostsim.wrap(g)[3]["degenerate"] > 0


@pytest.mark.parametrize("d", [2, 3, 4, 5, 6])
def test_gift_wrapping_of_tiny_clouds_and_spheres(d):
    """edge cases: a simplex (d + 1 generators: every seed step and no wrap), d + 2 generators, and a cloud whose generators are
    all hull vertices (a sphere)"""
    from scipy.spatial import ConvexHull as QHull
    rng = np.random.default_rng(100 + d)
    clouds = [rng.random((d + 1, d)), rng.random((d + 2, d))]
    if d <= 4:
        u = rng.normal(size=(200, d))
        clouds.append(u / np.linalg.norm(u, axis=1)[:, None])
    for xs in clouds:
        F, N, C, st = hostsim.wrap(xs, slots=2)
        want = {tuple(sorted(int(v) + 1 for v in f)) for f in QHull(xs).simplices}
        assert {tuple(f) for f in F.tolist()} == want and st["degenerate"] == 0
