"""ConvexHull(xs) (chull.jl:213-238) by gift wrapping on the device (hvb_convex_hull, csrc/hvb_wrap.cuh), against Qhull's hull,
against the walk around the unbounded 2-faces (hvb_convex_hull_via(ctx, 1), csrc/hvb_hull.cuh) and against round 1's way (facets
read off the unbounded edges of a complete search)."""
import numpy as np
import pytest

from util import points

pytestmark = pytest.mark.gpu


def check_against_qhull(cv, xs):
    from scipy.spatial import ConvexHull as QHull
    q = QHull(xs)
    want = {tuple(sorted(int(v) + 1 for v in f)): eq for f, eq in zip(q.simplices, q.equations)}
    got = {tuple(int(v) for v in sig): (r, u) for sig, r, u in cv}
    assert len(cv) == len(q.simplices) and set(got) == set(want)
    for sig, (r, u) in got.items():
        eq = want[sig]
        assert abs(np.linalg.norm(u) - 1.0) < 1e-12 and u @ eq[:-1] > 1.0 - 1e-9      # outer unit normal
        on_plane = (xs[np.array(sig) - 1] - r) @ u
        assert np.abs(on_plane).max() < 1e-10                                          # r lies in the facet's plane
        assert ((xs - r) @ u).max() < 1e-10                                            # every node behind the plane


@pytest.mark.parametrize("via", ["wrap", "walk", "search"])
@pytest.mark.parametrize("d,n", [(2, 3000), (3, 2000), (4, 600), (5, 200), (6, 100)])
def test_convex_hull_matches_qhull(hvb, d, n, via):
    xs = points(n, d, 700 + d)
    check_against_qhull(hvb.ConvexHull(xs, via=via), xs)


@pytest.mark.parametrize("d,n", [(3, 100000), (4, 20000), (5, 3000)])
def test_facet_walk_never_visits_the_interior(hvb, d, n):
    """the point of the walk: the hull without the tessellation.  Same facets as the complete search, a small
    fraction of its raycasts, no vertex returned."""
    xs = points(n, d, 710 + d)
    a = hvb.ConvexHull(xs)
    b = hvb.ConvexHull(xs, via="search")
    w = hvb.ConvexHull(xs, via="walk")
    for c in (a, w):
        assert np.array_equal(c.sig, b.sig)
        assert np.abs(c.u - b.u).max() < 1e-9 and np.abs(c.r - b.r).max() < 1e-8
        assert c.stats["vertices"] == 0 and c.stats["rays"] == len(c)
    assert w.stats["raycasts"] < {3: 0.02, 4: 0.2, 5: 0.6}[d] * b.stats["raycasts"]
    # one query per ridge at most: d / 2 per facet (+ the d - 1 seed steps)
    assert a.stats["raycasts"] <= len(a) * d / 2 + d
    if d == 3:
        check_against_qhull(a, xs)


def test_wrapping_is_reproducible_and_filter_independent(hvb):
    """normals and centres come from the facet's generators alone: bitwise equal from run to run and with the FP32 filter off"""
    xs = points(4000, 4, 77)
    a = hvb.ConvexHull(xs)
    b = hvb.ConvexHull(xs)
    c = hvb.ConvexHull(xs, options=hvb.RaycastParameter(fp32_filter=0))
    for o in (b, c):
        assert np.array_equal(a.sig, o.sig) and np.array_equal(a.u, o.u) and np.array_equal(a.r, o.r)
    assert c.stats["candidates_fp64"] > 100 * a.stats["candidates_fp64"]


@pytest.mark.parametrize("d,n", [(5, 50000), (2, 1000000), (6, 3000)])
def test_wrapping_at_config_size(hvb, d, n):
    """the hull of BASELINE's C4 / C3 clouds (and a d = 6 cloud): every generator behind every facet, every ridge shared by
    exactly two facets (a closed surface), Euler-Poincare for d = 2, 3 is implied by the comparison with Qhull above"""
    xs = points(n, d, 720 + d)
    cv = hvb.ConvexHull(xs)
    assert len(cv) > 0 and np.abs(np.linalg.norm(cv.u, axis=1) - 1.0).max() < 1e-12
    off = (cv.u * cv.r).sum(axis=1)
    step = max(1, len(cv) // 2000)
    for i in range(0, len(cv), step):
        assert (xs @ cv.u[i] - off[i]).max() < 1e-10
    ridges = {}
    for f in cv.sig:
        for j in range(d):
            key = tuple(np.delete(f, j))
            ridges[key] = ridges.get(key, 0) + 1
    assert set(ridges.values()) == {2}


def test_hull_needs_the_unbounded_domain(hvb):
    xs = points(500, 3, 1)
    s = hvb.Raycast(xs, domain=hvb.cuboid(3, periodic=[]))
    L = hvb._abi.lib()
    assert L.hvb_convex_hull(s._ctx) == hvb._abi.HVB_EINVAL


def test_hull_of_cospherical_points_is_reported(hvb):
    g = np.stack(np.meshgrid(*[np.arange(4.0)] * 3, indexing="ij"), -1).reshape(-1, 3)          # cube grid: square facets
    with pytest.raises(hvb.HVBError) as e:
        hvb.ConvexHull(g)
    assert e.value.code in (hvb._abi.HVB_EDEGENERATE, hvb._abi.HVB_EINCOMPLETE)
