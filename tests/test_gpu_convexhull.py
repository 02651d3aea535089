"""ConvexHull(xs) (chull.jl:213-238) by the device facet walk (hvb_convex_hull, csrc/hvb_hull.cuh), against Qhull's hull and
against round 1's way (facets read off the unbounded edges of a complete search)."""
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


@pytest.mark.parametrize("via", ["walk", "search"])
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
    assert np.array_equal(a.sig, b.sig)
    assert np.abs(a.u - b.u).max() < 1e-9 and np.abs(a.r - b.r).max() < 1e-9
    assert a.stats["vertices"] == 0 and a.stats["rays"] == len(a)
    assert a.stats["raycasts"] < {3: 0.02, 4: 0.2, 5: 0.6}[d] * b.stats["raycasts"]
    if d == 3:
        check_against_qhull(a, xs)


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
