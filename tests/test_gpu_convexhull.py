"""ConvexHull(xs) (chull.jl:213-238) from the unbounded edges of the search, against Qhull's hull."""
import numpy as np
import pytest

from util import points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,n", [(2, 3000), (3, 2000), (4, 600), (5, 200)])
def test_convex_hull_matches_qhull(hvb, d, n):
    from scipy.spatial import ConvexHull as QHull
    xs = points(n, d, 700 + d)
    cv = hvb.ConvexHull(xs)
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
