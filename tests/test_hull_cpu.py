"""The facet walk of csrc/hvb_hull.cuh compiled as host C++ (tests/hostsim): identical facets and normals to Qhull.
Runs without a GPU; the same code runs on the device behind hvb_convex_hull (tests/test_gpu_convexhull.py)."""
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
