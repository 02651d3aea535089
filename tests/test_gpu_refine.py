"""Refinement on the device: hvb_clean_affected + hvb_search over the new cells == a fresh tessellation of
all generators.  Mirrors systematic_refine! (meshrefine.jl:183-216): new nodes are prepended, old ids shift."""
import ctypes

import numpy as np
import pytest

from util import points

pytestmark = pytest.mark.gpu


def fresh(hvb, xs, dom):
    s = hvb.Raycast(xs, domain=dom)
    mesh, _ = hvb.voronoi(xs, searcher=s, copy=True)
    s.close()
    return mesh


@pytest.mark.parametrize("d,n0,m", [(2, 3000, 200), (3, 2000, 150), (4, 600, 40), (5, 150, 10)])
def test_refine_equals_fresh_tessellation(hvb, d, n0, m):
    old, new = points(n0, d, 600 + d), points(m, d, 610 + d)
    dom = hvb.cuboid(d, periodic=[])
    s = hvb.Raycast(old, domain=dom)
    mesh0, _ = hvb.voronoi(old, searcher=s, copy=True)
    xs_all, sig, r, affected = hvb.refine(s, new, old, mesh0.sig, mesh0.r)
    ref = fresh(hvb, xs_all, dom)
    assert np.array_equal(sig, ref.sig)
    assert np.array_equal(r, ref.r)                    # bitwise: canonical coordinates depend on the signature alone
    assert set(range(1, m + 1)) <= set(affected.tolist())
    # a second batch on top of the refined mesh
    new2 = points(m, d, 620 + d)
    xs2, sig2, r2, _ = hvb.refine(s, new2, xs_all, sig, r)
    ref2 = fresh(hvb, xs2, dom)
    assert np.array_equal(sig2, ref2.sig) and np.array_equal(r2, ref2.r)


def test_clean_affected_rule_and_affected_cells(hvb):
    """keep[v] <=> |x_sig1 - r| <= (1 + 1e-7) * dist(r, nearest new node)  (meshrefine.jl:133-138)"""
    from scipy.spatial import cKDTree
    d, n0, m = 3, 4000, 300
    old, new = points(n0, d, 630), points(m, d, 631)
    dom = hvb.cuboid(d, periodic=[])
    s = hvb.Raycast(old, domain=dom)
    mesh0, _ = hvb.voronoi(old, searcher=s, copy=True)
    xs_all = np.ascontiguousarray(np.vstack([new, old]))
    s.set_points(xs_all)
    sig = np.ascontiguousarray(mesh0.sig + m)
    r = np.ascontiguousarray(mesh0.r)
    keep = np.zeros(len(sig), dtype=np.uint8)
    aff = np.zeros(n0 + m, dtype=np.uint8)
    L = hvb._abi.lib()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    hvb._abi.check(L.hvb_clean_affected(s._ctx, P(sig), P(r), len(sig), d + 1, 1, m, P(keep), P(aff)), s._ctx)
    R = np.linalg.norm(xs_all[sig[:, 0] - 1] - r, axis=1)
    dist, _ = cKDTree(new).query(r)
    want = R <= (1.0 + 1e-7) * dist
    assert np.array_equal(keep.astype(bool), want)
    assert 0 < (~want).sum() < len(want)
    exp_aff = np.zeros(n0 + m, dtype=bool)
    exp_aff[:m] = True
    gone = sig[~want]
    exp_aff[gone[gone <= n0 + m] - 1] = True
    assert np.array_equal(aff.astype(bool), exp_aff)
    # argument errors
    assert L.hvb_clean_affected(s._ctx, P(sig), P(r), len(sig), d + 1, 0, m, P(keep), P(aff)) == hvb._abi.HVB_EINVAL
    assert L.hvb_clean_affected(s._ctx, P(sig), P(r), len(sig), d + 1, 1, n0 + m + 1, P(keep), P(aff)) == hvb._abi.HVB_EINVAL
