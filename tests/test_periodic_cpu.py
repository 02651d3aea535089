"""Periodic domains, CPU side: the certificate the product relies on, proven on the two oracles.

If every vertex of the explicit-halo problem that touches a caller generator has its ball inside the pushed planes, the
vertices touching caller generators are exactly the vertices of the periodic tessellation.  Here the halo problem is
solved by the CPU restatement of the reference algorithm and compared with Qhull on the 3^k replication."""
import numpy as np
import pytest

import periodic_oracle as po
from util import points


def certified_rows(oracle, xs, axes, margin):
    n, d = xs.shape
    origin, mult, hxs = po.halo(xs, axes, margin)
    ext = np.vstack([xs, hxs])
    base, normal = po.pushed_cuboid(d, axes, margin)
    o = oracle.run(ext, base, normal)
    touch = (o["sig"] <= n).any(axis=1)
    sig, r = o["sig"][touch], o["r"][touch]
    R = np.linalg.norm(r - ext[sig[:, 0] - 1], axis=1)
    excess = 0.0
    for a in axes:
        excess = max(excess, float((r[:, a - 1] + R - 1.0).max()), float((-(r[:, a - 1] - R)).max()))
    on_pushed = 0
    for a in axes:
        on_pushed += int(((sig == ext.shape[0] + 2 * a - 1) | (sig == ext.shape[0] + 2 * a)).any(axis=1).sum())
    return sig, r, origin, mult, excess, on_pushed


@pytest.mark.parametrize("d,n,axes,margin", [(2, 300, (1, 2), 0.35), (2, 300, (2,), 0.35), (3, 150, (1, 2, 3), 0.75), (3, 200, (1, 3), 0.6)])
def test_certified_halo_problem_equals_torus(oracle, d, n, axes, margin):
    xs = points(n, d, 42 + d)
    sig, r, origin, mult, excess, on_pushed = certified_rows(oracle, xs, axes, margin)
    assert on_pushed == 0 and excess <= margin, (on_pushed, excess)          # the certificate holds for this margin
    got = po.fold_rows(sig, n, origin, mult)
    if len(axes) == d:
        want = po.torus_simplices(xs, axes)
        assert got == want
        if d == 2:
            assert po.canonical_classes(got) == 2 * n                         # a triangulated torus has 2n triangles
    else:
        # mixed domain: compare the plane-free vertices only (the torus oracle has no Dirichlet faces)
        # (the torus oracle has no Dirichlet faces): they are the Delaunay simplices whose circumcentre lies in the domain
        want = po.torus_simplices(xs, axes, centres=True)
        assert got <= set(want)
        inside = {s for s, c in want.items() if (c > 0).all() and (c < 1).all()}
        assert inside <= got


def test_small_margin_fails_the_certificate(oracle):
    xs = points(300, 2, 44)
    *_, excess, on_pushed = certified_rows(oracle, xs, (1, 2), 0.02)
    assert on_pushed > 0 or excess > 0.02
