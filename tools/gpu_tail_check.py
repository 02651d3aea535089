"""Self-checking pass over the suites the closing check of round 2 (tools/gpu_last.sh) had no budget for: refinement,
non-general position, periodic contexts, the three hull algorithms, moments.  numpy + the C ABI only (no scipy, no pytest, no
oracle: a few seconds of box time), every check an invariant of the result itself.  Prints one line per check."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hvb200 as hvb  # noqa: E402


def points(n, d, seed):
    return np.random.default_rng(seed).random((n, d))


def grid(m, d):
    return (np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), -1).reshape(-1, d) + 0.5) / m


def run(xs, dom, periodic=False, **settings):
    s = hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(**settings), periodic=periodic)
    mesh, _ = hvb.voronoi(xs, searcher=s)
    return mesh, s


def check_refine():
    d, n0, m = 3, 2000, 150
    old, new = points(n0, d, 603), points(m, d, 613)
    dom = hvb.cuboid(d, periodic=[])
    s = hvb.Raycast(old, domain=dom)
    mesh0, _ = hvb.voronoi(old, searcher=s, copy=True)
    xs_all, sig, r, affected = hvb.refine(s, new, old, mesh0.sig, mesh0.r)
    s2 = hvb.Raycast(xs_all, domain=dom)
    ref, _ = hvb.voronoi(xs_all, searcher=s2, copy=True)
    assert np.array_equal(sig, ref.sig) and np.array_equal(r, ref.r)
    return "%d rows" % len(sig)


def check_lattice():
    out = []
    for d, m in ((2, 10), (3, 6), (4, 3)):
        xs = grid(m, d)
        mesh, s = run(xs, hvb.cuboid(d, periodic=[]))
        assert mesh.max_siglen == 2 ** d and mesh.number_of_vertices() == (m + 1) ** d
        vol = mesh.volumes()
        assert abs(vol.sum() - 1.0) < 1e-7 and np.abs(vol * m ** d - 1.0).max() < 1e-6
        out.append("d=%d: %d vertices" % (d, mesh.number_of_vertices()))
    return ", ".join(out)


def check_periodic():
    out = []
    for d, n in ((2, 3000), (3, 1500), (4, 500)):
        xs = points(n, d, 530 + d)
        mesh, s = run(xs, hvb.cuboid(d), periodic=True)
        vol = mesh.volumes()
        assert vol.shape == (n,) and (vol > 0).all() and abs(vol.sum() - 1.0) < 1e-11
        st = s.stats()
        if d == 2:
            assert st["unique_vertices"] == 2 * n                      # Euler on the 2-torus
        out.append("d=%d: %d unique vertices" % (d, st["unique_vertices"]))
    return ", ".join(out)


def check_hull():
    out = []
    for d, n in ((3, 20000), (4, 5000), (5, 1000)):
        xs = points(n, d, 710 + d)
        a, b, w = hvb.ConvexHull(xs), hvb.ConvexHull(xs, via="search"), hvb.ConvexHull(xs, via="walk")
        for c in (a, w):
            assert np.array_equal(c.sig, b.sig)
            assert np.abs(c.u - b.u).max() < 1e-9 and np.abs(c.r - b.r).max() < 1e-8
        assert ((xs[None, :64] - a.r[:, None, :])[:, :, :] * a.u[:, None, :]).sum(-1).max() < 1e-10
        out.append("d=%d: %d facets" % (d, len(a)))
    return ", ".join(out)


def check_moments():
    d, n = 4, 1000
    xs = points(n, d, 560)
    mesh, s = run(xs, hvb.cuboid(d, periodic=[]))
    vol = mesh.volumes()
    assert abs(vol.sum() - 1.0) < 1e-11
    m0, m1, m2 = mesh.moments()
    assert abs(m0.sum() - 1.0) < 1e-11 and np.abs(m1.sum(0) - 0.5).max() < 1e-11
    return "ok"


if __name__ == "__main__":
    bad = 0
    for f in (check_refine, check_lattice, check_periodic, check_hull, check_moments):
        t = time.time()
        try:
            print("%-20s ok   %s  (%.2f s)" % (f.__name__, f(), time.time() - t), flush=True)
        except Exception as e:                                         # noqa: BLE001
            bad += 1
            print("%-20s FAIL %s: %s  (%.2f s)" % (f.__name__, type(e).__name__, e, time.time() - t), flush=True)
    print("tail check: %s" % ("green" if bad == 0 else "%d FAILED" % bad), flush=True)
    sys.exit(1 if bad else 0)
