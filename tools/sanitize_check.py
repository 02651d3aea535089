"""Small end-to-end exercise of every kernel family for `compute-sanitizer --tool memcheck python tools/sanitize_check.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import hvb200

rng = np.random.default_rng(0)
ROUND1 = "r2" not in sys.argv                      # `python tools/sanitize_check.py r2`: only the kernels added in round 2
for d, n in ((3, 1500), (2, 2000), (5, 120)) if ROUND1 else ():
    xs = rng.random((n, d))
    s = hvb200.Raycast(xs, domain=hvb200.cuboid(d, periodic=[]), options=hvb200.RaycastParameter(neighbors=1))
    mesh, _ = hvb200.voronoi(xs, searcher=s, copy=True)
    vol, area = mesh.volumes(), mesh.areas()
    assert abs(vol.sum() - 1) < 1e-10
    if d == 3:
        new = rng.random((60, d))
        xs_all, sig, r, aff = hvb200.refine(s, new, xs, mesh.sig, mesh.r)
        assert len(sig) > len(mesh.sig)
    s.close()
if ROUND1:
    xs = rng.random((800, 3))
    s = hvb200.Raycast(xs, domain=hvb200.Boundary())
    mesh, _ = hvb200.voronoi(xs, searcher=s, copy=True)
    assert len(mesh.ray_edge) > 0 and np.isinf(mesh.volumes()).any()
    s.close()
    xs = rng.random((1500, 2))
    s = hvb200.Raycast(xs, domain=hvb200.cuboid(2), periodic=True)
    mesh, _ = hvb200.voronoi(xs, searcher=s)
    assert abs(mesh.volumes().sum() - 1) < 1e-10
    s.close()
for p in (1, 2, 0) if ROUND1 else ():
    xs = rng.random((1200, 3))
    s = hvb200.Raycast(xs, domain=hvb200.cuboid(3, periodic=[]), options=hvb200.RaycastParameter(persistent=p))
    mesh, _ = hvb200.voronoi(xs, searcher=s, copy=True)
    s.close()
# round 2: convex hull (gift wrapping: TMA ring, partial merge, commit; and the older walk), non-general position (perturb, merge,
# volumes from the records, moments), d = 2 bucket sort of the rows (the d = 2 searches above), small odd sizes for the bulk copies
xs = rng.random((1500, 2))                             # d = 2: rows ordered by the bucket sort
s = hvb200.Raycast(xs, domain=hvb200.cuboid(2, periodic=[]))
mesh, _ = hvb200.voronoi(xs, searcher=s, copy=True)
s.close()
for d, n in ((2, 1001), (3, 777), (4, 300), (5, 150), (6, 40)):
    xs = rng.random((n, d))
    a = hvb200.ConvexHull(xs)
    b = hvb200.ConvexHull(xs, via="walk")
    assert np.array_equal(a.sig, b.sig)
for d, m in ((2, 7), (3, 5)):
    g = (np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), -1).reshape(-1, d) + 0.5) / m
    s = hvb200.Raycast(g, domain=hvb200.cuboid(d, periodic=[]))
    mesh, _ = hvb200.voronoi(g, searcher=s)
    assert mesh.max_siglen == 2 ** d and abs(mesh.volumes().sum() - 1) < 1e-7
    vol, first, second = mesh.moments()
    off, ids = mesh.neighbors()
    s.close()
xs = rng.random((900, 3))
s = hvb200.Raycast(xs, domain=hvb200.cuboid(3, periodic=[]))
mesh, _ = hvb200.voronoi(xs, searcher=s)
vol, first, second = mesh.moments()
assert abs(vol.sum() - 1) < 1e-10 and np.abs(first.sum(0) - 0.5).max() < 1e-10
s.close()
print("sanitize_check ok")
