"""Small end-to-end exercise of every kernel family for `compute-sanitizer --tool memcheck python tools/sanitize_check.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import hvb200

rng = np.random.default_rng(0)
for d, n in ((3, 1500), (2, 2000), (5, 120)):
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
for p in (1, 2, 0):
    xs = rng.random((1200, 3))
    s = hvb200.Raycast(xs, domain=hvb200.cuboid(3, periodic=[]), options=hvb200.RaycastParameter(persistent=p))
    mesh, _ = hvb200.voronoi(xs, searcher=s, copy=True)
    s.close()
print("sanitize_check ok")
