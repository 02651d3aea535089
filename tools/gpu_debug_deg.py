import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim")):
    sys.path.insert(0, p)
import numpy as np
import hvb200, hostsim, qhull_oracle
M64 = (1 << 64) - 1
def mix64(x):
    x ^= x >> 33; x = (x * 0xff51afd7ed558ccd) & M64; x ^= x >> 33; x = (x * 0xc4ceb9fe1a85ec53) & M64; x ^= x >> 33
    return x
def perturb(g, rel=1e-9):
    ext = (g.max(0) - g.min(0)).max()
    flat = g.ravel().copy()
    for i in range(len(flat)):
        h = mix64((i * 0x9e3779b97f4a7c15 + 0x243f6a8885a308d3) & M64)
        flat[i] += rel * ext * (((h >> 11) + 0.5) / 9007199254740992.0 * 2 - 1)
    return flat.reshape(g.shape)
d, m = 3, 6
g = (np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), -1).reshape(-1, d) + 0.5) / m
xp = perturb(g)
base, normal = qhull_oracle.cuboid(d)
h = hostsim.run(xp, base, normal)
s = hvb200.Raycast(xp, domain=hvb200.cuboid(d, periodic=[]), options=hvb200.RaycastParameter(on_degenerate=1))
mesh, _ = hvb200.voronoi(xp, searcher=s)
hs = set(map(tuple, h["sig"].tolist())); ds = set(map(tuple, np.asarray(mesh.sig).tolist()))
print("pre-perturbed input: hostsim rows", len(hs), "device rows", len(ds), "equal", hs == ds, "only device", sorted(ds - hs)[:6], "only host", sorted(hs - ds)[:6])
s2 = hvb200.Raycast(g, domain=hvb200.cuboid(d, periodic=[]))
mesh2, _ = hvb200.voronoi(g, searcher=s2)
sg = [tuple(int(i) for i in a) for a in mesh2.sigs()]
print("resolved: vertices", len(sg), "stats", {k: s2.stats()[k] for k in ("vertices", "degenerate", "rejected", "raycasts")})
print([a for a in sg if 219 in a and len(a) == 5][:10])
want = qhull_oracle.voronoi_nongeneral(g, base, normal)
ws = set(tuple(sorted(k)) for k in want)
print("extra", sorted(set(sg) - ws)[:10], "missing", sorted(ws - set(sg))[:10])
