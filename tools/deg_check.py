import sys; sys.path.insert(0,'/root/repo')
import numpy as np, hvb200
g = np.stack(np.meshgrid(*[np.arange(6.0)] * 3, indexing="ij"), -1).reshape(-1, 3) / 6 + 1 / 12
for ts in (1,4,8):
    try:
        s = hvb200.Raycast(g, domain=hvb200.cuboid(3, periodic=[]), options=hvb200.RaycastParameter(tile_size=ts))
        m,_ = hvb200.voronoi(g, searcher=s)
        print(ts, "no error", m.sig.shape, s.stats())
    except hvb200.HVBError as e:
        print(ts, "error", e)
