"""Quick diagnostic run on the GPU box: a few configurations against the oracle with statistics and timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import hvb200, hv_oracle, qhull_oracle

def one(d, n, bounded, check=True, **kw):
    xs = np.random.default_rng(0).random((n, d))
    dom = hvb200.cuboid(d, periodic=[]) if bounded else hvb200.Boundary()
    t0 = time.time()
    s = hvb200.Raycast(xs, domain=dom, options=hvb200.RaycastParameter(**kw))
    t1 = time.time()
    try:
        mesh, _ = hvb200.voronoi(xs, searcher=s)
    except hvb200.HVBError as e:
        print("ERROR", d, n, bounded, e); return
    t2 = time.time()
    st = s.stats()
    line = dict(d=d, n=n, bounded=bounded, V=int(mesh.sig.shape[0]), t_create=round(t1 - t0, 4), t_search=round(t2 - t1, 4),
                **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
    if check:
        if bounded:
            b, nn = qhull_oracle.cuboid(d); o = hv_oracle.run(xs, b, nn)
        else:
            o = hv_oracle.run(xs)
        same = mesh.sig.shape == o["sig"].shape and np.array_equal(mesh.sig, o["sig"])
        line["sig_equal"] = bool(same)
        if same:
            x0 = xs[o["sig"][:, 0] - 1]
            line["rel_err"] = float((np.linalg.norm(mesh.r - o["r"], axis=1) / np.linalg.norm(o["r"] - x0, axis=1)).max())
        else:
            so = {tuple(x) for x in o["sig"].tolist()}; sg = {tuple(x) for x in mesh.sig.tolist()}
            line["missing"] = len(so - sg); line["extra"] = len(sg - so)
    print(json.dumps(line), flush=True)

if __name__ == "__main__":
    one(3, 1000, True); one(3, 1000, False); one(2, 5000, True); one(4, 500, True); one(5, 300, True); one(6, 150, True)
    one(3, 100000, True, check=False); one(3, 100000, True, check=False)
    one(2, 1000000, True, check=False)
    one(5, 50000, True, check=False)
    one(4, 30000, True, check=False)
