"""The table of DESIGN.md section 8: the restated reference (oracle/) on seeded uniform clouds against the statistics the reference
publishes in docs/src/index.md:93,96 (tests/golden/ref_published/index_md_statistics.json).  CPU only.

    python tools/published_stats.py            small entries (N <= 2000), RCOriginal, with descents and nn-searches per walk
    python tools/published_stats.py big        d = 4, N = 30 000 and d = 5, N = 20 000 (minutes), 8 threads; the golden counts of tests/util.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import hv_oracle  # noqa: E402
import qhull_oracle  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "ref_published", "index_md_statistics.json")) as f:
    PUB = {int(k): v for k, v in json.load(f).items()}


def small():
    for d, nodes in ((4, (200, 500, 1000, 1500, 2000)), (5, (200, 500, 1000))):
        base, normal = qhull_oracle.cuboid(d)
        tot, tp = np.zeros(4), np.zeros(4)
        p = PUB[d]
        for n in nodes:
            c = p["nodes"].index(n)
            V, B, D, W, NN = [], [], [], [], []
            for seed in range(4):
                xs = np.random.default_rng(5000 + 100 * d + seed).random((n, d))
                r = hv_oracle.run(xs, base, normal, method="RCOriginal")
                st = r["stats"]
                V.append(len(r["sig"])); B.append(int((r["sig"] > n).any(axis=1).sum())); D.append(st["descents"])
                W.append(st["raycasts"] - d * st["descents"]); NN.append(st["nn_calls"])
            print("d=%d N=%5d  vertices %9.1f / %9.1f   boundary %8.1f / %8.1f   walks %9.1f / %9.1f   descents %5.2f / %5.2f   nn per walk %.4f / %.4f" % (
                d, n, np.mean(V), p["vertices"][c], np.mean(B), p["boundary_vertices"][c], np.mean(W), p["walks"][c], np.mean(D),
                p["vertices"][c] - p["walks"][c], sum(NN) / sum(W), p["nn_per_walk"][c]))
            tot += (np.mean(V), np.mean(B), np.mean(W), np.mean(D))
            tp += (p["vertices"][c], p["boundary_vertices"][c], p["walks"][c], p["vertices"][c] - p["walks"][c])
        print("   pooled (ours / published - 1): vertices %+.4f  boundary %+.4f  walks %+.4f  descents %+.3f" % tuple(tot / tp - 1))


def big():
    for d, n in ((4, 30000), (5, 20000)):
        base, normal = qhull_oracle.cuboid(d)
        c = PUB[d]["nodes"].index(n)
        V, B = [], []
        for k in range(4):
            xs = np.random.default_rng(7000 + 100 * d + k).random((n, d))
            r = hv_oracle.run(xs, base, normal, nthreads=min(8, os.cpu_count() or 1))
            V.append(len(r["sig"])); B.append(int((r["sig"] > n).any(axis=1).sum()))
            print("d=%d N=%d cloud %d: %d vertices, %d on the boundary" % (d, n, k, V[-1], B[-1]), flush=True)
        pv, pb = PUB[d]["vertices"][c], PUB[d]["boundary_vertices"][c]
        print("d=%d N=%d: vertices %.1f (published %.2f, %+.3f %%), boundary %.1f (published %.2f, %+.3f %%)" % (
            d, n, np.mean(V), pv, 100 * (np.mean(V) / pv - 1), np.mean(B), pb, 100 * (np.mean(B) / pb - 1)))


if __name__ == "__main__":
    big() if len(sys.argv) > 1 and sys.argv[1] == "big" else small()
