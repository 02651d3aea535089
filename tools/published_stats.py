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


def counts():
    """every entry of both matrices up to 30 000 (d = 4) / 20 000 (d = 5) nodes on 4 seeded clouds points(n, d, 8000 + 1000 d + 10 c + k),
    c = column of the matrix, k = 0..3: the restatement's vertex / boundary-vertex counts -> tests/golden/ref_published/seeded_counts.json
    (the counts the device path has to return for the same clouds, tests/test_gpu_volumes.py)"""
    out = {}
    for d, maxn in ((4, 30000), (5, 20000)):
        base, normal = qhull_oracle.cuboid(d)
        p = PUB[d]
        for c, n in enumerate(p["nodes"]):
            if n > maxn:
                continue
            V, B = [], []
            for k in range(4):
                xs = np.random.default_rng(8000 + 1000 * d + 10 * c + k).random((n, d))
                r = hv_oracle.run(xs, base, normal, nthreads=min(8, os.cpu_count() or 1))
                V.append(len(r["sig"])); B.append(int((r["sig"] > n).any(axis=1).sum()))
            out["%d,%d" % (d, n)] = {"column": c, "vertices": V, "boundary_vertices": B}
            print("d=%d N=%5d  vertices %+.4f  boundary %+.4f  (ours / published - 1; times sqrt(N): %+.2f, %+.2f)" % (
                d, n, np.mean(V) / p["vertices"][c] - 1, np.mean(B) / p["boundary_vertices"][c] - 1,
                np.sqrt(n) * (np.mean(V) / p["vertices"][c] - 1), np.sqrt(n) * (np.mean(B) / p["boundary_vertices"][c] - 1)), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_published", "seeded_counts.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    {"big": big, "counts": counts}.get(sys.argv[1] if len(sys.argv) > 1 else "", small)()
