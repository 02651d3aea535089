"""Fuzz of the device algorithm's logic on the CPU: the host build of hvb_core.cuh (tests/hostsim) against Qhull (the
independent truth, oracle/qhull_oracle.py) on clouds the uniform bench inputs never produce -- clusters, shells, stretched and
offset boxes, tiny clouds, widely varying density.  Test infrastructure; prints every mismatch with the seed that reproduces it.

    python tools/fuzz_hostsim.py [minutes] [first_seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", os.path.join("tests", "hostsim")):
    sys.path.insert(0, os.path.join(ROOT, p))
import hostsim          # noqa: E402
import qhull_oracle     # noqa: E402


def cloud(rng, kind, n, d):
    if kind == "uniform":
        return rng.random((n, d))
    if kind == "clusters":
        k = int(rng.integers(2, 6))
        c = rng.random((k, d)) * 0.6 + 0.2
        s = 10.0 ** rng.uniform(-3, -1, size=k)
        w = rng.integers(0, k, size=n)
        return np.clip(c[w] + rng.standard_normal((n, d)) * s[w, None], 1e-6, 1 - 1e-6)
    if kind == "shell":
        v = rng.standard_normal((n, d))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        rad = 0.35 * (1.0 + 0.05 * rng.standard_normal((n, 1)))
        x = 0.5 + v * rad
        m = max(d + 2, n // 10)
        x[:m] = rng.random((m, d))                     # some interior points: a pure shell is close to cospherical
        return np.clip(x, 1e-6, 1 - 1e-6)
    if kind == "density":
        x = rng.random((n, d))
        x[:, 0] = x[:, 0] ** 4                          # density varies by orders of magnitude along axis 0
        return x
    if kind == "slab":
        x = rng.random((n, d))
        x[:, -1] = 0.5 + (x[:, -1] - 0.5) * 1e-2        # almost flat cloud
        return x
    raise ValueError(kind)


KINDS = ("uniform", "clusters", "shell", "density", "slab")
SIZES = {2: (5, 3000), 3: (6, 1500), 4: (7, 500), 5: (8, 160)}


def one(seed):
    rng = np.random.default_rng(seed)
    d = int(rng.integers(2, 6))
    lo, hi = SIZES[d]
    n = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
    kind = KINDS[int(rng.integers(0, len(KINDS)))]
    bounded = bool(rng.integers(0, 2))
    xs = cloud(rng, kind, n, d)
    if bounded:
        # a generator within 1e-12 (relative) of a boundary plane is degenerate input for the reference's own predicate
        # (u.x > c (1 + 1e-12), raycast.jl:802-805: the mirror image is not "ahead"); keep the cloud strictly inside
        xs = 1e-6 + xs * (1.0 - 2e-6)
    if len(np.unique(xs, axis=0)) != n:
        return None
    knobs = dict(ppc=int(rng.integers(0, 9)), probe_scale=float(rng.choice([0.0, 1.05, 1.5, 3.0])), seed_stride=int(rng.choice([0, 1, 7, 64])))
    tag = "seed=%d d=%d n=%d %s %s %s" % (seed, d, n, kind, "bounded" if bounded else "unbounded", knobs)
    if bounded:
        base, normal = qhull_oracle.cuboid(d)
        try:
            truth = qhull_oracle.bounded(xs, base, normal)
        except Exception as e:                              # Qhull refuses (nearly) degenerate input: not a finding
            return ("skip", tag + " qhull: " + str(e)[:60])
        s = hostsim.run(xs, base, normal, **knobs)
        rays_ok = True
    else:
        if kind == "slab" or n < d + 2:
            pass
        try:
            truth, rays = qhull_oracle.unbounded(xs)
        except Exception as e:
            return ("skip", tag + " qhull: " + str(e)[:60])
        s = hostsim.run(xs, **knobs)
        got_rays = {tuple(r) for r in s["ray_edge"].tolist()}
        rays_ok = got_rays == rays
    rmax = max(np.linalg.norm(c - xs[k[0] - 1]) for k, c in truth.items() if k[0] <= n)
    if rmax > 1e9:
        # a ray parameter of 1e9 cloud diameters leaves 1e-7 of absolute accuracy at the cloud: neither this walk nor the
        # restated reference (oracle/hv_oracle.cpp misses vertices on such clouds) resolves generators closer than that
        return ("skip", tag + " circumradius %.1e" % rmax)
    if s["stats"]["degenerate"]:
        return ("skip", tag + " flagged degenerate")
    got = {tuple(r) for r in s["sig"].tolist()}
    want = set(truth.keys())
    if got != want or not rays_ok or s["stats"]["seed_fail"]:
        return ("FAIL", tag + " missing=%d extra=%d rays_ok=%s seed_fail=%d" % (len(want - got), len(got - want), rays_ok, s["stats"]["seed_fail"]))
    # coordinates against Qhull's circumcentres, relative to the circumradius
    sig = s["sig"]
    ref = np.array([truth[tuple(r)] for r in sig.tolist()])
    x0 = xs[sig[:, 0] - 1]
    rad = np.linalg.norm(ref - x0, axis=1)
    err = float((np.linalg.norm(s["r"] - ref, axis=1) / np.maximum(rad, 1e-300)).max()) if len(sig) else 0.0
    return ("ok", tag, err)


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    cnt = {"ok": 0, "skip": 0, "FAIL": 0}
    worst = (0.0, "")
    while time.time() - t0 < minutes * 60:
        res = one(seed)
        seed += 1
        if res is None:
            continue
        cnt[res[0]] += 1
        if res[0] == "FAIL":
            print("FAIL", res[1], flush=True)
        elif res[0] == "ok" and res[2] > worst[0]:
            worst = (res[2], res[1])
    print("cases: %s; worst relative coordinate difference to Qhull %.2e (%s); next seed %d" % (cnt, worst[0], worst[1], seed))


if __name__ == "__main__":
    main()
