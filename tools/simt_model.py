"""Lock-step cost model of the min-t query for design decisions without a GPU.

The host build of the device algorithm (tests/hostsim, HVB_TRACE_EVENT hooks in hvb_core.cuh) records for every ray the
sequence of row tests and scanned rows (with their point counts and FP32 survivors).  Rays are grouped 32 at a time in
processing order (the queue order of the walk) and the warp-instruction cost of one warp is evaluated for
  nested : today's kernel -- one lane per ray, a loop per level (rows, chunks of a row), lanes wait for the slowest
  pooled : the next design -- the row search stays with the owner lane, the POINTS of the <= 32
           current rows are flattened over the warp
  ideal  : every lane-operation at 32 of 32 threads.
Costs are instruction estimates from the SASS of k_walk_coop (per row test R, per chunk of 4 points C, per FP32 survivor
S, per pooled point P, per pooled batch with a survivor S2).  Only the query is modelled; setup and commit are uniform.

usage: python tools/simt_model.py [d n [sort_window]]"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
for p in ("tests", os.path.join("tests", "hostsim"), "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import hostsim        # noqa: E402
import qhull_oracle   # noqa: E402

COST = {  # d: (R, C, S, P, S2)
    2: (60, 80, 45, 34, 25), 3: (70, 100, 45, 40, 25), 4: (100, 110, 50, 46, 25), 5: (110, 150, 55, 60, 25), 6: (120, 170, 55, 68, 25)}


def rays_from_trace(tr, d):
    """-> list of rays; a ray = list of stages; a stage = list of steps (tests, points, survivors): `tests` row tests,
    then one scanned row of `points` points (0: the row search ended without a hit)"""
    rays, steps, tests = [], None, 0
    stage_list = None
    for kind, val in tr:
        if kind == 0:
            if stage_list is not None:
                if steps is not None:
                    if tests:
                        steps.append([tests, 0, 0])
                    stage_list.append(steps)
                rays.append(stage_list)
            stage_list, steps, tests = [], None, 0
        elif kind == 3:
            if steps is not None:
                if tests:
                    steps.append([tests, 0, 0])
                stage_list.append(steps)
            steps, tests = [], 0
        elif kind == 1:
            tests += 1
        elif kind == 2:
            if d >= 4:
                steps.append([tests, int(val), 0]); tests = 0          # the tests before this hit, then the row
            else:
                steps.append([1, int(val), 0]); tests = 0              # d <= 3: every iteration is one geometry + one (maybe empty) row
        elif kind == 4:
            steps[-1][2] += 1
    if stage_list is not None:
        if steps is not None:
            if tests:
                steps.append([tests, 0, 0])
            stage_list.append(steps)
        rays.append(stage_list)
    return rays


def model(rays, d, order=None):
    R, C, S, P, S2 = COST[d]
    idx = np.arange(len(rays)) if order is None else order
    nested = pooled = ideal = 0.0
    lane_ops = 0.0
    for w in range(0, len(idx), 32):
        grp = [rays[i] for i in idx[w:w + 32]]
        nst = max(len(r) for r in grp)
        for s in range(nst):
            st = [r[s] for r in grp if len(r) > s]
            for k in range(max(len(x) for x in st)):
                steps = [x[k] for x in st if len(x) > k]
                t = np.array([a[0] for a in steps]); n = np.array([a[1] for a in steps]); sv = np.array([a[2] for a in steps])
                ch = (n + 3) // 4
                nested += R * t.max() + C * ch.max() + S * sv.max()
                batches = int(np.ceil(n.sum() / 32.0))
                pooled += R * t.max() + (P + (S2 if sv.sum() else 0)) * batches
                ops = R * t.sum() + C * ch.sum() + S * sv.sum()
                ideal += ops / 32.0
                lane_ops += ops
    return nested, pooled, ideal


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    win = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    xs = np.random.default_rng(0).random((n, d))
    base, normal = qhull_oracle.cuboid(d)
    o = hostsim.run(xs, base, normal, trace=True)
    rays = rays_from_trace(o["trace"], d)
    tests = sum(a[0] for r in rays for s in r for a in s)
    pts = sum(a[1] for r in rays for s in r for a in s)
    rows = sum(1 for r in rays for s in r for a in s if a[1] > 0)
    surv = sum(a[2] for r in rays for s in r for a in s)
    print("d=%d n=%d: %d rays, per ray %.1f row tests, %.1f non-empty rows, %.1f points, %.1f FP32 survivors, %.3f stages"
          % (d, n, len(rays), tests / len(rays), rows / len(rays), pts / len(rays), surv / len(rays), sum(len(r) for r in rays) / len(rays)))
    nested, pooled, ideal = model(rays, d)
    print("  queue order    : nested %.0f warp-instr/ray (efficiency %.2f)   pooled points %.0f (%.2fx fewer)   ideal %.0f"
          % (nested / len(rays), ideal / nested, pooled / len(rays), nested / pooled, ideal / len(rays)))
    # rays sorted by their number of row tests inside windows of `win` consecutive rays (what sorting by expected cost could give)
    key = np.array([sum(a[0] for s in r for a in s) for r in rays])
    order = np.concatenate([w0 + np.argsort(key[w0:w0 + win], kind="stable") for w0 in range(0, len(rays), win)])
    nested_s, pooled_s, _ = model(rays, d, order)
    print("  sorted (win %d): nested %.0f (%.2fx fewer than unsorted)   pooled points %.0f" % (win, nested_s / len(rays), nested / nested_s, pooled_s / len(rays)))


if __name__ == "__main__":
    main()
