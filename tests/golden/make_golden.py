"""Generates the committed golden vectors tests/golden/*.npz.

The reference ships no golden vectors for this path and Julia cannot run in the build container, so the vectors
are produced by the CPU restatement (oracle/hv_oracle.cpp) and accepted only if the independent Qhull oracle
(oracle/qhull_oracle.py) yields the identical vertex set and coordinates within 1e-11.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hv_oracle  # noqa: E402
import qhull_oracle  # noqa: E402

CASES = [  # name, d, n, seed, bounded
    ("d2_n60_cube", 2, 60, 11, True), ("d2_n60_free", 2, 60, 11, False),
    ("d3_n50_cube", 3, 50, 12, True), ("d3_n50_free", 3, 50, 12, False),
    ("d4_n40_cube", 4, 40, 13, True), ("d4_n40_free", 4, 40, 13, False),
    ("d5_n30_cube", 5, 30, 14, True), ("d6_n24_cube", 6, 24, 15, True),
]

for name, d, n, seed, bounded in CASES:
    xs = np.random.default_rng(seed).random((n, d))
    if bounded:
        base, normal = qhull_oracle.cuboid(d)
        o = hv_oracle.run(xs, base, normal)
        q = qhull_oracle.bounded(xs, base, normal)
        qrays = set()
    else:
        base = normal = np.zeros((0, d))
        o = hv_oracle.run(xs)
        q, qrays = qhull_oracle.unbounded(xs)
    so = [tuple(s) for s in o["sig"].tolist()]
    assert set(so) == set(q), name
    err = max(np.abs(o["r"][k] - q[s]).max() / max(1.0, np.abs(q[s]).max()) for k, s in enumerate(so))
    assert err < 1e-11, (name, err)
    assert {tuple(e) for e in o["ray_edge"].tolist()} == qrays, name
    assert o["stats"]["degenerate"] == 0 and o["stats"]["duplicates"] == 0
    np.savez_compressed(os.path.join(os.path.dirname(__file__), name + ".npz"), xs=xs, base=base, normal=normal,
                        sig=o["sig"], r=o["r"], ray_edge=o["ray_edge"], nb_off=o["nb_off"], nb_ids=o["nb_ids"])
    print(name, len(so), "vertices", len(o["ray_edge"]), "rays, qhull agreement", err)
