"""Periodic tessellations: test-side restatements.  TEST INFRASTRUCTURE ONLY (see oracle/hv_oracle.py).

Two independent truths for a cuboid domain with periodic axes:
  * halo(): the explicit halo problem the reference solves (reflect_nodes domain.jl:338-390 produces shifted copies
    with `references` / `reference_shifts`; expand_internal_boundary pushes the periodic planes outwards): caller
    generators + every periodic image within `margin` outside the periodic faces, in the numbering the product
    documents (include/hvb200.h: by generator, then by shift code, pair 0 fastest).  Feeding it to the CPU
    restatement (hv_oracle.run) gives the reference-algorithm answer on the extended point set.
  * torus_simplices(): Qhull on the full 3^k replication; a Delaunay simplex that touches the central copy, written
    as a set of (origin, shift) pairs, is a vertex of the periodic tessellation (shares no code with either).
"""
import itertools

import numpy as np
from scipy.spatial import Delaunay


def halo(xs, periodic_axes, margin, lo=0.0, hi=1.0):
    """-> (origin[nh] 1-based, mult[nh, k] int, hxs[nh, d]) for the unit-cuboid with the given 1-based periodic axes"""
    n, d = xs.shape
    axes = [a - 1 for a in periodic_axes]            # pair k <-> axis axes[k] (planes 2a+1 / 2a+2, ascending)
    width = hi - lo
    K = max(1, int(np.ceil(margin / width)))
    w = 2 * K + 1
    rows = []
    for code in range(w ** len(axes)):
        km, rem = [], code
        for _ in axes:
            km.append(rem % w - K)
            rem //= w
        if not any(km):
            continue
        y = xs.copy()
        for k, a in enumerate(axes):
            y[:, a] = y[:, a] + float(km[k]) * width          # same operation order as the device kernel
        ok = np.ones(n, dtype=bool)
        for a in axes:
            ok &= ~(y[:, a] > hi + margin) & ~(-y[:, a] > -lo + margin)
        idx = np.nonzero(ok)[0]
        rows.append((idx, np.full(idx.shape, code), np.tile(np.array(km, dtype=np.int32), (idx.shape[0], 1)), y[idx]))
    if not rows:
        return np.zeros(0, np.int64), np.zeros((0, len(axes)), np.int32), np.zeros((0, d))
    node = np.concatenate([r[0] for r in rows]); code = np.concatenate([r[1] for r in rows])
    mult = np.concatenate([r[2] for r in rows]); hx = np.concatenate([r[3] for r in rows])
    order = np.lexsort((code, node))
    return node[order].astype(np.int64) + 1, mult[order], hx[order]


def pushed_cuboid(d, periodic_axes, margin, lo=0.0, hi=1.0):
    """planes of cuboid(d) with the periodic faces pushed outwards by margin (plane 2i-1 upper, 2i lower face of axis i)"""
    base = np.zeros((2 * d, d)); normal = np.zeros((2 * d, d))
    for i in range(d):
        m = margin if (i + 1) in periodic_axes else 0.0
        base[2 * i, :] = lo; base[2 * i, i] = hi + m; normal[2 * i, i] = 1.0
        base[2 * i + 1, :] = lo; base[2 * i + 1, i] = lo - m; normal[2 * i + 1, i] = -1.0
    return base, normal


def torus_simplices(xs, periodic_axes, lo=0.0, hi=1.0, centres=False):
    """set of frozenset((origin0, shift tuple)) over the Delaunay simplices of the 3^k replication that touch the
    central copy.  Valid when no circumball of such a simplex leaves the replicated box (checked when every axis is
    periodic).  centres=True returns a dict simplex -> circumcentre."""
    n, d = xs.shape
    axes = [a - 1 for a in periodic_axes]
    shifts = list(itertools.product((-1, 0, 1), repeat=len(axes)))
    pts, tag = [], []
    for s in shifts:
        y = xs.copy()
        for k, a in enumerate(axes):
            y[:, a] += s[k] * (hi - lo)
        pts.append(y); tag += [s] * n
    X = np.vstack(pts)
    tri = Delaunay(X)
    zero = shifts.index(tuple([0] * len(axes)))
    central = (tri.simplices // n == zero).any(axis=1)
    simp = tri.simplices[central]
    # circumballs must stay inside the replicated box along the periodic axes
    P = X[simp]
    A = 2.0 * (P[:, 1:, :] - P[:, :1, :]); b = (P[:, 1:, :] ** 2).sum(-1) - (P[:, :1, :] ** 2).sum(-1)
    flat = np.abs(np.linalg.det(A)) < 1e-280          # zero-volume hull simplices of a partially replicated cloud
    A[flat] = np.eye(d)
    cc = np.linalg.solve(A, b[..., None])[..., 0]
    cc[flat] = np.inf
    R = np.linalg.norm(cc - P[:, 0, :], axis=1)
    if len(axes) == d:
        for a in axes:
            assert (cc[:, a] + R < hi + (hi - lo)).all() and (cc[:, a] - R > lo - (hi - lo)).all(), "replication too small"
    out = {}
    for row, c in zip(simp, cc):
        out[frozenset((int(v % n), tag[v]) for v in row)] = c
    return out if centres else set(out)


def fold_rows(sig, n, origin, mult):
    """rows of extended ids (1-based; caller 1..n, halo n+1..) -> set of frozenset((origin0, shift tuple)); rows with plane ids are skipped"""
    k = mult.shape[1]
    nh = origin.shape[0]
    out = set()
    zero = tuple([0] * k)
    for row in sig:
        if (row > n + nh).any():
            continue
        out.add(frozenset((int(g - 1), zero) if g <= n else (int(origin[g - n - 1] - 1), tuple(int(t) for t in mult[g - n - 1])) for g in row))
    return out


def canonical_classes(simplices):
    """number of classes modulo lattice translation"""
    cls = set()
    for s in simplices:
        items = sorted(s)
        o0, s0 = items[0]
        cls.add(frozenset((o, tuple(a - b for a, b in zip(sh, s0))) for o, sh in items))
    return len(cls)
