"""Shared helpers of the parity tests."""
import numpy as np


def points(n, d, seed):
    """the workload generator of BASELINE.md section 2: i.i.d. U[0,1)^d, numpy default_rng(seed)"""
    return np.random.default_rng(seed).random((n, d))


def cuboid_planes(d):
    import qhull_oracle
    return qhull_oracle.cuboid(d)


def rel_coord_error(r_a, r_b, xs, sig):
    """max over vertices of |r_a - r_b| / circumradius-scale (distance of the vertex to its first generator, at
    least the coordinate magnitude scale of the cloud)."""
    x0 = xs[sig[:, 0] - 1]
    scale = np.maximum(np.linalg.norm(r_b - x0, axis=1), 1e-300)
    return float((np.linalg.norm(r_a - r_b, axis=1) / scale).max()) if len(sig) else 0.0


def exact_vertex(xs, row, planes=None):
    """the point equidistant to the generators / on the planes of `row` (1-based ids, plane p = n + p), solved in exact
    rational arithmetic from the float inputs and rounded once: the arbiter for ill-conditioned vertices"""
    from fractions import Fraction
    n, d = xs.shape
    F = lambda v: Fraction(float(v))
    ids = [int(g) for g in row]
    x0 = [F(v) for v in xs[ids[0] - 1]]
    A, b = [], []
    for g in ids[1:]:
        if g <= n:
            dx = [F(v) - x0k for v, x0k in zip(xs[g - 1], x0)]
            A.append(dx); b.append(sum(t * t for t in dx) / 2)
        else:
            base, normal = planes
            nrm = [F(v) for v in normal[g - n - 1]]
            off = sum(nk * F(bk) for nk, bk in zip(nrm, base[g - n - 1]))
            A.append(nrm); b.append(off - sum(nk * x0k for nk, x0k in zip(nrm, x0)))
    M = [ai + [bi] for ai, bi in zip(A, b)]
    for c in range(d):
        pv = next(i for i in range(c, d) if M[i][c] != 0)
        M[c], M[pv] = M[pv], M[c]
        for i in range(d):
            if i != c and M[i][c] != 0:
                f = M[i][c] / M[c][c]
                M[i] = [u - f * w for u, w in zip(M[i], M[c])]
    return np.array([float(x0[k] + M[k][d] / M[k][k]) for k in range(d)])


def assert_same_mesh(got_sig, got_r, ref_sig, ref_r, xs, tol=1e-10, planes=None):
    """bit-exact combinatorics (rows already in lexicographic order on both sides), coordinates within tol relative.
    The reference coordinates of a restated walk depend on which walk found the vertex (the multi-threaded oracle
    differs from itself by 2e-10 on slivers of the d = 2 full-size cloud); where the two disagree by more than tol the
    arbiter is the exact rational solution: the checked result must be within tol of THAT."""
    assert got_sig.shape == ref_sig.shape, (got_sig.shape, ref_sig.shape)
    assert np.array_equal(got_sig, ref_sig)
    if not len(ref_sig):
        return 0.0
    x0 = xs[ref_sig[:, 0] - 1]
    scale = np.maximum(np.linalg.norm(ref_r - x0, axis=1), 1e-300)
    err = np.linalg.norm(got_r - ref_r, axis=1) / scale
    bad = np.nonzero(err > tol)[0]
    assert len(bad) <= 50, (len(bad), float(err.max()))
    if planes is None and len(bad):
        import qhull_oracle
        planes = qhull_oracle.cuboid(xs.shape[1])
    for v in bad:
        ex = exact_vertex(xs, ref_sig[v], planes)
        e2 = np.linalg.norm(got_r[v] - ex) / max(np.linalg.norm(ex - x0[v]), 1e-300)
        assert e2 <= tol, (int(v), float(err[v]), float(e2))
        err[v] = e2
    return float(err.max())


def neighbors_from_sig(sig, n):
    """neighbors_of_cell_new (neighbors.jl:219-262) recomputed on the host from a vertex list"""
    d1 = sig.shape[1]
    a = np.repeat(sig, d1, axis=1).ravel()
    b = np.tile(sig, (1, d1)).ravel()
    keep = (a != b) & (a <= n)
    pairs = np.unique(np.stack([a[keep], b[keep]], axis=1), axis=0)
    off = np.searchsorted(pairs[:, 0], np.arange(1, n + 2))
    return off.astype(np.int64), pairs[:, 1].astype(np.int64)


def empty_ball_violations(sig, r, xs, sample=2000, seed=0):
    """verify_vertex (raycast.jl:477-502) restated: no generator strictly inside the ball, generators of sig on it"""
    from scipy.spatial import cKDTree
    n = xs.shape[0]
    rng = np.random.default_rng(seed)
    idx = rng.choice(len(sig), size=min(sample, len(sig)), replace=False)
    tree = cKDTree(xs)
    bad = 0
    for v in idx:
        real = sig[v][sig[v] <= n] - 1
        rad = np.linalg.norm(xs[real[0]] - r[v])
        inside = tree.query_ball_point(r[v], rad * (1 - 1e-9))
        if len(inside) > 0:
            bad += 1
        dev = np.abs(np.linalg.norm(xs[real] - r[v], axis=1) - rad).max()
        if dev > 1e-9 * max(rad, 1e-12):
            bad += 1
    return bad


def area_invariants(xs, vol, off, ids, area, planes=None):
    """checks of interface areas that need no second implementation: (1) vol_i = 1/d * sum_j area_ij * height_ij,
    (2) the divergence theorem sum_j area_ij * normal_ij = 0, (3) symmetry area_ij = area_ji between generators.
    Returns (max relative volume defect, max |divergence| / surface, max symmetry defect relative to the largest area,
    area per boundary plane).  Cells with an infinite entry are skipped."""
    n, d = xs.shape
    base, normal = planes if planes is not None else (np.zeros((0, d)), np.zeros((0, d)))
    P = base.shape[0]
    face = np.zeros(P)
    dv = dd = 0.0
    for i in range(n):
        J, A = ids[off[i]:off[i + 1]], area[off[i]:off[i + 1]]
        if not np.isfinite(A).all() or not np.isfinite(vol[i]):
            continue
        acc, div = 0.0, np.zeros(d)
        for j, a in zip(J, A):
            if j <= n:
                w = xs[j - 1] - xs[i]; L = np.linalg.norm(w)
                acc += a * L / 2 / d; div += a * w / L
            elif j - n - 1 < P:
                p = j - n - 1
                nn = normal[p] / np.linalg.norm(normal[p])
                acc += a * (nn @ (base[p] - xs[i])) / d; div += a * nn; face[p] += a
        dv = max(dv, abs(acc / vol[i] - 1.0)); dd = max(dd, float(np.linalg.norm(div) / A.sum()))
    sym, amax = 0.0, float(area[np.isfinite(area)].max())
    for i in range(n):
        for k in range(off[i], off[i + 1]):
            j = ids[k]
            if j <= n and np.isfinite(area[k]):
                kk = off[j - 1] + np.searchsorted(ids[off[j - 1]:off[j]], i + 1)
                sym = max(sym, abs(area[k] - area[kk]) / amax)
    return dv, dd, sym, face


# vertex counts of the seeded clouds points(n, d, 7000 + 100 d + k), k = 0..3, by the restatement (8 threads): the golden counts the
# device path has to reproduce exactly (tests/test_gpu_volumes.py::test_published_vertex_counts_of_the_reference)
PUBLISHED_SCALE_CLOUDS = {(4, 30000): (840958, 842244, 842038, 840746), (5, 20000): (2683558, 2690646, 2684057, 2689046)}
