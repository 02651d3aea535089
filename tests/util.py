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


def assert_same_mesh(got_sig, got_r, ref_sig, ref_r, xs, tol=1e-10):
    """bit-exact combinatorics (rows already in lexicographic order on both sides), coordinates within tol relative"""
    assert got_sig.shape == ref_sig.shape, (got_sig.shape, ref_sig.shape)
    assert np.array_equal(got_sig, ref_sig)
    err = rel_coord_error(got_r, ref_r, xs, ref_sig)
    assert err <= tol, err
    return err


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
