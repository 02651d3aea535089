"""Independent truth for the Voronoi vertex set, built on Qhull (scipy.spatial).  TEST INFRASTRUCTURE ONLY.

It shares no code and no algorithm with the raycast search: interior vertices are circumcentres of Delaunay
simplices, vertices on boundary planes come from a per-cell half-space intersection.  It pins the CPU
restatement (oracle/hv_oracle.cpp), because the reference's own tests hold no golden vectors for this path
(SURVEY.md section 8c).
"""
import numpy as np
from scipy.spatial import Delaunay, HalfspaceIntersection


def cuboid(d, lo=0.0, hi=1.0):
    """Planes of HighVoronoi's cuboid(d, periodic=[]) (boundary.jl:510-534): plane 2i-1 = upper face of axis i,
    plane 2i = lower face; returns (base[P,d], normal[P,d])."""
    base = np.zeros((2 * d, d))
    normal = np.zeros((2 * d, d))
    for i in range(d):
        base[2 * i, :] = lo
        base[2 * i, i] = hi
        normal[2 * i, i] = 1.0
        base[2 * i + 1, :] = lo
        normal[2 * i + 1, i] = -1.0
    return base, normal


def circumcenters(X, simplices):
    P = X[simplices]                       # [S, d+1, d]
    A = 2.0 * (P[:, 1:, :] - P[:, :1, :])
    b = (P[:, 1:, :] ** 2).sum(-1) - (P[:, :1, :] ** 2).sum(-1)
    return np.linalg.solve(A, b[..., None])[..., 0]


def unbounded(X):
    """Boundary() (no planes): every Delaunay simplex is a vertex, every hull facet an unbounded edge.
    Returns (dict sig->r with 1-based sorted sig tuples, set of ray edges)."""
    tri = Delaunay(X)
    cc = circumcenters(X, tri.simplices)
    verts = {tuple(sorted((s + 1).tolist())): c for s, c in zip(tri.simplices, cc)}
    rays = {tuple(sorted((f + 1).tolist())) for f in tri.convex_hull}
    return verts, rays


def bounded(X, base, normal):
    """Vertices of the Voronoi diagram restricted to the convex domain {y : n_p.(y-b_p) <= 0}.
    Plane p (1-based) appears in sig as N+p.  Returns dict sig->r."""
    n, d = X.shape
    P = base.shape[0]
    tri = Delaunay(X)
    indptr, indices = tri.vertex_neighbor_vertices
    verts = {}
    for i in range(n):
        nb = indices[indptr[i]:indptr[i + 1]]
        xi = X[i]
        D = X[nb] - xi                                   # work in coordinates centred at x_i
        hs_n = np.hstack([2.0 * D, -(D ** 2).sum(1, keepdims=True)])
        hs_p = np.hstack([normal, -((base - xi) * normal).sum(1, keepdims=True)])
        hs = HalfspaceIntersection(np.vstack([hs_n, hs_p]), np.zeros(d))
        for pt, facet in zip(hs.intersections, hs.dual_facets):
            if len(facet) != d:
                raise RuntimeError("degenerate vertex in qhull oracle")
            ids = [int(nb[f]) + 1 if f < len(nb) else n + (f - len(nb)) + 1 for f in facet]
            sig = tuple(sorted(ids + [i + 1]))
            verts.setdefault(sig, pt + xi)
    return verts


def voronoi_nongeneral(X, base=None, normal=None, tol=1e-9):
    """Truth for clouds in NON-general position (cubic grids, ...): Qhull's Voronoi diagram of the generators plus their mirror
    images at the planes (scipy.spatial.Voronoi merges cospherical Delaunay facets into one Voronoi vertex).  Returns
    {frozenset of 1-based ids (plane p = n + p): coordinates} for the vertices inside the domain -- what the reference
    returns for such input: one vertex per cospherical set, listing ALL its generators (raycast.jl:926-949).
    TEST INFRASTRUCTURE ONLY."""
    from scipy.spatial import Voronoi
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    P = 0 if base is None else len(base)
    pts = [X]
    for p in range(P):
        s = ((base[p] - X) * normal[p]).sum(1)
        pts.append(X + 2.0 * s[:, None] * normal[p])
    allp = np.vstack(pts)
    vor = Voronoi(allp)
    v2p = {}
    for pi, ri in enumerate(vor.point_region):
        for v in vor.regions[ri]:
            if v >= 0:
                v2p.setdefault(v, set()).add(pi)
    out = {}
    for v, ps in v2p.items():
        c = vor.vertices[v]
        if P and not all(((c - base[p]) @ normal[p]) <= tol for p in range(P)):
            continue
        real = {p for p in ps if p < n}
        if not real:
            continue
        ids = {p + 1 for p in real}
        for p in ps:
            if p >= n and (p - n) % n in real:            # the mirror image of one of the vertex's own generators: a plane
                ids.add(n + (p - n) // n + 1)
        out[frozenset(ids)] = c
    return out
