"""VoronoiData mirror (highvoronoi.jl_b200/api.py, voronoidata.jl:571-620) on the CPU: the fields are assembled on the host from
what the C ABI returns, so the assembly is checked here on a stand-in mesh that carries the restated reference's rows and
neighbour lists and the host build's geometry products (tests/hostsim: the same formulas the device kernels run).  The device
wiring itself is tests/test_gpu_volumes.py::test_voronoi_data_fields."""
import numpy as np
import pytest

import hostsim
import qhull_oracle
from util import points


class StandInMesh:
    """duck type of hvb200.VoronoiMesh, filled from arrays"""

    def __init__(self, xs, sig, r, off, ids, base, normal, rays=None, halo=None):
        self.xs, self.sig, self.r, self._nb = xs, sig, r, (off, ids)
        self.base, self.normal = base, normal
        self.n, self.dim = xs.shape[0], xs.shape[1]
        self.n_halo = 0
        e = np.zeros((0, self.dim), dtype=np.int64)
        self.ray_edge, self.ray_base, self.ray_dir, self.ray_node = rays if rays is not None else (e, e * 1.0, e * 1.0, np.zeros(0, dtype=np.int64))
        if halo is not None:
            self.halo_origin, self.halo_xs = halo
            self.n_halo = len(self.halo_origin)
            self.n_user = self.n - self.n_halo

    def neighbors(self):
        return self._nb

    def volumes(self):
        return hostsim.volumes(self.xs, self.sig, self.base, self.normal)

    def moments(self):
        m = hostsim.moments(self.xs, self.sig, self.base, self.normal)               # 1, x_a, x_a x_b (a <= b)
        d = self.dim
        second = np.zeros((self.n, d, d))
        iu = np.triu_indices(d)
        second[:, iu[0], iu[1]] = m[:, 1 + d:]
        second[:, iu[1], iu[0]] = m[:, 1 + d:]
        return m[:, 0].copy(), m[:, 1:1 + d].copy(), second

    def areas(self):
        return hostsim.areas(self.xs, self.sig, self._nb[0], self._nb[1], self.base, self.normal)

    def area_moments(self):
        m = hostsim.area_moments(self.xs, self.sig, self._nb[0], self._nb[1], self.base, self.normal)
        return m[:, 0].copy(), m[:, 1:].copy()

    def origin_of(self, ids):
        import hvb200
        return hvb200.VoronoiMesh.origin_of(self, ids)

    def vertices_iterator(self, i):
        for sg, rr in zip(self.sig, self.r):
            if i in sg:
                yield sg, rr


class StandInGeometry:
    def __init__(self, xs, mesh, domain):
        self.nodes, self.mesh, self.domain = xs, mesh, domain


@pytest.mark.parametrize("d,n", [(2, 300), (3, 200), (4, 80)])
def test_fields_on_a_bounded_domain(hvb, oracle, d, n):
    hostsim.build()
    xs = points(n, d, 900 + d)
    dom = hvb.cuboid(d, periodic=[])
    o = oracle.run(xs, dom.base, dom.normal)
    mesh = StandInMesh(xs, o["sig"], o["r"], o["nb_off"], o["nb_ids"], dom.base, dom.normal)
    vd = hvb.VoronoiData(StandInGeometry(xs, mesh, dom), copyall=True)
    assert vd.offset == 0 and not hasattr(vd, "references")
    assert abs(vd.volume.sum() - 1.0) < 1e-11
    assert np.allclose(vd.bulk_integral[:, 0], vd.volume, rtol=0, atol=1e-13)
    assert vd.bulk_integral.shape == (n, 1 + d + d * d)
    assert np.abs(vd.bulk_integral[:, 1:1 + d].sum(0) - 0.5).max() < 1e-11          # int x_a over the unit cube
    touched = 0
    for i in range(n):
        nb, ori, area, ii = vd.neighbors[i], vd.orientations[i], vd.area[i], vd.interface_integral[i]
        assert len(nb) == len(ori) == len(area) == len(ii) and list(nb) == sorted(nb)
        assert np.array_equal(ii[:, 0], area)
        mid = xs[i] + 0.5 * ori                                                       # voronoidata.jl:788
        for k, j in enumerate(nb):
            if j <= n:
                assert np.array_equal(ori[k], xs[j - 1] - xs[i])
                assert abs(np.linalg.norm(mid[k] - xs[i]) - np.linalg.norm(mid[k] - xs[j - 1])) < 1e-14
            else:
                p = j - n - 1
                assert abs((mid[k] - dom.base[p]) @ dom.normal[p]) < 1e-14            # the midpoint lies on the plane
                assert np.allclose(vd.boundary_nodes[i + 1][p + 1], xs[i] + ori[k], rtol=0, atol=0)
                touched += 1
        # divergence theorem per cell: sum_j area_ij * unit orientation = 0
        unit = ori / np.linalg.norm(ori, axis=1, keepdims=True)
        assert np.abs((area[:, None] * unit).sum(0)).max() < 1e-10
        # volume = 1/d sum_j area_ij * h_ij with h_ij = |orientation| / 2
        assert abs((area * np.linalg.norm(ori, axis=1)).sum() / (2 * d) - vd.volume[i]) < 1e-11
    assert touched == sum(len(v) for v in vd.boundary_nodes.values()) > 0
    # onboundary=True: projections instead of mirror images
    vb = hvb.VoronoiData(StandInGeometry(xs, mesh, dom), getboundary_nodes=True, onboundary=True)
    for i, planes in vb.boundary_nodes.items():
        for p, y in planes.items():
            assert abs((y - dom.base[p - 1]) @ dom.normal[p - 1]) < 1e-14
    # vertices: every cell lists the rows that name it
    assert sum(len(v) for v in vd.vertices) == int((o["sig"] <= n).sum())


def test_boundary_vertices_on_the_unbounded_domain(hvb, oracle):
    xs = points(150, 3, 950)
    o = oracle.run(xs)
    mesh = StandInMesh(xs, o["sig"], o["r"], o["nb_off"], o["nb_ids"], None, None,
                       rays=(o["ray_edge"], o["ray_base"], o["ray_dir"], o["ray_node"]))
    vd = hvb.VoronoiData(StandInGeometry(xs, mesh, hvb.Boundary()), getboundary_vertices=True, getneighbors=True)
    truth, rays = qhull_oracle.unbounded(xs)
    assert set(vd.boundary_vertices) == rays
    for edge, (b, u, node) in vd.boundary_vertices.items():
        assert abs(np.linalg.norm(u) - 1.0) < 1e-12 and 1 <= node <= len(xs)
        # every point of the ray keeps the edge's generators equidistant
        gen = xs[np.array(edge) - 1]
        dist = np.linalg.norm(gen - (b + 0.37 * u), axis=1)
        assert np.abs(dist - dist[0]).max() < 1e-9 * max(1.0, dist[0])


def test_periodic_view_folds_and_sorts(hvb):
    """stand-in for a periodic context in d = 1+1: four nodes on a 2-torus strip are enough to exercise the folding, the
    multiplicities, sorted=True and references / reference_shifts"""
    xs = np.array([[0.1, 0.2], [0.6, 0.3], [0.3, 0.8], [0.8, 0.7]])
    halo_origin = np.array([2, 4, 1])
    shifts = np.array([[-1.0, 0.0], [-1.0, 0.0], [1.0, 0.0]])
    halo_xs = xs[halo_origin - 1] + shifts
    allx = np.vstack([xs, halo_xs])
    # neighbour lists in the extended numbering (4 official, 3 halo): cell 1 touches node 2 twice (directly and through its image 5)
    off = np.array([0, 4, 6, 8, 10, 10, 10, 10])
    ids = np.array([2, 3, 5, 6, 1, 7, 1, 4, 3, 7])
    mesh = StandInMesh(allx, np.zeros((0, 3), dtype=np.int64), np.zeros((0, 2)), off, ids, None, None, halo=(halo_origin, halo_xs))
    mesh.areas = lambda: np.arange(10, dtype=float)
    dom = hvb.cuboid(2)
    vd = hvb.VoronoiData(StandInGeometry(xs, mesh, dom), getneighbors=True, getorientations=True, getarea=True, sorted=True)
    assert [list(nb) for nb in vd.neighbors] == [[2, 2, 3, 4], [1, 1], [1, 4], [1, 3]]
    assert list(vd.area[0]) == [0.0, 2.0, 1.0, 3.0]                                   # areas travel with their neighbours
    assert np.array_equal(vd.orientations[0][1], halo_xs[0] - xs[0])                  # the image the cell really touches
    assert np.array_equal(vd.orientations[1][1], halo_xs[2] - xs[1])
    raw = hvb.VoronoiData(StandInGeometry(xs, mesh, dom), getneighbors=True, getreferences=True, reduce_to_periodic=False)
    assert raw.offset == 3 and list(raw.neighbors[0]) == [2, 3, 5, 6]
    assert np.array_equal(raw.references, halo_origin) and np.array_equal(raw.reference_shifts, shifts)
    assert np.array_equal(allx[4:], xs[raw.references - 1] + raw.reference_shifts)


@pytest.mark.parametrize("d,n,margin", [(2, 300, 0.35), (3, 150, 0.75)])
def test_periodic_view_on_the_explicit_halo_problem(hvb, oracle, d, n, margin):
    """the assertions of tests/test_gpu_volumes.py::test_voronoi_data_fields (periodic part) on the halo problem solved by the
    restated reference: halo copies within `margin`, periodic planes pushed out by it, the backend's numbering
    (caller 1..n, halo n+1..n+nh, planes behind)"""
    import periodic_oracle as po
    xs = points(n, d, 960 + d)
    axes = tuple(range(1, d + 1))
    origin, mult, hxs = po.halo(xs, axes, margin)
    ext = np.vstack([xs, hxs])
    base, normal = po.pushed_cuboid(d, axes, margin)
    o = oracle.run(ext, base, normal)
    mesh = StandInMesh(ext, o["sig"], o["r"], o["nb_off"], o["nb_ids"], base, normal, halo=(origin, hxs))
    mesh.volumes = lambda: hostsim.volumes(ext, o["sig"], base, normal)[:n]
    vp = hvb.VoronoiData(StandInGeometry(xs, mesh, hvb.cuboid(d)), getneighbors=True, getorientations=True, getarea=True, getvolume=True,
                         sorted=True)
    assert abs(vp.volume.sum() - 1.0) < 1e-11
    for i in range(n):
        nb, ori, area = vp.neighbors[i], vp.orientations[i], vp.area[i]
        assert int(nb.max()) <= n and list(nb) == sorted(nb)
        L = np.linalg.norm(ori, axis=1)
        assert abs((area * L).sum() / (2 * d) / vp.volume[i] - 1.0) < 1e-10
        assert np.linalg.norm((area[:, None] * ori / L[:, None]).sum(0)) / area.sum() < 1e-10
        shift = ori - (xs[nb - 1] - xs[i])                                  # a whole number of periods
        assert np.abs(shift - np.round(shift)).max() < 1e-12
    raw = hvb.VoronoiData(StandInGeometry(xs, mesh, hvb.cuboid(d)), getneighbors=True, getreferences=True, reduce_to_periodic=False)
    assert raw.offset == len(origin) and max(int(nb.max()) for nb in raw.neighbors) > n
    assert np.allclose(ext[n:], xs[raw.references - 1] + raw.reference_shifts, rtol=0, atol=0)
