"""Non-general position resolved by perturbation + merge, without a GPU: the search of the
perturbed cloud runs on the host build of the device algorithm (tests/hostsim), the merge of the rows and the neighbour lists are
the library's own host functions (csrc/hvb_nongeneral.hpp) -- the pipeline of Ctx::resolve_degenerate.  Truth: Qhull's Voronoi
diagram of the same cloud, which merges cospherical facets (oracle/qhull_oracle.py::voronoi_nongeneral).  The same checks run
against the CUDA path in tests/test_gpu_degenerate.py."""
import numpy as np
import pytest

import hostsim
import qhull_oracle
from util import points


def grid(m, d):
    return (np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), -1).reshape(-1, d) + 0.5) / m


def check(xs):
    d = xs.shape[1]
    base, normal = qhull_oracle.cuboid(d)
    o = hostsim.resolve(xs, base, normal)
    want = qhull_oracle.voronoi_nongeneral(xs, base, normal)
    got = {frozenset(int(i) for i in o["ids"][o["off"][v]:o["off"][v + 1]]): o["r"][v] for v in range(len(o["off"]) - 1)}
    assert len(got) == len(o["off"]) - 1 and set(got) == set(want)
    assert max(np.abs(got[k] - want[k]).max() for k in want) < 1e-10
    rows = [tuple(int(i) for i in o["ids"][o["off"][v]:o["off"][v + 1]]) for v in range(len(o["off"]) - 1)]
    assert rows == sorted(rows)
    return o, want


@pytest.mark.parametrize("d,m", [(2, 9), (3, 5), (4, 3)])
def test_lattice_on_the_host(d, m):
    o, want = check(grid(m, d))
    assert o["max_siglen"] == 2 ** d and len(want) == (m + 1) ** d and o["simplicial"] > len(want)


def test_lattice_patch_in_a_random_cloud_on_the_host():
    g = 0.3 + 0.4 * grid(5, 3)
    rnd = points(200, 3, 53)
    xs = np.vstack([g, rnd[np.any((rnd < 0.28) | (rnd > 0.72), axis=1)]])
    o, want = check(xs)
    assert o["max_siglen"] == 8 and min(len(k) for k in want) == 4


def test_neighbour_lists_of_a_lattice_on_the_host():
    """a neighbour shares a FULL interface (neighbors.jl:205-212): the face neighbours and the walls, not the cells met at an
    edge or a corner"""
    m, d = 4, 3
    xs = grid(m, d)
    base, normal = qhull_oracle.cuboid(d)
    o = hostsim.resolve(xs, base, normal)
    n = len(xs)
    cell = lambda i, j, k: (i * m + j) * m + k + 1
    for i in range(m):
        for j in range(m):
            for k in range(m):
                c = cell(i, j, k)
                got = [int(v) for v in o["nb_ids"][o["nb_off"][c - 1]:o["nb_off"][c]]]
                face = {cell(*q) for q in ((i - 1, j, k), (i + 1, j, k), (i, j - 1, k), (i, j + 1, k), (i, j, k - 1), (i, j, k + 1)) if all(0 <= t < m for t in q)}
                assert got == sorted(got) and {v for v in got if v <= n} == face and len([v for v in got if v > n]) == 6 - len(face)


def test_perturbation_is_a_function_of_the_index_alone():
    """the same cloud resolves to the same rows twice (k_perturb / perturb_unit depend on the caller's id and the axis)"""
    xs = grid(4, 3)
    base, normal = qhull_oracle.cuboid(3)
    a, b = hostsim.resolve(xs, base, normal), hostsim.resolve(xs, base, normal)
    assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["r"], b["r"]) and np.array_equal(a["nb_ids"], b["nb_ids"])


@pytest.mark.parametrize("d,m", [(2, 6), (3, 4)])
def test_periodic_lattice_through_the_host_side_halo_on_the_host(d, m):
    """a lattice on the unit torus the way the reference periodises (halo generators made by the caller, mirror planes pushed out
    by the margin; domain.jl:175-213): in [0, 1)^d one vertex per class of the torus, m^d of them, each with 2^d generators"""
    import itertools
    g = grid(m, d)
    margin = 1.5 / m
    pts = [g]
    for shift in itertools.product((-1, 0, 1), repeat=d):
        if any(shift):
            c = g + np.array(shift, dtype=float)
            pts.append(c[np.all((c > -margin) & (c < 1 + margin), axis=1)])
    xs = np.vstack(pts)
    base = np.zeros((2 * d, d)); normal = np.zeros((2 * d, d))
    for i in range(d):
        base[2 * i] = -margin; base[2 * i, i] += 1 + 2 * margin; normal[2 * i, i] = 1.0
        base[2 * i + 1] = -margin; normal[2 * i + 1, i] = -1.0
    o = hostsim.resolve(xs, base, normal)
    inside = np.all((o["r"] > -1e-9) & (o["r"] < 1 - 1e-9), axis=1)
    lens = np.diff(o["off"])
    assert int(inside.sum()) == m ** d and np.all(lens[inside] == 2 ** d)
    assert np.abs(o["r"][inside] * m - np.round(o["r"][inside] * m)).max() < 1e-9
