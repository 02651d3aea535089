"""CPU tests of the oracle: pins the restatement against the independent Qhull oracle and the golden vectors."""
import glob
import os

import numpy as np
import pytest

import qhull_oracle
from util import PUBLISHED_SCALE_CLOUDS, empty_ball_violations, neighbors_from_sig, points

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    P = g["base"].shape[0]
    o = oracle.run(g["xs"], g["base"], g["normal"]) if P else oracle.run(g["xs"])
    assert np.array_equal(o["sig"], g["sig"])
    assert np.abs(o["r"] - g["r"]).max() <= 1e-12 * max(1.0, np.abs(g["r"]).max())
    assert np.array_equal(o["ray_edge"], g["ray_edge"])
    assert np.array_equal(o["nb_off"], g["nb_off"]) and np.array_equal(o["nb_ids"], g["nb_ids"])


@pytest.mark.parametrize("d,n", [(2, 400), (3, 300), (4, 120), (5, 60)])
def test_oracle_matches_qhull_bounded(oracle, d, n):
    xs = points(n, d, 100 + d)
    base, normal = qhull_oracle.cuboid(d)
    o = oracle.run(xs, base, normal)
    q = qhull_oracle.bounded(xs, base, normal)
    so = [tuple(s) for s in o["sig"].tolist()]
    assert set(so) == set(q)
    assert max(np.abs(o["r"][k] - q[s]).max() for k, s in enumerate(so)) < 1e-11
    assert o["stats"]["degenerate"] == 0
    # the reference's invariant: one raycast per vertex (docs/src/index.md:93,96), no vertex found twice
    assert o["stats"]["duplicates"] == 0


@pytest.mark.parametrize("d,n", [(2, 400), (3, 300), (4, 120)])
def test_oracle_matches_qhull_unbounded(oracle, d, n):
    xs = points(n, d, 200 + d)
    o = oracle.run(xs)
    qv, qr = qhull_oracle.unbounded(xs)
    assert {tuple(s) for s in o["sig"].tolist()} == set(qv)
    assert {tuple(e) for e in o["ray_edge"].tolist()} == qr


def test_oracle_verify_vertex_and_neighbors(oracle):
    xs = points(2000, 3, 7)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    assert empty_ball_violations(o["sig"], o["r"], xs, sample=500) == 0
    off, ids = neighbors_from_sig(o["sig"], 2000)
    assert np.array_equal(off, o["nb_off"]) and np.array_equal(ids, o["nb_ids"])


def test_oracle_multithread_equals_singlethread(oracle):
    xs = points(3000, 3, 9)
    base, normal = qhull_oracle.cuboid(3)
    a = oracle.run(xs, base, normal, nthreads=1)
    b = oracle.run(xs, base, normal, nthreads=4)
    assert np.array_equal(a["sig"], b["sig"])
    assert np.abs(a["r"] - b["r"]).max() < 1e-11


# ---- the other search_settings methods restated (raycast.jl:972-1012 RCOriginal, :504-528 RCCombined, :542-631 RCNonGeneralFast)
@pytest.mark.parametrize("method", ["RCOriginal", "RCCombined", "RCNonGeneralFast"])
@pytest.mark.parametrize("d,n", [(2, 1500), (3, 800), (4, 300), (5, 100)])
def test_oracle_methods_find_the_same_mesh(oracle, method, d, n):
    """test/rcmethods.jl:10-13 runs the four methods and checks sum(volume) = 1 for each; restated, they return the same
    vertex set for a cloud in general position -- which is what lets the device map every `method` value onto one exact
    min-t kernel.  Checked against the default method, against Qhull, bounded and unbounded."""
    xs = points(n, d, 500 + d)
    base, normal = qhull_oracle.cuboid(d)
    a, b = oracle.run(xs, base, normal), oracle.run(xs, base, normal, method=method)
    assert np.array_equal(a["sig"], b["sig"]) and np.abs(a["r"] - b["r"]).max() < 1e-11
    assert np.array_equal(a["nb_ids"], b["nb_ids"]) and b["stats"]["degenerate"] == 0
    assert {tuple(s) for s in b["sig"].tolist()} == set(qhull_oracle.bounded(xs, base, normal))
    ua, ub = oracle.run(xs), oracle.run(xs, method=method)
    assert np.array_equal(ua["sig"], ub["sig"])
    assert sorted(map(tuple, ua["ray_edge"].tolist())) == sorted(map(tuple, ub["ray_edge"].tolist()))
    # the procedures differ: the default method spends 2.6 nn + 1 inrange per ray (docs/src/index.md:93), these fewer
    assert b["stats"]["nn_calls"] < a["stats"]["nn_calls"]


def test_oracle_refuses_a_method_it_does_not_restate(oracle):
    with pytest.raises(ValueError):
        oracle.run(points(50, 3, 1), method=6)          # RCOriginalHP (hull walks only) is not restated


def test_oracle_rejects_too_few_points(oracle):
    with pytest.raises(RuntimeError):
        oracle.run(np.random.default_rng(0).random((3, 3)))


# ---- fixtures produced by the REAL reference (oracle/make_reference_fixtures.jl) --------------------------------------
# None are committed yet: no Julia was available to the builders (the header of oracle/hv_oracle.cpp says "parity unpinned
# by the reference itself").  The reader and the comparison are exercised on a file written in the same format from the
# oracle's own output, so that dropping real files into tests/golden/ref/ is all it takes to pin the oracle.
REF_FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref", "*.txt")))


def read_reference_fixture(path):
    with open(path) as f:
        d, n, P, V = map(int, f.readline().split())
        xs = np.array([list(map(float, f.readline().split())) for _ in range(n)]).reshape(n, d)
        planes = np.array([list(map(float, f.readline().split())) for _ in range(P)]).reshape(P, 2 * d)
        rows = [f.readline().split() for _ in range(V)]
    sig = np.array([[int(t) for t in r[:d + 1]] for r in rows], dtype=np.int64).reshape(V, d + 1)
    r = np.array([[float(t) for t in r[d + 1:]] for r in rows]).reshape(V, d)
    return xs, planes[:, :d], planes[:, d:], sig, r


def check_against_fixture(run, path, tol=1e-10):
    xs, base, normal, sig, r = read_reference_fixture(path)
    o = run(xs, base, normal) if len(base) else run(xs)
    order = np.lexsort(sig.T[::-1])
    assert np.array_equal(o["sig"], sig[order])
    x0 = xs[sig[order][:, 0] - 1]
    rel = np.linalg.norm(o["r"] - r[order], axis=1) / np.maximum(np.linalg.norm(r[order] - x0, axis=1), 1e-300)
    assert rel.max() <= tol, rel.max()


@pytest.mark.parametrize("path", REF_FIXTURES, ids=[os.path.basename(p)[:-4] for p in REF_FIXTURES])
def test_reference_fixtures(oracle, path):
    check_against_fixture(oracle.run, path)


def test_reference_fixture_reader_roundtrip(oracle, tmp_path):
    xs = points(200, 3, 5)
    base, normal = qhull_oracle.cuboid(3)
    o = oracle.run(xs, base, normal)
    p = tmp_path / "ref_selftest.txt"
    with open(p, "w") as f:
        f.write("3 200 6 %d\n" % len(o["sig"]))
        for x in xs:
            f.write(" ".join("%.17g" % v for v in x) + "\n")
        for b, nm in zip(base, normal):
            f.write(" ".join("%.17g" % v for v in list(b) + list(nm)) + "\n")
        for s, r in zip(o["sig"][::-1], o["r"][::-1]):                 # any row order is accepted
            f.write(" ".join(map(str, s)) + " " + " ".join("%.17g" % v for v in r) + "\n")
    check_against_fixture(oracle.run, str(p))


# ---- known answers the reference itself publishes (docs/src/index.md:93,96; tests/golden/ref_published/README.md) -----------
def published():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_published", "index_md_statistics.json")) as f:
        return {int(d): v for d, v in json.load(f).items()}


@pytest.mark.parametrize("d,nodes", [(4, (200, 500, 1000, 1500, 2000)), (5, (200, 500, 1000))])
def test_published_statistics_of_the_reference(oracle, d, nodes):
    """The matrices in docs/src/index.md were made by the real package with its own harness (statistics.jl:98-143: 4 unseeded
    uniform clouds per entry in the unit cube).  The restatement has to reproduce them statistically, on 4 seeded clouds per entry:
      * vertices and boundary vertices (rows 4, 5): properties of the cloud.  Measured spread of ONE cloud: 2.3 % / 0.7 % / 0.6 % / 0.4 %
        of the vertex count at N = 200 / 500 / 1000 / >= 1500 (d = 5: 2.4 % / 1.0 % / 1.0 %), 1.3-1.8 % of the boundary count; two means
        of 4 differ by sigma / sqrt(2): the per-entry bars below are ~4 sigma of that difference, the pooled bars 1 % and 1.5 %;
      * walks (row 6): every vertex is found by exactly one walk except the first vertex of a descent, vertices - walks = descents:
        the reference's counts (39.5 at d = 4, N = 1000) and the restatement's agree entry by entry;
      * nn-searches per walk (row 7): the table dates from the classic incircle iteration (RCOriginal, raycast.jl:972-1012): the
        restated RCOriginal needs 2.57-2.62 nn-searches per walk where the reference printed 2.58-2.63 (within 1.5 %); today's
        default method spends one more (2.85) and RCCombined exactly one nested traversal."""
    pub = published()[d]
    base, normal = qhull_oracle.cuboid(d)
    tot = np.zeros(4)
    tot_pub = np.zeros(4)
    for n in nodes:
        c = pub["nodes"].index(n)
        V, B, D, W, NN = [], [], [], [], []
        for seed in range(4):
            xs = points(n, d, 5000 + 100 * d + seed)
            o = oracle.run(xs, base, normal, method="RCOriginal")
            st = o["stats"]
            V.append(len(o["sig"])); B.append(int((o["sig"] > n).any(axis=1).sum()))
            walks = st["raycasts"] - d * st["descents"]                 # a descent casts d rays (raycast.jl:45-109)
            assert walks == len(o["sig"]) - st["descents"]               # one walk per vertex, the descents' first vertices excepted
            D.append(st["descents"]); W.append(walks); NN.append(st["nn_calls"])
        V, B, D, W = np.mean(V), np.mean(B), np.mean(D), np.mean(W)
        bar = 0.065 if n <= 200 else 0.03 if n <= 500 else 0.02
        assert abs(V / pub["vertices"][c] - 1.0) < bar, (n, V, pub["vertices"][c])
        assert abs(B / pub["boundary_vertices"][c] - 1.0) < 0.06, (n, B, pub["boundary_vertices"][c])
        assert abs(W / pub["walks"][c] - 1.0) < bar
        pub_desc = pub["vertices"][c] - pub["walks"][c]
        assert abs(D - pub_desc) < 0.5 * pub_desc + 4, (n, D, pub_desc)
        assert abs(sum(NN) / (4 * W) / pub["nn_per_walk"][c] - 1.0) < 0.015, (n, sum(NN) / (4 * W), pub["nn_per_walk"][c])
        tot += (V, B, W, D); tot_pub += (pub["vertices"][c], pub["boundary_vertices"][c], pub["walks"][c], pub_desc)
    rel = tot / tot_pub - 1.0
    assert abs(rel[0]) < 0.01 and abs(rel[1]) < 0.015 and abs(rel[2]) < 0.01 and abs(rel[3]) < 0.25, rel


def test_published_vertex_count_at_30000_nodes(oracle):
    """the largest entry of the d = 4 matrix (docs/src/index.md:93): 841 395.0 vertices, 98 515.75 of them on the boundary, averaged over
    4 clouds.  One cloud scatters by 0.08 % (vertices) / 0.55 % (boundary): the means of 4 have to agree to 0.25 % / 1.6 % (4 sigma of
    the difference of two such means); measured: +0.012 % / -0.89 %"""
    d, n = 4, 30000
    c = published()[d]["nodes"].index(n)
    base, normal = qhull_oracle.cuboid(d)
    V, B = [], []
    for k in range(4):
        o = oracle.run(points(n, d, 7000 + 100 * d + k), base, normal, nthreads=min(8, os.cpu_count() or 1))
        V.append(len(o["sig"])); B.append(int((o["sig"] > n).any(axis=1).sum()))
    assert tuple(V) == PUBLISHED_SCALE_CLOUDS[(d, n)]
    assert abs(np.mean(V) / published()[d]["vertices"][c] - 1.0) < 0.0025
    assert abs(np.mean(B) / published()[d]["boundary_vertices"][c] - 1.0) < 0.016


def seeded_counts():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_published", "seeded_counts.json")) as f:
        return {tuple(int(x) for x in k.split(",")): v for k, v in json.load(f).items()}


def test_seeded_counts_follow_the_published_curves():
    """tests/golden/ref_published/seeded_counts.json: the restatement's vertex / boundary-vertex counts on 4 seeded clouds for EVERY
    entry of the published matrices up to 30 000 (d = 4) / 20 000 (d = 5) nodes (tools/published_stats.py counts; minutes of CPU, so
    only spot-checked below).  Their means follow the reference's published means along the whole curve: the relative difference of
    two means of 4 clouds shrinks like 1 / sqrt(N) -- measured |.| sqrt(N) <= 0.46 (vertices) and 1.15 (boundary) over all 36 entries --
    and the totals agree to 0.1 % / 0.3 %"""
    pub, got = published(), seeded_counts()
    assert len(got) == 18 + 14
    tot = np.zeros(4)
    for (d, n), g in got.items():
        c = g["column"]
        assert pub[d]["nodes"][c] == n and len(g["vertices"]) == len(g["boundary_vertices"]) == 4
        pv, pb = pub[d]["vertices"][c], pub[d]["boundary_vertices"][c]
        assert abs(np.mean(g["vertices"]) / pv - 1.0) < 0.8 / np.sqrt(n), (d, n)
        assert abs(np.mean(g["boundary_vertices"]) / pb - 1.0) < 2.0 / np.sqrt(n), (d, n)
        tot += (np.mean(g["vertices"]), pv, np.mean(g["boundary_vertices"]), pb)
    assert abs(tot[0] / tot[1] - 1.0) < 1e-3 and abs(tot[2] / tot[3] - 1.0) < 3e-3


@pytest.mark.parametrize("d,n", [(4, 3000), (4, 8000), (5, 1500)])
def test_seeded_counts_are_the_restatement_s(oracle, d, n):
    g = seeded_counts()[(d, n)]
    base, normal = qhull_oracle.cuboid(d)
    for k in range(4):
        o = oracle.run(points(n, d, 8000 + 1000 * d + 10 * g["column"] + k), base, normal, nthreads=min(8, os.cpu_count() or 1))
        assert len(o["sig"]) == g["vertices"][k] and int((o["sig"] > n).any(axis=1).sum()) == g["boundary_vertices"][k]
