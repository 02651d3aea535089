"""Multi-GPU paths.  (1) hvb_create_multi: ONE process, one call, several GPUs behind one context -- also run with one
device listed several times, so that the decomposition, the ownership rule and the host-side union are exercised on a
one-GPU box; (2) one process per GPU with the in-library NCCL collectives (hvb_comm_init / hvb_exchange_counts /
hvb_allgather); (3) the older exchange with the collective issued by the host language (hvb_export_device /
hvb_adopt_device_padded / hvb_merge_device).  (2) and (3) are skipped on boxes with fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from util import points

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, n, d, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import hvb200
    from hvb200 import multigpu
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    xs = points(n, d, 31)
    dom = hvb200.cuboid(d, periodic=[])
    res = []
    for dedup in (False, True):          # every rank issues the same collectives: adopt path first, generic dedup + sort second
        s = hvb200.Raycast(xs, domain=dom, options=hvb200.RaycastParameter(threading=hvb200.B200Thread(rank, rank, world), neighbors=1))
        mesh, _ = hvb200.voronoi(xs, searcher=s)
        local = mesh.sig.shape[0]
        multigpu.gather_and_merge(s, dedup=dedup)
        merged = hvb200.VoronoiMesh(s, copy=True)
        off, ids = merged.neighbors()
        res.append((merged.sig.copy(), merged.r.copy(), np.array(off), np.array(ids), local))
        s.close()
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_two_gpu_merge_equals_single(hvb):
    import torch.multiprocessing as mp
    n, d, world = 20000, 3, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, n, d, out), nprocs=world, join=True)
        res = dict(out)
    xs = points(n, d, 31)
    single = hvb.Raycast(xs, domain=hvb.cuboid(d, periodic=[]))
    mesh, _ = hvb.voronoi(xs, searcher=single, copy=True)
    off, ids = mesh.neighbors()
    lo1, hi1 = n * 1 // world, n
    for rank in range(world):
        for dedup, (sig, r, o2, i2, local) in zip((False, True), res[rank]):
            if not dedup:                                                        # adopt path: sorted runs in rank order
                order = np.lexsort(sig.T[::-1])
                sig, r = sig[order], r[order]
            assert np.array_equal(sig, mesh.sig) and np.array_equal(r, mesh.r)  # bitwise: canonical coordinates
            assert local < mesh.sig.shape[0]
            # neighbour lists were built from the slab result: complete for the rank's own cells (grid order slabs)
            assert o2.shape == off.shape
    assert res[0][0][4] + res[1][0][4] == mesh.sig.shape[0]                      # disjoint owned shards


# ---- (1) one process, one context, several GPUs (hvb_create_multi) --------------------------------------------------
def _canon(sig, r):
    order = np.lexsort(sig.T[::-1])
    return sig[order], r[order]


def _multi_devices(k):
    ng = _ngpus()
    return list(range(k)) if ng >= k else [i % max(ng, 1) for i in range(k)]


@pytest.mark.parametrize("d,n,slabs", [(3, 20000, 4), (2, 30000, 3), (5, 1500, 2)])
def test_multi_context_equals_single(hvb, d, n, slabs):
    xs = points(n, d, 51)
    dom = hvb.cuboid(d, periodic=[])
    single = hvb.Raycast(xs, domain=dom)
    ref, _ = hvb.voronoi(xs, searcher=single)
    off1, ids1 = ref.neighbors()
    vol1 = ref.volumes()
    multi = hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(threading=hvb.B200Thread(devices=_multi_devices(slabs))))
    mesh, _ = hvb.voronoi(xs, searcher=multi)
    sig, r = _canon(mesh.sig, mesh.r)
    assert np.array_equal(sig, ref.sig) and np.array_equal(r, ref.r)            # bitwise: canonical coordinates
    off, ids = mesh.neighbors()
    assert np.array_equal(off, off1) and np.array_equal(ids, ids1)
    vol = mesh.volumes()
    assert np.allclose(vol, vol1, rtol=1e-12, atol=0) and abs(vol.sum() - 1.0) < 1e-11
    st = multi.stats()
    assert st["vertices"] == len(ref.sig) and st["raycasts"] >= single.stats()["raycasts"]
    # a second cloud on the same context
    xs2 = points(n // 2, d, 52)
    multi.set_points(xs2)
    single.set_points(xs2)
    a, _ = hvb.voronoi(xs2, searcher=multi)
    b, _ = hvb.voronoi(xs2, searcher=single)
    sa, ra = _canon(a.sig, a.r)
    assert np.array_equal(sa, b.sig) and np.array_equal(ra, b.r)


def test_multi_context_unbounded_rays(hvb):
    xs = points(4000, 3, 53)
    ref, _ = hvb.voronoi(xs, searcher=hvb.Raycast(xs))
    multi = hvb.Raycast(xs, options=hvb.RaycastParameter(threading=hvb.B200Thread(devices=_multi_devices(3))))
    mesh, _ = hvb.voronoi(xs, searcher=multi)
    sig, r = _canon(mesh.sig, mesh.r)
    assert np.array_equal(sig, ref.sig) and np.array_equal(r, ref.r)
    assert sorted(map(tuple, mesh.ray_edge.tolist())) == sorted(map(tuple, ref.ray_edge.tolist()))   # every ray once
    assert np.isinf(mesh.volumes()).sum() == np.isinf(ref.volumes()).sum()


def test_multi_context_periodic(hvb):
    """periodic domain sharded over slabs: halo numbering identical on every slab, rows = the single-GPU rows"""
    xs = points(6000, 3, 54)
    dom = hvb.cuboid(3)
    ref, s1 = hvb.voronoi(xs, searcher=hvb.Raycast(xs, domain=dom, periodic=True))
    multi = hvb.Raycast(xs, domain=dom, periodic=True, options=hvb.RaycastParameter(threading=hvb.B200Thread(devices=_multi_devices(2))))
    mesh, _ = hvb.voronoi(xs, searcher=multi)
    order = np.lexsort(mesh.sig.T[::-1])
    assert np.array_equal(mesh.sig[order], ref.sig) and np.array_equal(mesh.r[order], ref.r)
    assert np.array_equal(mesh.canonical[order], ref.canonical)
    assert multi.stats()["unique_vertices"] == s1.stats()["unique_vertices"]
    assert np.array_equal(mesh.halo_origin, ref.halo_origin)


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_multi_context_allgather(hvb):
    xs = points(30000, 3, 55)
    dom = hvb.cuboid(3, periodic=[])
    ref, _ = hvb.voronoi(xs, searcher=hvb.Raycast(xs, domain=dom))
    off1, ids1 = ref.neighbors()
    multi = hvb.Raycast(xs, domain=dom, options=hvb.RaycastParameter(threading=hvb.B200Thread(ngpus=2)))
    hvb.voronoi(xs, searcher=multi)
    L = hvb._abi.lib()
    hvb._abi.check(L.hvb_allgather(multi._ctx), multi._ctx)
    mesh = hvb.VoronoiMesh(multi, copy=True)
    sig, r = _canon(mesh.sig, mesh.r)
    assert np.array_equal(sig, ref.sig) and np.array_equal(r, ref.r)
    off, ids = mesh.neighbors()
    assert np.array_equal(off, off1) and np.array_equal(ids, ids1)
    assert multi.stats()["exchange_bytes"] > 0
    area = mesh.areas()
    assert np.allclose(area, ref.areas(), rtol=1e-12)


# ---- (2) one process per GPU, collectives issued inside the library -------------------------------------------------
def _worker_inlib(rank, world, port, n, d, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import hvb200
    from hvb200 import multigpu
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # carries the 128-byte NCCL id, nothing else
    res = {}
    for periodic in (False, True):
        xs = points(n, d, 61)
        dom = hvb200.cuboid(d) if periodic else hvb200.cuboid(d, periodic=[])
        s = hvb200.Raycast(xs, domain=dom, periodic=periodic,
                           options=hvb200.RaycastParameter(threading=hvb200.B200Thread(rank, rank, world), periodic_margin=0.01 if periodic else 0.0))
        multigpu.init_comm(s)
        mesh, _ = hvb200.voronoi(xs, searcher=s)
        counts = multigpu.exchange_counts(s)
        assert counts[rank] == mesh.sig.shape[0]
        retries = s.stats()["periodic_retries"]
        multigpu.allgather(s)
        merged = hvb200.VoronoiMesh(s, copy=True)
        off, ids = merged.neighbors()
        halo = (merged.halo_origin.copy(), merged.halo_mult.copy()) if periodic else None
        res[periodic] = (merged.sig.copy(), merged.r.copy(), np.array(off), np.array(ids), counts.copy(), retries, s.stats()["exchange_bytes"], halo)
        s.close()
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_two_process_inlibrary_nccl(hvb):
    import torch.multiprocessing as mp
    n, d, world = 20000, 3, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker_inlib, args=(world, port, n, d, out), nprocs=world, join=True)
        res = dict(out)
    xs = points(n, d, 61)
    for periodic in (False, True):
        dom = hvb.cuboid(d) if periodic else hvb.cuboid(d, periodic=[])
        mesh, _ = hvb.voronoi(xs, searcher=hvb.Raycast(xs, domain=dom, periodic=periodic))
        off, ids = mesh.neighbors()
        for rank in range(world):
            sig, r, o2, i2, counts, retries, xbytes, halo = res[rank][periodic]
            assert (counts > 0).all() and xbytes > 0
            if not periodic:
                order = np.lexsort(sig.T[::-1])
                assert np.array_equal(sig[order], mesh.sig) and np.array_equal(r[order], mesh.r)
                assert counts.sum() == mesh.sig.shape[0]
                assert np.array_equal(o2, off) and np.array_equal(i2, ids)
            else:
                # the tiny first margin forces the ranks to agree on a larger one (ncclAllReduce max inside hvb_search); the
                # halo (and its numbering) then differs from the single-GPU run's, the folded signatures do not
                assert retries >= 1
                assert np.array_equal(res[0][periodic][7][0], halo[0]) and np.array_equal(res[0][periodic][7][1], halo[1])
                import periodic_oracle as po
                assert po.fold_rows(sig, n, halo[0], halo[1]) == po.fold_rows(mesh.sig, n, mesh.halo_origin, mesh.halo_mult)
