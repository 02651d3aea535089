"""2-GPU test of the sharded path: slab search per rank + NCCL all-gather + device merge == single-GPU result.
Skipped on boxes with fewer than two GPUs (the slab logic itself is also covered on one GPU by
test_gpu_parity.py::test_slab_union_equals_full and on the CPU by test_multigpu_gloo.py)."""
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
