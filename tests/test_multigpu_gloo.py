"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slab partition, variable-length all-gather, dedup.
The per-rank slab results are cut out of the oracle's result (a rank finds every vertex that touches its slab)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import qhull_oracle
from conftest import ROOT
from util import points


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sig_full, r_full, n, out):
    sys.path.insert(0, ROOT)
    import hvb200
    from hvb200 import multigpu
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = multigpu.slab_bounds(n, rank, world)
    touches = ((sig_full > lo) & (sig_full <= hi)).any(axis=1)          # 1-based ids lo+1 .. hi
    sig = torch.from_numpy(sig_full[touches])
    r = torch.from_numpy(r_full[touches])
    pad = 7 + rank                                                       # buffers larger than the valid part
    sig_b = torch.cat([sig, torch.full((pad, sig.shape[1]), -1, dtype=torch.int64)])
    r_b = torch.cat([r, torch.zeros((pad, r.shape[1]), dtype=torch.float64)])
    sig_all, r_all, sent = multigpu.all_gather_rows(sig_b, r_b, sig.shape[0])
    uniq, first = np.unique(sig_all.numpy(), axis=0, return_index=True)  # deterministic dedup (sorted rows)
    out[rank] = (uniq, r_all.numpy()[first], int(sig.shape[0]), sent)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_gather_dedup_world2(oracle):
    n, d, world = 1500, 3, 2
    xs = points(n, d, 21)
    base, normal = qhull_oracle.cuboid(d)
    o = oracle.run(xs, base, normal)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), o["sig"], o["r"], n, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for rank in range(world):
        uniq, r, local, sent = res[rank]
        assert np.array_equal(uniq, o["sig"])                 # union of the slabs == full set, identical on every rank
        assert np.array_equal(r, o["r"])
        assert local < len(o["sig"]) and sent > 0             # each rank held a strict subset
    assert res[0][2] + res[1][2] > len(o["sig"])              # vertices straddling the slab boundary are found twice


def test_slab_bounds_partition():
    sys.path.insert(0, ROOT)
    from hvb200 import multigpu
    for n, w in [(10, 3), (100000, 8), (7, 7), (5, 8)]:
        b = [multigpu.slab_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
