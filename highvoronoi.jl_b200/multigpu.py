"""Multi-GPU exchange step of the slab-sharded search (parallelmesh.jl:52-87 for the slabs).

One process per GPU.  Every rank searches its slab (RaycastParameter(threading=B200Thread(device, rank, world))), then
the per-rank vertex lists are merged by ONE variable-length all-gather (counts first, rows padded to the maximum) and a
device-side dedup + sort (hvb_merge_device).  torch.distributed is only plumbing here: it works with NCCL on CUDA
tensors and -- for the CPU tests -- with gloo on host tensors."""
import ctypes

import torch
import torch.distributed as dist

from . import _abi


def init_comm(searcher, group=None):
    """process-per-GPU mode: gives the searcher's context its NCCL communicator (hvb_comm_init).  Rank 0 draws the id
    (hvb_comm_unique_id); torch.distributed only carries the 128 bytes -- every collective on the data path is issued
    by the library itself."""
    L, ctx = _abi.lib(), searcher._ctx
    buf = ctypes.create_string_buffer(128)
    if dist.get_rank(group) == 0:
        _abi.check(L.hvb_comm_unique_id(buf), None)
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=0, group=group)
    _abi.check(L.hvb_comm_init(ctx, ctypes.create_string_buffer(box[0], 128)), ctx)


def exchange_counts(searcher):
    """rows owned by every rank (hvb_exchange_counts: ncclAllGather of one word inside the library)"""
    import numpy as np
    L, ctx = _abi.lib(), searcher._ctx
    world = max(1, searcher.parameters.world)
    cnt = np.zeros(world, dtype=np.int64)
    _abi.check(L.hvb_exchange_counts(ctx, cnt.ctypes.data_as(ctypes.c_void_p)), ctx)
    return cnt


def allgather(searcher):
    """replaces the rank's shard by the rows of all ranks (hvb_allgather: counts + compact rows, one fused NCCL group)"""
    L, ctx = _abi.lib(), searcher._ctx
    _abi.check(L.hvb_allgather(ctx), ctx)


def slab_bounds(n, rank, world):
    """contiguous, near-equal index ranges of the spatially sorted order (partition_indices, parallelmesh.jl:52-87)"""
    return n * rank // world, n * (rank + 1) // world


def all_gather_rows(sig, r, count, group=None):
    """sig [cap, d+1] int64, r [cap, d] float64 (first `count` rows valid) on every rank -> concatenation of the valid
    rows of all ranks, identical on every rank.  Returns (sig_all, r_all, bytes_sent_per_rank)."""
    world = dist.get_world_size(group)
    dev = sig.device
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cnts = [int(c.item()) for c in cnts]
    cap = max(max(cnts), 1)
    sig_p = torch.zeros((cap, sig.shape[1]), dtype=torch.int64, device=dev)
    r_p = torch.zeros((cap, r.shape[1]), dtype=torch.float64, device=dev)
    sig_p[:count] = sig[:count]
    r_p[:count] = r[:count]
    sig_l = [torch.empty_like(sig_p) for _ in range(world)]
    r_l = [torch.empty_like(r_p) for _ in range(world)]
    dist.all_gather(sig_l, sig_p, group=group)
    dist.all_gather(r_l, r_p, group=group)
    sig_all = torch.cat([t[:c] for t, c in zip(sig_l, cnts)]).contiguous()
    r_all = torch.cat([t[:c] for t, c in zip(r_l, cnts)]).contiguous()
    return sig_all, r_all, (sig_p.numel() + r_p.numel()) * 8


def gather_and_merge(searcher, group=None, dedup=False, cache=None):
    """exchange step for a searcher that finished its slab search; afterwards the context holds the global result.
    The ranks return disjoint, sorted, owned rows, so the padded all-gather output is installed as it is
    (hvb_adopt_device_padded: no dedup, no re-sort); dedup=True runs the generic hash dedup + global sort instead
    (hvb_merge_device).  `cache` (a dict) keeps the exchange buffers between calls."""
    import numpy as np
    L, ctx = _abi.lib(), searcher._ctx
    d = searcher.dim
    world = dist.get_world_size(group)
    nv = ctypes.c_int64()
    _abi.check(L.hvb_counts(ctx, ctypes.byref(nv), None, None), ctx)
    if dedup:
        cap = max(nv.value, 1)
        sig = torch.empty((cap, d + 1), dtype=torch.int64, device="cuda")
        r = torch.empty((cap, d), dtype=torch.float64, device="cuda")
        got = ctypes.c_int64()
        _abi.check(L.hvb_export_device(ctx, sig.data_ptr(), r.data_ptr(), cap, ctypes.byref(got)), ctx)
        sig_all, r_all, sent = all_gather_rows(sig, r, got.value, group)
        torch.cuda.synchronize()
        _abi.check(L.hvb_merge_device(ctx, sig_all.data_ptr(), r_all.data_ptr(), sig_all.shape[0]), ctx)
        return sent
    # counts of all ranks (one small collective, one host read)
    cnt = torch.tensor([nv.value], dtype=torch.int64, device="cuda")
    cnts_t = torch.empty(world, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(cnts_t, cnt, group=group)
    cnts = cnts_t.cpu().numpy().astype(np.int64)
    cap = int(cnts.max()) if cnts.max() > 0 else 1
    cache = cache if cache is not None else {}
    if cache.get("cap", 0) < cap:                       # exchange buffers grow geometrically and are re-used
        newcap = int(cap * 1.2) + 1024
        cache.update(cap=newcap,
                     sig=torch.empty((newcap, d + 1), dtype=torch.int64, device="cuda"),
                     r=torch.empty((newcap, d), dtype=torch.float64, device="cuda"),
                     sig_all=torch.empty((world, newcap, d + 1), dtype=torch.int64, device="cuda"),
                     r_all=torch.empty((world, newcap, d), dtype=torch.float64, device="cuda"))
    bufcap = cache["cap"]
    got = ctypes.c_int64()
    _abi.check(L.hvb_export_device(ctx, cache["sig"].data_ptr(), cache["r"].data_ptr(), bufcap, ctypes.byref(got)), ctx)
    # all-gather straight into one buffer per array, padded only to the largest count of THIS step (the buffers
    # themselves keep 20 % headroom so that they are not re-allocated from step to step)
    sig_out = cache["sig_all"].view(-1)[:world * cap * (d + 1)]
    r_out = cache["r_all"].view(-1)[:world * cap * d]
    dist.all_gather_into_tensor(sig_out, cache["sig"].view(-1)[:cap * (d + 1)], group=group)
    dist.all_gather_into_tensor(r_out, cache["r"].view(-1)[:cap * d], group=group)
    torch.cuda.synchronize()
    _abi.check(L.hvb_adopt_device_padded(ctx, sig_out.data_ptr(), r_out.data_ptr(), world, cap,
                                         cnts.ctypes.data_as(ctypes.c_void_p)), ctx)
    return cap * (2 * d + 1) * 8
