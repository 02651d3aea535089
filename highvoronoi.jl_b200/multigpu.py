"""Multi-GPU exchange step of the slab-sharded search (parallelmesh.jl:52-87 for the slabs).

One process per GPU.  Every rank searches its slab (RaycastParameter(threading=B200Thread(device, rank, world))), then
the per-rank vertex lists are merged by ONE variable-length all-gather (counts first, rows padded to the maximum) and a
device-side dedup + sort (hvb_merge_device).  torch.distributed is only plumbing here: it works with NCCL on CUDA
tensors and -- for the CPU tests -- with gloo on host tensors."""
import ctypes

import torch
import torch.distributed as dist

from . import _abi


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


def gather_and_merge(searcher, group=None, dedup=False):
    """exchange step for a searcher that finished its slab search; afterwards the context holds the global result.
    The ranks return disjoint, sorted, owned rows, so the gathered concatenation is installed as-is
    (hvb_adopt_device); dedup=True runs the generic hash dedup + global sort instead (hvb_merge_device)."""
    L, ctx = _abi.lib(), searcher._ctx
    d = searcher.dim
    nv = ctypes.c_int64()
    _abi.check(L.hvb_counts(ctx, ctypes.byref(nv), None, None), ctx)
    cap = max(nv.value, 1)
    sig = torch.empty((cap, d + 1), dtype=torch.int64, device="cuda")
    r = torch.empty((cap, d), dtype=torch.float64, device="cuda")
    got = ctypes.c_int64()
    _abi.check(L.hvb_export_device(ctx, sig.data_ptr(), r.data_ptr(), cap, ctypes.byref(got)), ctx)
    sig_all, r_all, sent = all_gather_rows(sig, r, got.value, group)
    torch.cuda.synchronize()
    fn = L.hvb_merge_device if dedup else L.hvb_adopt_device
    _abi.check(fn(ctx, sig_all.data_ptr(), r_all.data_ptr(), sig_all.shape[0]), ctx)
    return sent
