#!/usr/bin/env python
"""bench.py -- Voronoi vertices / second of the raycast vertex search (BASELINE.json metric).

One "step" = one pass of the hot path over one synthetic point cloud: voronoi(xs; searcher=Raycast(xs; domain))
through the C ABI of libhvb200.so.  N = 1: configs[1] of BASELINE.json (C2: 100 000 uniform points, d = 3, cuboid
boundary, vertex search + neighbours).  N > 1: the same density per GPU -- N x 100 000 points, generators
replicated, GPU k walks slab k of the spatially sorted order (parallelmesh.jl:52-87), vertex lists merged by one
NCCL all-gather + deterministic dedup -- i.e. weak scaling along the reference's own decomposition.

  value : whole-job vertices/s of the hot path on the device, result (sorted vertex rows + neighbour lists) complete in
          HBM: spatial index build (the reference times Raycast(xs) + voronoi(), statistics.jl:98-126) + hvb_search,
          device time from CUDA events on the library's stream (ms_build - ms_upload + ms_search + ms_finalize: the
          generators count as resident, their upload and the page-locked D2H staging of the result (ms_stage_wait)
          belong to e2e, not here); for N > 1 plus the all-gather + merge, max over ranks
  e2e   : the same through the public API from HOST buffers: hvb_create (H2D + index build) + hvb_search +
          hvb_fetch_vertices + hvb_fetch_neighbors (D2H), wall clock with the device idle on both sides
  --impl reference : the CPU restatement of the reference algorithm (oracle/, the reference is Julia and cannot
          run here) on all host threads, on a bounded sample of the same workload
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per vertex, SURVEY.md section 8(d) / BASELINE.md section 3
B_ALG = {2: 240.0, 3: 429.0, 4: 740.0, 5: 1344.0, 6: 2755.0}
WORKLOADS = {  # name -> (points per GPU, dim)
    "C2": (100000, 3), "C2x4": (400000, 3), "C1": (1000, 3), "C3": (1000000, 2), "C4": (50000, 5), "C4s": (20000, 5), "D4": (30000, 4), "D6": (4000, 6),
    # periodic unit cube, cuboid(d) with every axis periodic (hvb_create_periodic): C5 = configs[4] of BASELINE.json
    "C5": (20000, 6), "C5s": (4000, 6), "P3": (100000, 3), "P2": (1000000, 2),
}
PERIODIC = {"C5", "C5s", "P3", "P2"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed ncu capture
    (profiles/r1_traffic.json; null when no capture exists for this workload)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return json.load(f).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cloud(n, d, seed):
    return np.random.default_rng(seed).random((n, d))


def run_reference(args, n_per_gpu, d, rank, world):
    """the reference arm: CPU restatement of the reference algorithm, all host threads, bounded sample"""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hv_oracle
    import qhull_oracle
    hv_oracle.build()
    cores = min(os.cpu_count() or 1, 8)              # the reference cannot use more than 8 threads (chull.jl:185-193)
    n_sample = min(n_per_gpu * world, max(2000, int(args.ref_points)))
    base, normal = qhull_oracle.cuboid(d)
    times, verts = [], 0
    for it in range(args.warmup + args.steps):
        xs = cloud(n_per_gpu * world, d, it)[:n_sample]
        o = hv_oracle.run(xs, base, normal, nthreads=cores)
        # the reference's protocol (statistics.jl:98-126) times Raycast(xs) + voronoi(): index build + cell loop;
        # the oracle's sorting / neighbour post-processing for the tests is not part of it
        dt = (o["stats"]["search_us"] + o["stats"]["build_us"]) * 1e-6
        if it >= args.warmup:
            times.append(dt)
            verts += len(o["sig"])
    T = sum(times)
    val = verts / T
    sample = ("the whole workload, %d points" % n_sample) if n_sample == n_per_gpu * world else \
        "first %d of the %d points of the workload (same density is not preserved: fewer, larger cells)" % (n_sample, n_per_gpu * world)
    line = {"impl": "reference", "metric": "voronoi_vertices_per_sec", "value": val, "unit": "vertices/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s sample: %d uniform points, d=%d, cuboid(d,periodic=[])" % (args.workload, n_sample, d),
                       "threads": cores},
            "cpu_baseline": {"value": val, "unit": "vertices/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of the reference algorithm (oracle/hv_oracle.cpp), not Julia"},
            "e2e": {"value": val, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-points", type=float, default=100000)
    ap.add_argument("--cpu-points", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--setting", action="append", default=[], help="backend knob, e.g. tile_size=8")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_per_gpu, d = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, n_per_gpu, d, rank, world)

    import torch
    import hvb200
    from hvb200 import _abi
    L = _abi.lib()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    settings = {}
    for kv in args.setting:
        k, v = kv.split("=")
        settings[k] = float(v) if "." in v else int(v)
    n_total = n_per_gpu * world
    periodic = args.workload in PERIODIC
    if periodic and world > 1:
        raise SystemExit("periodic workloads run on one GPU (periodic contexts are not sharded yet)")
    dom = hvb200.cuboid(d) if periodic else hvb200.cuboid(d, periodic=[])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_merge(searcher):
        """multi-GPU exchange: counts + padded rows by NCCL all-gather, dedup + sort on every rank"""
        from hvb200 import multigpu
        return multigpu.gather_and_merge(searcher, cache=state.setdefault("xchg", {}))

    state = {"s": None}
    phases = np.zeros(4)

    def step(it, timed):
        """returns (device_ms, e2e_s, vertices, stats, launches, h2d_bytes, d2h_bytes)"""
        if "xs_pin" not in state:                        # the step's input lives in page-locked host memory
            state["xs_pin"] = torch.empty((n_total, d), dtype=torch.float64, pin_memory=True)
        xs = state["xs_pin"].numpy()
        xs[:] = cloud(n_total, d, it)                    # new synthetic cloud every step
        flush.fill_(it & 0xff)
        barrier()
        t0 = time.perf_counter()
        if state["s"] is None:                           # the context (device + page-locked buffers) is re-used
            opts = hvb200.RaycastParameter(threading=hvb200.B200Thread(local_rank, rank, world), neighbors=1, **settings)
            state["s"] = hvb200.Raycast(xs, domain=dom, options=opts, periodic=periodic)
        else:
            state["s"].set_points(xs)                    # H2D + index build
        s = state["s"]
        st0 = s.stats()                                  # index build of THIS step (hvb_create / hvb_set_points) without the upload
        build_ms = st0["ms_build"] - st0["ms_upload"]
        t1 = time.perf_counter()
        hvb200_mesh_rc = L.hvb_search(s._ctx, None, 0, None, None, 0, 0)
        _abi.check(hvb200_mesh_rc, s._ctx)
        st = s.stats()
        dev_ms = build_ms + st["ms_search"] + st["ms_finalize"]
        if dist is not None:
            te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            te0.record()
            gather_merge(s)
            te1.record()
            torch.cuda.synchronize()
            dev_ms += te0.elapsed_time(te1)
        t1b = time.perf_counter()
        if dist is None:
            mesh = hvb200.VoronoiMesh(s)                           # D2H of vertices (and rays)
            sig_h, r_h = mesh.sig, mesh.r
            # periodic: vertices are counted once per image class ("excl. halo duplicates")
            V = st["unique_vertices"] if periodic else sig_h.shape[0]
            t1c = time.perf_counter()
            off, ids = mesh.neighbors()                            # neighbour lists + D2H
        else:
            # every rank keeps its shard of the merged, sorted list and the neighbour lists of its own cells
            nv = ctypes.c_int64()
            _abi.check(L.hvb_counts(s._ctx, ctypes.byref(nv), None, None), s._ctx)
            V = nv.value
            lo, hi = V * rank // world, V * (rank + 1) // world
            if "sig_h" not in state:                         # page-locked host buffers of the caller, allocated once
                cap_h = int(1.3 * V / world) + 1024
                state["sig_h"] = torch.empty((cap_h, d + 1), dtype=torch.int64, pin_memory=True).numpy()
                state["r_h"] = torch.empty((cap_h, d), dtype=torch.float64, pin_memory=True).numpy()
            sig_h, r_h = state["sig_h"], state["r_h"]
            _abi.check(L.hvb_fetch_vertices_range(s._ctx, lo, hi - lo, sig_h.ctypes.data_as(ctypes.c_void_p),
                                                  r_h.ctypes.data_as(ctypes.c_void_p)), s._ctx)
            sig_h, r_h = sig_h[:hi - lo], r_h[:hi - lo]
            t1c = time.perf_counter()
            po, pi, tot = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
            _abi.check(L.hvb_view_neighbors(s._ctx, ctypes.byref(po), ctypes.byref(pi), ctypes.byref(tot)), s._ctx)
            off = np.empty(n_total + 1, dtype=np.int64); ids = np.empty(int(tot.value), dtype=np.int64)   # byte accounting of the staged CSR
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if timed:
            phases[:] += (t1 - t0, t1b - t1, t1c - t1b, t2 - t1c)
        h2d = xs.nbytes
        d2h = sig_h.nbytes + r_h.nbytes + off.nbytes + ids.nbytes
        st2 = s.stats()
        return dev_ms, t2 - t0, V, st, st2["kernel_launches"], h2d, d2h

    for it in range(args.warmup):
        step(1000 + it, False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms_tot, e2e_tot, verts, launches, kern_ms, kern_launches, h2d, d2h = 0.0, 0.0, 0, 0, 0.0, 0, 0, 0
    stats_last = None
    step_ms = []
    for it in range(args.steps):
        dm, es, V, st, nl, hb, db = step(it, True)
        dev_ms_tot += all_max(dm)
        step_ms.append(round(dm, 3))
        e2e_tot += all_max(es)
        verts += V
        launches += nl
        kern_ms += st["ms_expand_kernel"]
        kern_launches += st["expand_launches"]
        h2d, d2h, stats_last = hb, db, st
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = verts / (dev_ms_tot * 1e-3)
    e2e = verts / e2e_tot
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel (k_expand): algorithmic bytes per launch / average launch duration
    v_rank = stats_last["vertices"]
    bytes_per_launch = B_ALG[d] * (v_rank * args.steps) / max(kern_launches, 1)
    avg_launch_s = kern_ms * 1e-3 / max(kern_launches, 1)
    achieved = bytes_per_launch / avg_launch_s / 1e9
    line = {
        "metric": "voronoi_vertices_per_sec", "value": value, "unit": "vertices/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d uniform points per GPU (%d total), d=%d, %s, vertices + neighbours"
                               % (args.workload, n_per_gpu, n_total, d,
                                  "cuboid(d) all axes periodic: halo generators + certificate on the device" if periodic else "cuboid(d,periodic=[])"),
                   "parallelism": "slab%d" % world, "l2": "256 MiB L2 flush before every step; steps timed one by one and summed",
                   "settings": settings},
        "e2e": {"value": e2e, "unit": "vertices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_tot / args.steps,
                "phase_ms": dict(zip(("set_points", "search", "fetch_vertices", "neighbors"), (1e3 * phases / args.steps).round(3).tolist()))},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": {0: "k_expand<%d>", 1: "k_walk<%d>", 2: "k_walk_coop<%d,pooled query>", 3: "k_walk_coop<%d>"}[settings.get("persistent", 3)] % d, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": measured_traffic(args.workload), "peak_source": peak_src,
                     "bytes_per_vertex": B_ALG[d], "launches_per_step": kern_launches / args.steps,
                     "kernel_ms_per_step": kern_ms / args.steps},
        "vertices_per_step": verts / args.steps,
        "stats_last_step": {k: stats_last[k] for k in ("raycasts", "duplicate_hits", "closed_skips", "candidates_fp32",
                                                        "candidates_fp64", "rounds", "seeds", "ms_build", "ms_upload", "ms_search", "ms_finalize", "ms_seed", "ms_neighbors", "ms_rows_sort", "ms_stage_wait", "capacity_retries",
                                                        "vertices", "unique_vertices", "halo_nodes", "periodic_retries")},
        "step_ms_list": step_ms,
    }
    if not args.no_cpu_baseline and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import hv_oracle
        import qhull_oracle
        base, normal = qhull_oracle.cuboid(d)
        n_s = min(n_per_gpu, args.cpu_points)
        xs = cloud(n_per_gpu, d, 0)[:n_s]
        o = hv_oracle.run(xs, base, normal, nthreads=1)
        dt = (o["stats"]["search_us"] + o["stats"]["build_us"]) * 1e-6
        line["cpu_baseline"] = {"value": len(o["sig"]) / dt, "unit": "vertices/s", "cores": 1, "kind": "port",
                                "sample": "first %d points of the step-0 cloud, single thread, %.1f s" % (n_s, dt)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
