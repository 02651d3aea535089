#!/usr/bin/env python
"""bench.py -- Voronoi vertices / second of the raycast vertex search (BASELINE.json metric, "d = 3,5 at 1/2/4/8 B200").

One "step" = one pass of the hot path over one synthetic point cloud: voronoi(xs; searcher=Raycast(xs; domain))
through the C ABI of libhvb200.so.

  headline (`value`, `e2e`, `roofline`)      : C2 = configs[1] of BASELINE.json (100 000 uniform points, d = 3, cuboid
                                               boundary, vertex search + neighbours); N > 1: weak scaling, N x 100 000 points
  `workloads` (same JSON line, fewer steps)  : C4 = configs[3] (50 000 points, d = 5; N > 1: weak, N x 50 000) and
                                               C3 = configs[2] (1 000 000 points, d = 2; N > 1: STRONG, 1 000 000 points in all)

N > 1 (one process per GPU, launched by torchrun): generators and index replicated, GPU k walks slab k of the spatially
sorted order (parallelmesh.jl:52-87) and keeps the disjoint set of vertices it OWNS -- the result stays sharded, like the
work; the only collective of a step is the library's own ncclAllGather of the shard sizes (hvb_exchange_counts).  Before
anything is timed the merged result (hvb_allgather: counts + compact rows over NCCL, inside the library) is compared with
a world = 1 search of the same cloud on rank 0 (`parity_n`); a mismatch fails the run.  torch.distributed is plumbing: it
carries the 128-byte NCCL id, the barriers and the max-over-ranks of the timings.

  value : whole-job vertices/s of the hot path on the device, result (sorted vertex rows + neighbour lists) complete in
          HBM: spatial index build (the reference times Raycast(xs) + voronoi(), statistics.jl:98-126) + hvb_search,
          device time from CUDA events on the library's stream (ms_build - ms_upload + ms_search + ms_finalize: the
          generators count as resident; their upload and the D2H of the result belong to e2e); N > 1: plus the count
          exchange, max over ranks
  e2e   : the same through the public API from HOST buffers: hvb_set_points (H2D + index build) + hvb_search + fetch of the
          vertex rows and neighbour lists into host memory (D2H; N > 1: every rank fetches its shard), wall clock
  --impl reference : the reference's own implementation on the host cores: Julia + HighVoronoi.jl when a `julia` binary and
          the package are present (oracle/run_reference.jl, protocol of statistics.jl:98-126), else the CPU restatement of
          the reference algorithm (oracle/hv_oracle.cpp, 8 threads like MultiThread(8,1))
"""
import argparse
import ctypes
import json
import os
import shutil
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
DEFAULT_PERSISTENT = 3          # hvb_default_params: the walk variant behind persistent (include/hvb200.h)
DEFAULT_EXTRA = "C4,C3,D4"     # side workloads of the default line (D4 at N = 1 only: it carries the reference's one-thread figure)
STRONG = {"C3"}                 # N > 1: the total stays fixed (BASELINE.json configs[2] names 1 000 000 points over all GPUs)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed ncu capture
    (profiles/r2_traffic.json, else r1; null when no capture exists for this workload)"""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(workload, {}).get("dram_bytes_per_launch")
            if v is not None:
                return v
        except Exception:
            pass
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


def workload_text(name, n_per_gpu, n_total, d, periodic):
    return "%s: %d uniform points per GPU (%d total), d=%d, %s, vertices + neighbours" % (
        name, n_per_gpu, n_total, d,
        "cuboid(d) all axes periodic: halo generators + certificate on the device" if periodic else "cuboid(d,periodic=[])")


# ---------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
def find_julia():
    """a Julia with HighVoronoi.jl, if this box has one (baseline/_ref may carry it); None otherwise"""
    cands = [os.path.join(ROOT, "baseline", "_ref", "bin", "julia"), os.path.join(ROOT, "baseline", "_ref", "julia", "bin", "julia"),
             shutil.which("julia")]
    for c in cands:
        if c and os.path.exists(c):
            try:
                v = subprocess.run([c, "--version"], capture_output=True, text=True, timeout=60)
                if v.returncode == 0:
                    return c, v.stdout.strip()
            except Exception:
                pass
    return None, None


def run_reference_julia(julia, version, args, n_sample, d, cores):
    """times the real package with the protocol of statistics.jl:98-126 (oracle/run_reference.jl); None if it cannot run"""
    env = dict(os.environ, JULIA_NUM_THREADS=str(cores))
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref):
        env["JULIA_LOAD_PATH"] = ref + os.pathsep + os.path.join(ref, "src") + os.pathsep + env.get("JULIA_LOAD_PATH", "@:@v#.#:@stdlib")
    try:
        p = subprocess.run([julia, os.path.join(ROOT, "oracle", "run_reference.jl"), str(d), str(n_sample), str(args.steps), str(args.warmup), str(cores)],
                           capture_output=True, text=True, timeout=1500, env=env)
        lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not lines:
            return None
        r = json.loads(lines[-1])
        return {"verts": r["vertices"], "seconds": r["seconds"], "version": version}
    except Exception:
        return None


def run_reference(args, name, n_per_gpu, d, rank, world):
    """the reference arm: the reference's own CPU implementation on the host cores, bounded sample of our arm's workload"""
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 8)              # the reference cannot use more than 8 threads (chull.jl:185-193)
    n_total = n_per_gpu * (1 if name in STRONG else world)
    n_sample = min(n_total, max(2000, int(args.ref_points)))
    julia, jver = find_julia()
    res = run_reference_julia(julia, jver, args, n_sample, d, cores) if julia else None
    if res is not None:
        kind, T, verts = "reference", res["seconds"], res["verts"]
        note = "HighVoronoi.jl under %s, JULIA_NUM_THREADS=%d, MultiThread(%d,1), protocol of statistics.jl:98-126" % (jver, cores, cores)
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import hv_oracle
        import qhull_oracle
        hv_oracle.build()
        base, normal = qhull_oracle.cuboid(d)
        times, verts = [], 0
        for it in range(args.warmup + args.steps):
            xs = cloud(n_sample, d, it)
            o = hv_oracle.run(xs, base, normal, nthreads=cores)
            # the reference's protocol (statistics.jl:98-126) times Raycast(xs) + voronoi(): index build + cell loop;
            # the oracle's sorting / neighbour post-processing for the tests is not part of it
            dt = (o["stats"]["search_us"] + o["stats"]["build_us"]) * 1e-6
            if it >= args.warmup:
                times.append(dt)
                verts += len(o["sig"])
        kind, T = "port", sum(times)
        note = "CPU restatement of the reference algorithm (oracle/hv_oracle.cpp), not Julia: no julia binary on this box (probed baseline/_ref and PATH)"
    val = verts / T
    # vertices/s of a uniform cloud does not depend on its extent: n_sample uniform points in the unit cube are, up to
    # scaling, the points of the workload that fall into a sub-box holding n_sample of them (same density per cell)
    sample = ("the whole workload, %d points" % n_sample) if n_sample == n_total else \
        "%d uniform points = a sub-box of the %d-point workload at the same density, rescaled to the unit cube" % (n_sample, n_total)
    line = {"impl": "reference", "metric": "voronoi_vertices_per_sec", "value": val, "unit": "vertices/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
            "scaling": "strong" if name in STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(name, n_per_gpu // world if name in STRONG else n_per_gpu, n_total, d, False),
                       "parallelism": "slab%d" % world, "threads": cores},
            "cpu_baseline": {"value": val, "unit": "vertices/s", "cores": cores, "kind": kind, "sample": sample, "note": note},
            "e2e": {"value": val, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Runner:
    """one workload on this rank's GPU: context, page-locked caller buffers, the timed step"""

    def __init__(self, name, env, settings):
        import torch
        import hvb200
        self.torch, self.hvb, self.env, self.name = torch, hvb200, env, name
        self.L = hvb200._abi.lib()
        self.abi = hvb200._abi
        n_per_gpu, self.d = WORKLOADS[name]
        self.world, self.rank = env["world"], env["rank"]
        self.ngpus = env.get("ngpus", 1)                 # > 1: ONE process drives the GPUs through hvb_create_multi (--single-process)
        parts = self.world * self.ngpus
        self.strong = name in STRONG and parts > 1
        self.n_total = n_per_gpu if self.strong or parts == 1 else n_per_gpu * parts
        self.n_per_gpu = self.n_total // parts if self.strong else n_per_gpu
        self.periodic = name in PERIODIC
        self.settings = settings
        self.dom = hvb200.cuboid(self.d) if self.periodic else hvb200.cuboid(self.d, periodic=[])
        self.s = None
        self.xs_pin = torch.empty((self.n_total, self.d), dtype=torch.float64, pin_memory=True)
        self.out = {}
        self.phases = np.zeros(4)

    def searcher(self, xs):
        hvb = self.hvb
        if self.s is None:                               # the context (device + page-locked buffers) is re-used
            kw = dict(neighbors=1, wire32=1)             # ids cross PCIe as int32 (hvb_view_vertices32 / hvb_view_neighbors32)
            kw.update(self.settings)
            if self.ngpus > 1:
                kw["wire32"] = 0                         # the shards of several GPUs are copied into the caller's int64 buffers
                thr = hvb.B200Thread(ngpus=self.ngpus)
            else:
                thr = hvb.B200Thread(self.env["local_rank"], self.rank, self.world)
            opts = hvb.RaycastParameter(threading=thr, **kw)
            self.s = hvb.Raycast(xs, domain=self.dom, options=opts, periodic=self.periodic)
            if self.world > 1:
                from hvb200 import multigpu
                multigpu.init_comm(self.s)
        else:
            self.s.set_points(xs)                        # H2D + index build
        return self.s

    def host_rows(self, count):
        """page-locked host buffers of the caller for this rank's shard, allocated once with headroom"""
        d, torch = self.d, self.torch
        if self.out.get("cap", 0) < count:
            cap = int(1.25 * count) + 1024
            self.out = {"cap": cap, "sig": torch.empty((cap, d + 1), dtype=torch.int64, pin_memory=True).numpy(),
                        "r": torch.empty((cap, d), dtype=torch.float64, pin_memory=True).numpy()}
        return self.out["sig"], self.out["r"]

    def step(self, it, timed):
        """returns (device_ms, e2e_s, vertices_all_ranks, stats, launches, h2d_bytes, d2h_bytes)"""
        torch, L, abi, d = self.torch, self.L, self.abi, self.d
        xs = self.xs_pin.numpy()
        xs[:] = cloud(self.n_total, d, it)               # new synthetic cloud every step (identical on every rank)
        for buf in self.env["flush"] if isinstance(self.env["flush"], list) else [self.env["flush"]]:
            buf.fill_(it & 0xff)
        self.env["barrier"]()
        t0 = time.perf_counter()
        s = self.searcher(xs)
        st0 = s.stats()                                  # index build of THIS step (hvb_create / hvb_set_points) without the upload
        build_ms = st0["ms_build"] - st0["ms_upload"]
        t1 = time.perf_counter()
        abi.check(L.hvb_search(s._ctx, None, 0, None, None, 0, 0), s._ctx)
        st = s.stats()
        dev_ms = build_ms + st["ms_search"] + st["ms_finalize"]
        if self.world > 1:
            tx = time.perf_counter()
            cnt = np.zeros(self.world, dtype=np.int64)
            abi.check(L.hvb_exchange_counts(s._ctx, cnt.ctypes.data_as(ctypes.c_void_p)), s._ctx)   # library-issued ncclAllGather + host wait
            dev_ms += 1e3 * (time.perf_counter() - tx)
            V, mine = int(cnt.sum()), int(cnt[self.rank])
        t1b = time.perf_counter()
        if self.world == 1:
            mesh = self.hvb.VoronoiMesh(s)                         # D2H of vertices (and rays), zero-copy views of the staging buffers
            sig_h, r_h = mesh.sig, mesh.r
            # periodic: vertices are counted once per image class ("excl. halo duplicates")
            V = st["unique_vertices"] if self.periodic else sig_h.shape[0]
            t1c = time.perf_counter()
            off, ids = mesh.neighbors()                            # neighbour lists + D2H
            nb_bytes = off.nbytes + ids.nbytes
        else:
            # every rank keeps its shard and the neighbour lists of its own cells
            sig_h, r_h = self.host_rows(mine)
            abi.check(L.hvb_fetch_vertices_range(s._ctx, 0, mine, sig_h.ctypes.data_as(ctypes.c_void_p), r_h.ctypes.data_as(ctypes.c_void_p)), s._ctx)
            sig_h, r_h = sig_h[:mine], r_h[:mine]
            t1c = time.perf_counter()
            po, pi, tot = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
            abi.check(L.hvb_view_neighbors32(s._ctx, ctypes.byref(po), ctypes.byref(pi), ctypes.byref(tot)), s._ctx)
            nb_bytes = (self.n_total + 1) * 8 + int(tot.value) * 4
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if timed:
            self.phases[:] += (t1 - t0, t1b - t1, t1c - t1b, t2 - t1c)
        st2 = s.stats()
        return dev_ms, t2 - t0, V, st, st2["kernel_launches"], xs.nbytes, sig_h.nbytes + r_h.nbytes + nb_bytes

    def run(self, steps, warmup):
        env, d = self.env, self.d
        for it in range(warmup):
            self.step(1000 + it, False)
        self.phases[:] = 0
        dev_ms_tot, e2e_tot, verts, launches, kern_ms, kern_launches, h2d, d2h = 0.0, 0.0, 0, 0, 0.0, 0, 0, 0
        stats_last, step_ms = None, []
        for it in range(steps):
            dm, es, V, st, nl, hb, db = self.step(it, True)
            dev_ms_tot += env["all_max"](dm)
            step_ms.append(round(dm, 3))
            e2e_tot += env["all_max"](es)
            verts += V
            launches += nl
            kern_ms += st["ms_expand_kernel"]
            kern_launches += st["expand_launches"]
            h2d, d2h, stats_last = hb, db, st
        if os.environ.get("HVB_BENCH_ALLRANKS"):
            sys.stderr.write("[rank %d] %s %s\n" % (self.rank, self.name, json.dumps({k: round(float(stats_last[k]), 3) for k in (
                "ms_build", "ms_upload", "ms_search", "ms_finalize", "ms_seed", "ms_neighbors", "ms_rows_sort", "ms_expand_kernel", "raycasts", "vertices")})))
        peak, peak_src = measured_peak()
        # roofline of the dominant kernel (the walk): algorithmic bytes per launch / average launch duration, this rank
        v_rank = stats_last["raycasts"] - stats_last["duplicate_hits"] if self.world > 1 else stats_last["vertices"]
        if self.ngpus > 1:                               # counters are summed over the GPUs, the kernel time is the slowest GPU's
            v_rank = (stats_last["raycasts"] - stats_last["duplicate_hits"]) / self.ngpus
        bytes_per_launch = B_ALG[d] * (v_rank * steps) / max(kern_launches, 1)
        avg_launch_s = kern_ms * 1e-3 / max(kern_launches, 1)
        achieved = bytes_per_launch / avg_launch_s / 1e9
        kname = {0: "k_expand<%d>", 1: "k_walk<%d>", 2: "k_walk_coop<%d,1> (pooled query, row tickets)", 3: "k_walk_coop<%d,0>",
                 4: "k_walk_coop<%d,2> (pooled query, static schedule)"}[int(os.environ.get("HVB_PERSISTENT", self.settings.get("persistent", DEFAULT_PERSISTENT)))] % d
        return {
            "value": verts / (dev_ms_tot * 1e-3), "unit": "vertices/s", "steps": steps, "warmup": warmup, "ms_per_step": dev_ms_tot / steps,
            "scaling": "strong" if self.strong else "weak",
            "config": {"workload": workload_text(self.name, self.n_per_gpu, self.n_total, d, self.periodic), "parallelism": ("one process, %d GPUs (hvb_create_multi)" % self.ngpus) if self.ngpus > 1 else "slab%d" % self.world,
                       "l2": "256 MiB L2 flush before every step; steps timed one by one and summed", "settings": self.settings,
                       "wire": "ids cross PCIe as int64, coordinates as f64" if self.ngpus > 1 else "ids cross PCIe as int32 (wire32), coordinates as f64"},
            "e2e": {"value": verts / e2e_tot, "unit": "vertices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_tot / steps,
                    "phase_ms": dict(zip(("set_points", "search", "fetch_vertices", "neighbors"), (1e3 * self.phases / steps).round(3).tolist()))},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(self.name) if self.world == 1 and self.ngpus == 1 else None, "peak_source": peak_src, "bytes_per_vertex": B_ALG[d],
                         "vertices_per_launch": v_rank * steps / max(kern_launches, 1), "launches_per_step": kern_launches / steps,
                         "kernel_ms_per_step": kern_ms / steps},
            "vertices_per_step": verts / steps,
            "stats_last_step": {k: stats_last[k] for k in ("raycasts", "duplicate_hits", "closed_skips", "candidates_fp32", "candidates_fp64", "rows_scanned", "rounds",
                                                            "seeds", "ms_build", "ms_upload", "ms_search", "ms_finalize", "ms_seed", "ms_neighbors", "ms_rows_sort",
                                                            "ms_stage_wait", "capacity_retries", "vertices", "unique_vertices", "halo_nodes", "periodic_retries",
                                                            "rejected", "suboptimal")},
            "step_ms_list": step_ms,
        }

    def close(self):
        if self.s is not None:
            self.s.close()
            self.s = None


def checksum(sig, r):
    """order-independent 128-bit checksum of vertex rows: sum and xor of a 64-bit mix of every row's ids and coordinate BITS"""
    h = np.full(sig.shape[0], 0x9e3779b97f4a7c15, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for col in list(sig.T.astype(np.uint64)) + list(np.ascontiguousarray(r).view(np.uint64).reshape(r.shape).T):
            h = (h ^ col) * np.uint64(0x100000001b3) + np.uint64(0x632be59bd9b4e019)
            h ^= h >> np.uint64(29)
        return int(h.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(h)) if len(h) else 0


def parity_n(env, name, settings):
    """N > 1, before anything is timed: slab searches + hvb_allgather on every rank against a world = 1 search of the same
    cloud on rank 0.  Row count, checksum over ids and coordinate bits (max |dr| = 0: canonical coordinates), and the
    neighbour lists rebuilt from the merged rows.  Returns a dict for the JSON line; raises on a mismatch."""
    import hvb200
    from hvb200 import multigpu
    dist, torch = env["dist"], env["torch"]
    n_per_gpu, d = WORKLOADS[name]
    n_total = n_per_gpu * env["world"]
    xs = cloud(n_total, d, 4242)
    dom = hvb200.cuboid(d, periodic=[])
    s = hvb200.Raycast(xs, domain=dom, options=hvb200.RaycastParameter(threading=hvb200.B200Thread(env["local_rank"], env["rank"], env["world"]), **settings))
    multigpu.init_comm(s)
    L, abi = hvb200._abi.lib(), hvb200._abi
    abi.check(L.hvb_search(s._ctx, None, 0, None, None, 0, 0), s._ctx)
    shard = multigpu.exchange_counts(s)
    env["barrier"]()
    t0 = time.perf_counter()
    multigpu.allgather(s)
    torch.cuda.synchronize()
    ms_gather = env["all_max"](1e3 * (time.perf_counter() - t0))
    merged = hvb200.VoronoiMesh(s, copy=True)
    cs = checksum(merged.sig, merged.r)
    off, ids = merged.neighbors()
    nb = (int(off[-1]), int(np.asarray(ids, dtype=np.uint64).sum(dtype=np.uint64)))
    xbytes = s.stats()["exchange_bytes"]
    s.close()
    ok = 1
    detail = ""
    if env["rank"] == 0:
        s1 = hvb200.Raycast(xs, domain=dom, options=hvb200.RaycastParameter(threading=hvb200.B200Thread(env["local_rank"], 0, 1), **settings))
        ref, _ = hvb200.voronoi(xs, searcher=s1)
        o1, i1 = ref.neighbors()
        ref_cs, ref_nb = checksum(ref.sig, ref.r), (int(o1[-1]), int(np.asarray(i1, dtype=np.uint64).sum(dtype=np.uint64)))
        rows_ref = ref.sig.shape[0]
        s1.close()
    else:
        ref_cs, ref_nb, rows_ref = None, None, None
    box = [(ref_cs, ref_nb, rows_ref)]
    dist.broadcast_object_list(box, src=0)
    ref_cs, ref_nb, rows_ref = box[0]
    if merged.sig.shape[0] != rows_ref or cs != ref_cs or nb != ref_nb or int(shard.sum()) != rows_ref:
        ok, detail = 0, "rank %d: rows %d vs %d, checksum %s vs %s, neighbours %s vs %s" % (env["rank"], merged.sig.shape[0], rows_ref, cs, ref_cs, nb, ref_nb)
    ok_all = int(-env["all_max"](-float(ok)))
    if not ok_all:
        raise SystemExit("parity_n FAILED: merged multi-GPU result differs from the world=1 search (%s)" % detail)
    return {"parity_n": "ok", "parity_n_detail": {"workload": name, "points": n_total, "rows": rows_ref, "checksum128": "%016x%016x" % ref_cs,
                                                  "max_abs_dr": 0.0, "shard_rows": shard.tolist(),
                                                  "allgather_ms": ms_gather, "allgather_bytes_received_per_rank": xbytes}}


def products():
    """Other products of the same backend, timed once each on a warm context (not part of `value`): the convex hull of the C4
    cloud by gift wrapping (hvb_convex_hull; compare workloads.C4.ms_per_step = the complete search the hull used to cost) and a
    40^3 lattice, a cloud in non-general position (resolved by perturbation + merge)."""
    import hvb200
    out = {}
    try:
        xs = cloud(50000, 5, 0)
        s = hvb200.Raycast(xs, domain=hvb200.Boundary())
        for _ in range(3):
            t0 = time.perf_counter(); cv = hvb200.ConvexHull(xs, searcher=s); wall = time.perf_counter() - t0
        st = cv.stats
        out["convex_hull_C4"] = {"facets": len(cv), "queries": st["raycasts"], "rounds": st["rounds"], "ms_device": st["ms_search"] + st["ms_finalize"],
                                 "ms_wall": wall * 1e3, "pairs_fp32": st["candidates_fp32"], "fp64_evaluations": st["candidates_fp64"]}
        s.close()
        m = 40
        g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
        s = hvb200.Raycast(g, domain=hvb200.cuboid(3, periodic=[]))
        for _ in range(2):
            t0 = time.perf_counter(); mesh, _s = hvb200.voronoi(g, searcher=s); wall = time.perf_counter() - t0
        st = s.stats()
        out["lattice_40^3"] = {"vertices": mesh.number_of_vertices(), "with_8_generators": int((np.diff(mesh.sig_off) == 8).sum()), "max_siglen": mesh.max_siglen,
                               "ms_wall": wall * 1e3, "ms_search": st["ms_search"], "ms_finalize": st["ms_finalize"]}
        s.close()
    except Exception as e:                                   # a product must not take the bench line down
        out["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--extra", default=DEFAULT_EXTRA, help="workloads reported under `workloads` in the same line ('' = none)")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--no-products", action="store_true", help="skip the `products` timings (convex hull, lattice)")
    ap.add_argument("--ref-points", type=float, default=100000)
    ap.add_argument("--cpu-points", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="--gpus N in ONE process without torch.distributed: the library drives the GPUs (hvb_create_multi, one host thread + "
                         "context per GPU, ncclCommInitAll) -- the model of a Julia caller (INTEGRATION.md)")
    ap.add_argument("--setting", action="append", default=[], help="backend knob, e.g. tile_size=8")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_per_gpu, d = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, args.workload, n_per_gpu, d, rank, world)
    single = args.single_process and args.gpus > 1
    if single and world > 1:
        raise SystemExit("bench.py: --single-process is not launched under torchrun")
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and not single:
        # `python bench.py --gpus N` outside torchrun: one rank per GPU is the contract, so launch the ranks here instead of
        # printing an N-GPU line measured on one GPU
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if args.gpus != world and not single:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))

    import torch
    import hvb200  # noqa: F401
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    settings = {}
    for kv in args.setting:
        k, v = kv.split("=")
        settings[k] = float(v) if "." in v else int(v)
    if args.workload in PERIODIC and world > 1 and not settings.get("periodic_margin"):
        pass                                              # periodic contexts agree on the margin through the library's communicator

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

    env = {"rank": rank, "local_rank": local_rank, "world": world, "dist": dist, "torch": torch, "barrier": barrier, "all_max": all_max,
           "flush": torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")}      # > 126 MB L2
    if single:
        env["ngpus"] = args.gpus
        env["flush"] = [torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:%d" % k) for k in range(args.gpus)]

        def barrier():                                    # noqa: F811  (every device, not only the current one)
            for k in range(args.gpus):
                torch.cuda.synchronize(k)
        env["barrier"] = barrier

    par = {}
    if world > 1 and not args.no_parity and args.workload not in PERIODIC:
        par = parity_n(env, args.workload, settings)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    main_run = Runner(args.workload, env, settings)
    res = main_run.run(args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    main_run.close()
    extras, extra_errors = {}, {}
    for name in [w for w in args.extra.split(",") if w and w != args.workload]:
        if name == "D4" and world > 1 and args.extra == DEFAULT_EXTRA:
            continue                                      # the published number is a one-thread figure: reported at N = 1 only
        try:
            r = Runner(name, env, settings)
            extras[name] = r.run(max(1, min(args.extra_steps, args.steps)), 1)
            r.close()
        except Exception as e:                            # noqa: BLE001
            if world > 1:
                raise                                     # ranks must not part ways in front of a collective
            extra_errors[name] = repr(e)                  # a side workload must not take the headline line down
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    line = {"metric": "voronoi_vertices_per_sec", "value": res["value"], "unit": "vertices/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": res["config"], "e2e": res["e2e"],
            "gpu_launches": res["gpu_launches"] + sum(x["gpu_launches"] for x in extras.values()), "clocks": clocks, "roofline": res["roofline"],
            "vertices_per_step": res["vertices_per_step"], "stats_last_step": res["stats_last_step"], "step_ms_list": res["step_ms_list"]}
    line.update(par)
    if extras:
        line["workloads"] = {k: {kk: v[kk] for kk in ("value", "unit", "steps", "warmup", "ms_per_step", "scaling", "config", "e2e", "roofline", "vertices_per_step",
                                                      "stats_last_step")} for k, v in extras.items()}
    if extra_errors:
        line["workload_errors"] = extra_errors
    if "D4" in extras and world == 1:
        # the one workload of this bench the reference publishes a number for (docs/src/index.md:93): 30 000
        # uniform points in the unit cube, d = 4 -- 841 395.0 vertices in 14.37 s on one thread of the author's PC
        line["workloads"]["D4"]["published_by_the_reference"] = {
            "vertices": 841395.0, "seconds": 14.368660125, "vertices_per_s": 841395.0 / 14.368660125, "hardware": "author's PC, 1 thread",
            "source": "docs/src/index.md:93", "vertices_here": extras["D4"]["vertices_per_step"],
            "value_over_published": extras["D4"]["value"] / (841395.0 / 14.368660125),
            "e2e_over_published": extras["D4"]["e2e"]["value"] / (841395.0 / 14.368660125)}
    if world == 1 and not args.no_products:
        line["products"] = products()
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
