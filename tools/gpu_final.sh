#!/bin/bash
# round-2 closing run on ONE GPU: ncu --set full of the hull's stream kernel, hull timings, then the full validation
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/hull5.py <<'PY'
import numpy as np, hvb200
xs = np.random.default_rng(0).random((50000, 5))
cv = hvb200.ConvexHull(xs)
print(len(cv), cv.stats["raycasts"], cv.stats["ms_search"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wrap_scan -s 12 -c 1 -f -o gpurun_out/ncu_wrap_d5 python /tmp/hull5.py > gpurun_out/ncu_wrap_d5.log 2>&1
echo "ncu rc=$?"
python - <<'PY' > gpurun_out/hull_times.log 2>&1
import time, numpy as np, hvb200
for d, n in ((3, 100000), (4, 30000), (5, 50000), (2, 1000000), (6, 20000)):
    xs = np.random.default_rng(0).random((n, d))
    s = hvb200.Raycast(xs, domain=hvb200.Boundary())
    for rep in range(3):
        t = time.perf_counter(); cv = hvb200.ConvexHull(xs, searcher=s); dt = time.perf_counter() - t
    st = cv.stats
    s.close()
    print("d=%d n=%d facets %d queries %d rounds %d ms_search %.3f ms_finalize %.3f wall %.2f ms cand32 %d cand64 %d closed %d dup %d launches %d" % (d, n, len(cv), st["raycasts"], st["rounds"], st["ms_search"], st["ms_finalize"], dt * 1e3, st["candidates_fp32"], st["candidates_fp64"], st["closed_skips"], st["duplicate_hits"], st["kernel_launches"]), flush=True)
# a lattice resolved from non-general position: timing of the two searches + merge
import time
m = 40
g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) / m
s = hvb200.Raycast(g, domain=hvb200.cuboid(3, periodic=[]))
for rep in range(2):
    t = time.perf_counter(); mesh, _ = hvb200.voronoi(g, searcher=s); dt = time.perf_counter() - t
    print("lattice %d^3: %d vertices (%d with 8 generators), wall %.1f ms, stats %s" % (m, mesh.number_of_vertices(), int((np.diff(mesh.sig_off) == 8).sum()), dt * 1e3, {k: s.stats()[k] for k in ("ms_search", "ms_finalize", "raycasts", "degenerate")}), flush=True)
PY
cat gpurun_out/hull_times.log
bash tools/gpu_validate.sh
