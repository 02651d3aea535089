#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_convexhull.py -q -x ) > gpurun_out/pytest_hull.log 2>&1
tail -5 gpurun_out/pytest_hull.log
python - <<'PY' > gpurun_out/hull_times.log 2>&1
import time, numpy as np, hvb200
for d, n in ((3, 100000), (4, 30000), (5, 50000), (2, 1000000), (6, 20000)):
    xs = np.random.default_rng(0).random((n, d))
    t = time.perf_counter(); s = hvb200.Raycast(xs, domain=hvb200.Boundary()); t_create = time.perf_counter() - t
    for rep in range(3):
        t = time.perf_counter(); cv = hvb200.ConvexHull(xs, searcher=s); dt = time.perf_counter() - t
    st = cv.stats
    t = time.perf_counter(); s.close(); t_close = time.perf_counter() - t
    print("d=%d n=%d facets %d queries %d rounds %d ms_search %.3f ms_finalize %.3f wall %.2f ms (create %.1f close %.1f) cand32 %d cand64 %d closed %d dup %d launches %d" % (d, n, len(cv), st["raycasts"], st["rounds"], st["ms_search"], st["ms_finalize"], dt * 1e3, t_create * 1e3, t_close * 1e3, st["candidates_fp32"], st["candidates_fp64"], st["closed_skips"], st["duplicate_hits"], st["kernel_launches"]), flush=True)
PY
cat gpurun_out/hull_times.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/hull_launches.csv python - <<'PY' > /dev/null 2>&1
import numpy as np, hvb200
xs = np.random.default_rng(0).random((50000, 5))
cv = hvb200.ConvexHull(xs)
PY
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/hull_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][-40:]; v = float(r[vi].replace(",", "")); 
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
for k, a in agg.items(): print("%-42s n=%4d total %10.1f us  max %9.1f us" % (k, a[0], a[1], a[2]))
PY
