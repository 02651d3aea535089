#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_convexhull.py -q ) > gpurun_out/pytest_hull.log 2>&1
tail -25 gpurun_out/pytest_hull.log
python - <<'PY' > gpurun_out/hull_times.log 2>&1
import time, numpy as np, hvb200
for d, n in ((3, 100000), (4, 30000), (5, 50000), (2, 1000000)):
    xs = np.random.default_rng(0).random((n, d))
    for rep in range(2):
        t = time.perf_counter(); cv = hvb200.ConvexHull(xs); dt = time.perf_counter() - t
    st = cv.stats
    print("d=%d n=%d facets %d raycasts %d rounds %d ms_search %.3f kernel %.3f wall %.1f ms cand32 %d rows %d" % (d, n, len(cv), st["raycasts"], st["rounds"], st["ms_search"], st["ms_expand_kernel"], dt * 1e3, st["candidates_fp32"], st["rows_scanned"]), flush=True)
PY
cat gpurun_out/hull_times.log
