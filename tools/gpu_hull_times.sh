#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
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
PY
cat gpurun_out/hull_times.log
python -m pytest tests/test_gpu_convexhull.py -q -x -k "matches_qhull or reproducible" 2>&1 | tail -2
