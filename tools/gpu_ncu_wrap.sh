#!/bin/bash
# ncu --set full of the convex hull's stream kernel (k_wrap_scan<5>) on the C4 cloud, one launch of a large round
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > gpurun_out/hull5.py <<'PY'
import numpy as np, hvb200
xs = np.random.default_rng(0).random((50000, 5))
cv = hvb200.ConvexHull(xs)
print(len(cv), cv.stats["raycasts"], cv.stats["ms_search"])
PY
PYTHONPATH=. timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wrap_scan -s 12 -c 1 -f -o gpurun_out/ncu_wrap_d5 python gpurun_out/hull5.py > gpurun_out/ncu_wrap_d5.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/ncu_wrap_d5.ncu-rep --page raw --csv > gpurun_out/ncu_wrap_d5_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_wrap_d5.log
