#!/bin/bash
# closing check of round 2 (90 s of GPU budget left): the regression of the far-vertex fix and a parity subset on the device,
# then the bench (the fix touches the stage head of the walk kernel's query), then smoke()
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 45 python -m pytest tests/test_gpu_parity.py -q -x -k "1e7 or golden or matches_oracle or c1_seeds or qhull_directly or fp32_filter" ) > gpurun_out/pytest_last.log 2>&1
tail -4 gpurun_out/pytest_last.log
( time timeout 30 python bench.py --steps 10 --warmup 3 --no-products --no-cpu-baseline ) > gpurun_out/bench_last.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_last.log").readline())
    print("C2 value %.4g ms_per_step %.3f kernel %.3f | C4 kernel %.2f | C3 kernel %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"],
          d["workloads"]["C4"]["roofline"]["kernel_ms_per_step"], d["workloads"]["C3"]["roofline"]["kernel_ms_per_step"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1; tail -1 gpurun_out/smoke_last.log
