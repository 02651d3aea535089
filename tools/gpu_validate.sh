#!/bin/bash
# full GPU test suite + default bench (+ reference arm) on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_N1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
tail -2 gpurun_out/smoke.log
( timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_check.py r2 ) > gpurun_out/sanitize.log 2>&1
echo "sanitize rc=$?" >> gpurun_out/sanitize.log
tail -4 gpurun_out/sanitize.log
