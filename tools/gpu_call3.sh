#!/bin/bash
# kernel experiment: pooled query with static scheduling (persistent=4) against the default (persistent=3)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( HVB_PERSISTENT=4 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or matches_oracle or c1_seeds or fp32_filter or every_tile or qhull" ) > gpurun_out/pytest_mode4.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_mode4.log
tail -3 gpurun_out/pytest_mode4.log
for W in C2 C4s C3 D4; do
  for P in 3 4; do
    echo "== $W persistent=$P" >> gpurun_out/sweep_mode4.log
    timeout 300 python bench.py --workload $W --steps 5 --warmup 2 --no-cpu-baseline --extra '' --setting persistent=$P 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); s = j['stats_last_step']
        print('value=%.4g kernel_ms=%.3f step_ms=%.3f raycasts=%d dups=%d cand32=%d rows=%d' % (j['value'], j['roofline']['kernel_ms_per_step'], j['ms_per_step'], s['raycasts'], s['duplicate_hits'], s['candidates_fp32'], s['rows_scanned']))
    else: print(l[-300:])
" >> gpurun_out/sweep_mode4.log 2>&1
  done
done
cat gpurun_out/sweep_mode4.log
