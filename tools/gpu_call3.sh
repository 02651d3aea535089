#!/bin/bash
# kernel experiment: pooled query with static scheduling (persistent=4) against the default (persistent=3)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
true
true
true
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
