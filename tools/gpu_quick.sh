#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -x -k "golden or matches_oracle or multi or slab or known_vertices or reuse" ) > gpurun_out/pytest_quick.log 2>&1
tail -4 gpurun_out/pytest_quick.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/bench_quick.log").read().splitlines() if l.startswith("{")][-1])
def show(name,x):
    s=x.get("stats_last_step",{})
    print(name, "value %.4g ms/step %.3f e2e %.3f ms kern %.3f"%(x["value"],x["ms_per_step"],x["e2e"].get("ms_per_step",0),x["roofline"]["kernel_ms_per_step"]), {k:round(v,3) for k,v in s.items() if k.startswith("ms_")})
show("C2",j)
for k,v in j.get("workloads",{}).items(): show(k,v)
PY
