#!/bin/bash
# ncu --set full captures of the walk kernel: args = list of "WORKLOAD:PERSISTENT"
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for spec in "$@"; do
  W=${spec%%:*}; P=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_coop -s 1 -c 1 -f -o gpurun_out/ncu_${W}_p${P} \
      python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline --extra '' --setting persistent=$P > gpurun_out/ncu_${W}_p${P}.log 2>&1
  echo "$spec rc=$?"
done
ls -la gpurun_out/*.ncu-rep
