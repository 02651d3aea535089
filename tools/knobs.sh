#!/bin/bash
# usage: tools/knobs.sh <workload> <steps> "<setting list>" ...   e.g. tools/knobs.sh C2 5 "points_per_cell=4" "points_per_cell=8 probe_scale=1.5"
W=$1; S=$2; shift 2
for cfg in "$@"; do
  args=""; for kv in $cfg; do args="$args --setting $kv"; done
  f=gpurun_out/knob_${W}_$(echo $cfg | tr ' =.' '___').log
  timeout 300 python bench.py --workload $W --steps $S --warmup 2 --no-cpu-baseline $args > $f 2>&1
  python tools/showline.py "$W [$cfg]" < $f 2>&1 | tail -1
done
