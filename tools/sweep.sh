#!/bin/bash
# usage: tools/sweep.sh <tag> <workload> ; prints value / e2e / kernel ms for lib variants x tile sizes
W=${2:-C2}
for lib in ${LIBS:-base mb4 mb5 mb6}; do
  for g in ${TILES:-1 2 4 8}; do
    if [ "$lib" != "base" ]; then export HVB_LIB=$PWD/highvoronoi.jl_b200/lib/libhvb200_$lib.so; else unset HVB_LIB; fi
    timeout 300 python bench.py --workload $W --steps ${STEPS:-5} --warmup 2 --no-cpu-baseline --setting tile_size=$g > gpurun_out/sweep_$1_${lib}_g$g.log 2>&1
    tail -1 gpurun_out/sweep_$1_${lib}_g$g.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib g=$g value=%.3e e2e=%.3e kernel_ms=%.3f step_ms=%.3f e2e_ms=%.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['e2e']['ms_per_step']), d['e2e']['phase_ms'])" 2>&1 | tail -1
  done
done
