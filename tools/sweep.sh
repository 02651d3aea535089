#!/bin/bash
# usage: [LIBS=..] [TILES=..] [STEPS=..] tools/sweep.sh <tag> <workload>
W=${2:-C2}
for lib in ${LIBS:-base mb4}; do
  for g in ${TILES:-1 2 4 8}; do
    if [ "$lib" != "base" ]; then export HVB_LIB=$PWD/highvoronoi.jl_b200/lib/libhvb200_$lib.so; else unset HVB_LIB; fi
    f=gpurun_out/sweep_$1_${W}_${lib}_g$g.log
    timeout 300 python bench.py --workload $W --steps ${STEPS:-5} --warmup 2 --no-cpu-baseline --setting tile_size=$g $EXTRA > $f 2>&1
    python tools/showline.py "$W lib=$lib g=$g" < $f 2>&1 | tail -1
  done
done
