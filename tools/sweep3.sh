#!/bin/bash
# usage: tools/sweep3.sh : register cap / chunk width builds and seed stride on the default walk kernel (k_walk_coop)
mkdir -p gpurun_out
for W in C2 C4s; do
  for lib in base mb3 mb5 u8; do
    if [ "$lib" != "base" ]; then export HVB_LIB=$PWD/highvoronoi.jl_b200/lib/libhvb200_$lib.so; else unset HVB_LIB; fi
    f=gpurun_out/sw3_${W}_${lib}.log
    timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $f 2>&1
    python tools/showline.py "$W lib=$lib" < $f 2>&1 | tail -1
  done
done
unset HVB_LIB
for ss in 4 6 12 16; do
  f=gpurun_out/sw3_C2_ss$ss.log
  timeout 300 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --setting seed_stride=$ss > $f 2>&1
  python tools/showline.py "C2 seed_stride=$ss" < $f 2>&1 | tail -1
done
