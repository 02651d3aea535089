#!/bin/bash
# usage: tools/sweep2.sh <tag> : scan chunk width variants + seed stride on the persistent walk
mkdir -p gpurun_out
for W in C2 C4s; do
  for lib in base u2 u8; do
    if [ "$lib" != "base" ]; then export HVB_LIB=$PWD/highvoronoi.jl_b200/lib/libhvb200_$lib.so; else unset HVB_LIB; fi
    f=gpurun_out/sweep_$1_${W}_${lib}.log
    timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $f 2>&1
    python tools/showline.py "$W lib=$lib" < $f 2>&1 | tail -1
  done
done
unset HVB_LIB
for ss in 4 8 32; do
  f=gpurun_out/sweep_$1_C2_ss$ss.log
  timeout 300 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --setting seed_stride=$ss > $f 2>&1
  python tools/showline.py "C2 seed_stride=$ss" < $f 2>&1 | tail -1
done
