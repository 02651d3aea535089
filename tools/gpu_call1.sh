#!/bin/bash
# GPU call: tests, default bench, launch lists (C2, C4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default.log 2>&1
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extra '' > gpurun_out/ncu_C2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_C4.csv python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --extra '' > gpurun_out/ncu_C4.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
tail -c 1500 gpurun_out/bench_default.log
