#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export HVB_BENCH_ALLRANKS=1
run() { name=$1; N=$2; shift 2
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+RANDOM%100)) bench.py --gpus $N --steps 20 --warmup 5 "$@" ) > gpurun_out/scale4_$name.log 2>&1
  echo "$name rc=$?"; }
run N8 8
run N4 4
run N2 2
( time python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/scale4_N1.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_multi.py -q ) > gpurun_out/pytest_multi.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_multi.log
