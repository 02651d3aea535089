#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export HVB_BENCH_ALLRANKS=1
run() { name=$1; N=$2; shift 2
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+RANDOM%100)) bench.py --gpus $N --steps 20 --warmup 5 "$@" ) > gpurun_out/scale3_$name.log 2>&1
  echo "$name rc=$?"; }
run N8_blocks 8
run N8_slabs 8 --setting decomposition=0 --no-parity
run N4_blocks 4 --no-parity
