#!/bin/bash
# GPU call (2 GPUs): full GPU test suite, bench at N=1 and N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q --durations=25 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_N1.log 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_N2.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
tail -c 600 gpurun_out/bench_N2.log
