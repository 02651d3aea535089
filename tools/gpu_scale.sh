#!/bin/bash
# scaling run on one 8-GPU box: N = 1, 2, 4, 8 back to back (the driver's SCALE protocol) + multi-GPU tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    ( time python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/scale_N$N.log 2>&1
  else
    ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/scale_N$N.log 2>&1
  fi
  echo "N=$N rc=$?"
done
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q ) > gpurun_out/pytest_multi.log 2>&1
tail -3 gpurun_out/pytest_multi.log
