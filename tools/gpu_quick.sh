#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -k "multi or slab or int32 or golden" ) > gpurun_out/pytest_quick.log 2>&1
tail -15 gpurun_out/pytest_quick.log
