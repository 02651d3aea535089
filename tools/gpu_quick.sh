#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -k "multi or slab" ) > gpurun_out/pytest_quick.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_quick.log
