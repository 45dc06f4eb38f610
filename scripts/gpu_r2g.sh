#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 300 python -m pytest tests/test_gpu_ppo_gemm.py -x -q > gpurun_out/pytest_$TAG.log 2>&1; RC=$?; echo "pytest rc=$RC"; tail -5 gpurun_out/pytest_$TAG.log
if [ $RC -ne 0 ]; then tail -40 gpurun_out/pytest_$TAG.log; exit 0; fi
timeout 300 python scripts/profile_ppo_update.py > gpurun_out/ppo_prof_$TAG.log 2>&1; echo "prof rc=$?"
grep -v "^---" gpurun_out/ppo_prof_$TAG.log | cut -c1-200 | grep -i "update_impl\|kernel\|Name\|memcpy\|elementwise\|gemm" | head -40
