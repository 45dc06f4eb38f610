#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 300 python -m pytest tests/test_gpu_ppo_gemm.py -x -q > gpurun_out/pytest_$TAG.log 2>&1; RC=$?; echo "pytest rc=$RC"; tail -5 gpurun_out/pytest_$TAG.log
if [ $RC -ne 0 ]; then tail -40 gpurun_out/pytest_$TAG.log; exit 0; fi
timeout 200 python scripts/exp_ppo_gemm.py > gpurun_out/exp_gemm_$TAG.log 2>&1; head -7 gpurun_out/exp_gemm_$TAG.log; tail -3 gpurun_out/exp_gemm_$TAG.log
timeout 300 python scripts/profile_ppo_update.py > gpurun_out/ppo_prof_$TAG.log 2>&1; echo "prof rc=$?"
grep "update_impl\|ppo_gemm" gpurun_out/ppo_prof_$TAG.log | cut -c1-200
