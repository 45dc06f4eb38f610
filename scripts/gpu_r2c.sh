#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 300 python -m pytest tests/test_gpu_mlp_twoply.py -x -q -k "twoply" > gpurun_out/pytest_$TAG.log 2>&1; RC=$?; echo "pytest rc=$RC"
tail -5 gpurun_out/pytest_$TAG.log
if [ $RC -ne 0 ]; then exit 0; fi
timeout 200 python scripts/microbench_twoply.py > gpurun_out/mb2_$TAG.log 2>&1; echo "mb rc=$?"; tail -3 gpurun_out/mb2_$TAG.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:twoply_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f \
  python scripts/microbench_twoply.py > gpurun_out/ncu_fused_$TAG.log 2>&1; echo "ncu rc=$?"
