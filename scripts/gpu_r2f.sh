#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 300 python -m pytest tests/test_gpu_ppo_gemm.py -x -q > gpurun_out/pytest_$TAG.log 2>&1; RC=$?; echo "pytest rc=$RC"
tail -40 gpurun_out/pytest_$TAG.log
