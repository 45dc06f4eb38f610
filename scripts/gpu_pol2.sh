#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_ppo.py tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/pytest_pol.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_pol.log
python scripts/microbench.py 2>&1 | grep -E "N1"
timeout 300 python scripts/train_ppo.py --games 16384 --horizon 64 --updates 4 --eval-every 4 --eval-games 2048 2>&1 | tail -2 | cut -c1-500
