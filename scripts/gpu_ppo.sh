#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_ppo.py -m gpu -x -q > gpurun_out/pytest_ppo.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_ppo.log
timeout 600 python scripts/train_ppo.py --games 16384 --horizon 64 --updates 6 --eval-every 3 --eval-games 4096 > gpurun_out/train_ppo.log 2>&1; echo "train rc=$?"; tail -8 gpurun_out/train_ppo.log | cut -c1-600
