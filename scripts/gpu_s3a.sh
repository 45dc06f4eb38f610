#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_edges.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_s3a.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_s3a.log
timeout 300 python scripts/profile_ppo_update.py > gpurun_out/prof_ppo_s3a.log 2>&1; echo "prof rc=$?"; grep -v "^-" gpurun_out/prof_ppo_s3a.log | cut -c1-72,170-260 | head -12
timeout 300 python scripts/host_overhead.py
