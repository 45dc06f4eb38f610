#!/bin/bash
# PPO update kernels: parity, then the update's profile (1 M samples x 4 epochs) and the per-update times at the bench's size
timeout 600 python -m pytest tests/test_gpu_ppo_gemm.py tests/test_gpu_ppo.py -x -q 2>&1 | tail -3
timeout 300 python scripts/profile_ppo_update.py > gpurun_out/ppo_profile.log 2>&1; grep "ms per update" gpurun_out/ppo_profile.log
grep -A9 "update_impl=tcgen05, " gpurun_out/ppo_profile.log | cut -c1-62,150-225 | sed -n 4,10p
timeout 300 python scripts/exp_ppo_update_times.py 2>&1 | grep "^update" | tail -3
