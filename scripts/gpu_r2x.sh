#!/bin/bash
# fused class-A loss epilogue: parity, then the update's profile with and without it
set -x
timeout 600 python -m pytest tests/test_gpu_ppo_gemm.py -x -q 2>&1 | tail -8
timeout 300 python scripts/profile_ppo_update.py > gpurun_out/ppo_profile_r2x.log 2>&1; tail -3 gpurun_out/ppo_profile_r2x.log
BG_PPO_FUSE_LOSS=0 timeout 300 python scripts/profile_ppo_update.py > gpurun_out/ppo_profile_r2x_unfused.log 2>&1; grep "ms per update" gpurun_out/ppo_profile_r2x_unfused.log
grep "ms per update" gpurun_out/ppo_profile_r2x.log
