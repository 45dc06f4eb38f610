#!/bin/bash
# ncu --set full of the kernels changed in round 2 (HEAD): policy_kernel, the PPO GEMMs (fused logits/loss among them), K1 tier 0
timeout 500 ncu --set full --import-source on --clock-control none -k regex:policy_kernel -s 3 -c 1 -o gpurun_out/prof_policy_r2 -f python scripts/exp_policy_time.py > gpurun_out/ncu_policy_r2.log 2>&1
GAMES=16384 timeout 500 ncu --set full --import-source on --clock-control none -k regex:ppo_gemm_nt_kernel -s 12 -c 6 -o gpurun_out/prof_ppo_nt_r2 -f python scripts/profile_ppo_update.py > gpurun_out/ncu_ppo_r2.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:movegen_kernel -s 150 -c 1 -o gpurun_out/prof_k1_r2 -f python scripts/exp_k1_variants.py > gpurun_out/ncu_k1_r2.log 2>&1
ls -la gpurun_out/prof_*_r2.ncu-rep
