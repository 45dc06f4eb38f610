#!/bin/bash
# ncu evidence at HEAD: the launch list of a short bench run, and `ncu --set full` of K1 tier 0, the policy kernel and the PPO GEMMs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 3 --warmup 3 --min-seconds 0 --no-extras --no-cpu-baseline --pyref-seconds 0 --e2e-steps 4 --e2e-segments 1 > gpurun_out/ncu_launches.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:movegen_kernel -s 150 -c 1 -o gpurun_out/prof_k1_r2 -f python scripts/exp_k1_variants.py > gpurun_out/ncu_k1_r2.log 2>&1
if [ "$1" == "all" ]; then
timeout 500 ncu --set full --import-source on --clock-control none -k regex:policy_kernel -s 3 -c 1 -o gpurun_out/prof_policy_r2 -f python scripts/exp_policy_time.py > gpurun_out/ncu_policy_r2.log 2>&1
GAMES=16384 timeout 500 ncu --set full --import-source on --clock-control none -k regex:ppo_gemm_nt_kernel -s 12 -c 6 -o gpurun_out/prof_ppo_nt_r2 -f python scripts/profile_ppo_update.py > gpurun_out/ncu_ppo_r2.log 2>&1
fi
ls -la gpurun_out/launches_r2_final.csv gpurun_out/prof_k1_r2.ncu-rep
