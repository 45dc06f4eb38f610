#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2j}
export GAMES=16384
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppo_gemm -s 10 -c 10 -o gpurun_out/prof_ppogemm_$TAG -f \
  python scripts/profile_ppo_update.py > gpurun_out/ncu_ppogemm_$TAG.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_ppogemm_$TAG.log
