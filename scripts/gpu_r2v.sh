#!/bin/bash
# round 2: ncu launch list of the bench command (kernel shares of a step), full captures of K1 tier 0 and the NT GEMM kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2v}
timeout 300 python bench.py --steps 4 --warmup 3 --dephase 64 --no-cpu-baseline --e2e-steps 2 --e2e-segments 1 --no-extras --min-seconds 0.001 > gpurun_out/b_$TAG.json 2> gpurun_out/b_$TAG.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 4 --warmup 3 --dephase 64 --no-cpu-baseline --e2e-steps 2 --e2e-segments 1 --no-extras --min-seconds 0.001 > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:movegen_kernel -s 2 -c 1 -o gpurun_out/prof_k1_$TAG -f \
  python scripts/microbench.py > gpurun_out/ncu_k1_$TAG.log 2>&1; echo "k1 rc=$?"
tail -5 gpurun_out/launches_$TAG.csv | cut -c1-200
