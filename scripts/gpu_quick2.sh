#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
python scripts/microbench.py 2>&1 | grep -E "K1|update|K3"
python scripts/microbench_twoply.py
ncu --set full --clock-control none --import-source on -k regex:movegen_team_kernel -s 300 -c 1 -o gpurun_out/prof_team_mid_i -f python scripts/microbench.py > gpurun_out/ncu_team_i.log 2>&1; echo "ncu rc=$?"
