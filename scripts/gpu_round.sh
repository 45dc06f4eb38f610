#!/bin/bash
# One GPU session: tests, microbench, bench, ncu launch list + full captures (each only after the plain run exits 0).
set -u
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
python scripts/microbench.py > gpurun_out/mb_$TAG.log 2>&1; echo "mb rc=$?"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; echo "ref rc=$?"
if [ "${NCU:-1}" = "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 64 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 4 --warmup 3 --dephase 64 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:movegen_kernel -s 2 -c 1 -o gpurun_out/prof_k1_$TAG -f \
  python scripts/microbench.py > gpurun_out/ncu_k1_$TAG.log 2>&1
fi
tail -3 gpurun_out/pytest_$TAG.log; cat gpurun_out/mb_$TAG.log; cat gpurun_out/bench_$TAG.json
if [ "${NCU:-1}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:ppo_loss_grad -s 5 -c 1 -o gpurun_out/prof_loss_$TAG -f \
  python scripts/profile_ppo_update.py > gpurun_out/ncu_loss_$TAG.log 2>&1
fi
