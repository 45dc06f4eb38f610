#!/bin/bash
# microbench + ncu full captures of K1 tier 0, K3 bf16, K4, policy kernel (each after the plain run exits 0)
mkdir -p gpurun_out
python scripts/microbench.py > gpurun_out/mb_h.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_h.log
for k in movegen_kernel encode_bf16_kernel mlp_value_kernel policy_kernel movegen_team_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_${k}_h -f python scripts/microbench.py > gpurun_out/ncu_${k}_h.log 2>&1; echo "ncu $k rc=$?"
done
