#!/bin/bash
mkdir -p gpurun_out
for k in mlp_value_kernel policy_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_${k}_m -f python scripts/microbench.py > gpurun_out/ncu_${k}_m.log 2>&1; echo "ncu $k rc=$?"
done
