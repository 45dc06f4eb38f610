#!/bin/bash
# learning check: the same training with the hand-written update and with the library-GEMM update, two seeds each
for impl in tcgen05 cublas; do for seed in 0 1; do
  timeout 300 python scripts/train_ppo.py --games 16384 --horizon 64 --updates 200 --eval-every 100 --seed $seed --update-impl $impl > gpurun_out/train_${impl}_s$seed.log 2>&1
  python - <<PY
import json
for line in open("gpurun_out/train_${impl}_s$seed.log"):
    if '"win_rate"' in line:
        d = json.loads(line); print("$impl seed $seed update", d["update"], "win rate", round(d["win_rate"], 3), "points", round(d["mean_points"], 3), "entropy", round(d["entropy"], 3), "Msteps/s", round(d["env_steps_per_s"] / 1e6, 1))
PY
done; done
