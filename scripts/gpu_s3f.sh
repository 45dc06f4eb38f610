#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_small.py 2>&1 | tail -2
timeout 900 python scripts/train_ppo.py --games 16384 --horizon 64 --updates 400 --eval-every 50 --eval-games 4096 > gpurun_out/train_ppo_long.log 2>&1; echo "train rc=$?"
grep win_rate gpurun_out/train_ppo_long.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['update'], round(d['value_loss'],3), round(d['entropy'],3), d['episodes'], 'win', round(d['win_rate'],3), 'pts', round(d['mean_points'],3), 'steps/s %.3g' % d['env_steps_per_s'])"
