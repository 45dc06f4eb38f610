#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/fuzz_k1.py --positions 30000 > gpurun_out/fuzz_k1.json 2> gpurun_out/fuzz_k1.err; echo "fuzz rc=$?"; cat gpurun_out/fuzz_k1.json; tail -3 gpurun_out/fuzz_k1.err
timeout 300 python bench.py --games 131072 --steps 100 --no-cpu-baseline --rows-per-game 48 > gpurun_out/bench_131k.json 2> gpurun_out/bench_131k.err; echo "bench131k rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_131k.json')); print('131072 games: value',d['value'],'greedy',d['extra']['greedy_1ply'], 'twoply', d['extra']['twoply']['root_afterstates_per_s'])"
