#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-k}
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e', d['e2e']['value'], d['e2e'].get('segments_ms_per_step')); print(json.dumps(d['roofline'])[:420]); print(json.dumps(d['extra'])[:1800]); print(d['cpu_baseline']['value'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --dephase 64 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:movegen_kernel -s 2 -c 1 -o gpurun_out/prof_k1_$TAG -f python scripts/microbench.py > gpurun_out/ncu_k1_$TAG.log 2>&1; echo "ncu k1 rc=$?"
