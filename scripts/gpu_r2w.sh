#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2w}; export TAG
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_single_env.py -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - <<'PY'
import json,os
d=json.loads(open('gpurun_out/bench_%s.json'%os.environ['TAG']).read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'rows', d['config']['legal_plays_per_step_mean'], 'launches', d['gpu_launches'])
print(d['roofline']['kernels_us'], d['roofline']['frac'], d['roofline']['k1_ms_per_launch'])
PY
