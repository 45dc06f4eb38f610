#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_policy.py -m gpu -x -q > gpurun_out/pytest_pol.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_pol.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_policy.py > gpurun_out/pytest_q.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/pytest_q.log
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-extras > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_q.json; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print('e2e', d['e2e']['value'], 'value', d['value'])"
