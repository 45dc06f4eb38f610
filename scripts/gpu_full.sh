#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_full.log
timeout 600 python bench.py --steps 500 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_j.err; python -c "
import json; d=json.load(open('gpurun_out/bench_j.json')); print(d['value'], d['e2e']['value'], json.dumps(d['roofline'])[:600]); print(json.dumps(d['extra'])[:1500])"
