#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_s3c.log
timeout 300 python scripts/microbench.py 2>&1 | grep -i "encode_f32\|K2\|K1 movegen_slab (3"
timeout 600 python bench.py --steps 400 --no-extras > gpurun_out/bench_s3c.json 2> gpurun_out/bench_s3c.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s3c.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['segments_ms_per_step'], d['roofline']['issue'])
PY
