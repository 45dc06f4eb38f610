#!/bin/bash
# policy kernel with two CTAs per SM: parity, then the bench extras
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_ppo.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --no-cpu-baseline --pyref-seconds 0 > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err; tail -c 600 gpurun_out/bench_r2aa.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2aa.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in d.get("extra", {}).items():
    print(k, json.dumps(v)[:400])
PY
