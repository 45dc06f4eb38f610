#!/bin/bash
# full GPU suite + bench on HEAD (what the driver runs at round end)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_r2ae.json 2> gpurun_out/bench_r2ae.err; tail -c 300 gpurun_out/bench_r2ae.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2ae.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernels_us"])
for k, v in d.get("extra", {}).items():
    print(k, json.dumps(v)[:330])
PY
