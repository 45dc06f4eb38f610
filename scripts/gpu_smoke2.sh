#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
PPO=0 bash scripts/gpu_multi.sh 2 > /dev/null 2>&1
wc -l gpurun_out/bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('value',d['value'],'e2e',d['e2e']['value']); print(json.dumps(d.get('extra'))[:600])"
