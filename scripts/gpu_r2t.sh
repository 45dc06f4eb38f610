#!/bin/bash
# round 2: the whole GPU suite + bench (both arms) + smoke
set -u
mkdir -p gpurun_out
TAG=${1:-r2t}; export TAG
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_'+__import__('sys').argv[1] if False else 'gpurun_out/bench_TAG.json'.replace('TAG', __import__('os').environ.get('TAG','r2t'))).read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
for k,v in d['extra'].items(): print(k, json.dumps(v)[:300])
PY
