#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fullsize.py tests/test_gpu_mlp_twoply.py tests/test_gpu_ppo.py -m gpu -x -q > gpurun_out/pytest_s3b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_s3b.log
timeout 300 python scripts/profile_ppo_update.py > gpurun_out/prof_ppo_s3b.log 2>&1; echo "prof rc=$?"; grep -v "^-" gpurun_out/prof_ppo_s3b.log | cut -c1-72,170-260 | head -8; tail -3 gpurun_out/prof_ppo_s3b.log
timeout 300 python scripts/host_overhead.py
timeout 600 python bench.py --steps 400 > gpurun_out/bench_s3b.json 2> gpurun_out/bench_s3b.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s3b.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['segments_ms_per_step'], d['roofline']['k1_ms_per_launch'], d['roofline']['k1_ms_per_launch_alone'])
print({k:(v.get('root_afterstates_per_s') or v.get('env_steps_per_s') or v.get('positions_per_s')) for k,v in d['extra'].items()})
PY
