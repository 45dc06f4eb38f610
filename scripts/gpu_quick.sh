#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
for t1 in 128 256; do echo "== BG_TEAM_MID=$t1"; BG_TEAM_MID=$t1 python scripts/microbench.py 2>&1 | grep -E "K1|update|K3"; done
python bench.py --steps 300 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; cat gpurun_out/bench_q.json
