#!/bin/bash
# tests, then K1 team-size sweep and the microbench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_g.log
for t1 in 128 256 512; do for t2 in 512 1024; do
  echo "== BG_TEAM_MID=$t1 BG_TEAM_BIG=$t2"; BG_TEAM_MID=$t1 BG_TEAM_BIG=$t2 python scripts/microbench.py 2>&1 | grep -E "K1|update"
done; done
python scripts/microbench.py
python bench.py --steps 300 --no-extras --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; cat gpurun_out/bench_g.json | cut -c1-400
python bench.py --steps 300 --no-extras --no-cpu-baseline --no-overlap | cut -c1-300
