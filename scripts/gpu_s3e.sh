#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/san_mem_s3.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_mem_s3.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/san_race_s3.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_race_s3.log
timeout 900 python tests/fuzz_k1.py --positions 30000 > gpurun_out/fuzz_k1_s3.json 2> gpurun_out/fuzz_k1_s3.err; echo "fuzz rc=$?"; cat gpurun_out/fuzz_k1_s3.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
