#!/bin/bash
# K1 with tier 2 beside tier 1: parity + timings
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fuzz_large.py tests/test_gpu_fullsize.py tests/test_gpu_team_race.py tests/test_gpu_mlp_twoply.py tests/test_gpu_policy.py -x -q 2>&1 | tail -5
TAG=tier2-beside-tier1 timeout 200 python scripts/exp_k1_variants.py 2>&1 | head -1
timeout 200 python scripts/microbench_twoply.py 2>&1 | tail -7 | head -3
