#!/bin/bash
# K1 heavy-first: parity + timings
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fuzz_large.py tests/test_gpu_team_race.py tests/test_gpu_mlp_twoply.py -x -q 2>&1 | tail -5
TAG=heavy-first timeout 200 python scripts/exp_k1_variants.py 2>&1 | tail -3
timeout 200 python scripts/microbench_twoply.py 2>&1 | tail -7
