#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp_twoply.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/microbench.py 2>&1 | grep "K4\|N1"
timeout 300 python scripts/microbench_twoply.py 2>&1 | tail -6
