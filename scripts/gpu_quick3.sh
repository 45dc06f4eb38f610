#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_q.log
python scripts/microbench.py 2>&1 | grep -E "K1|update|K3|K4|N1"
python scripts/microbench_twoply.py
