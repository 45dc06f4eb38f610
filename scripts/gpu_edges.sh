#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/pytest_edges.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_edges.log
