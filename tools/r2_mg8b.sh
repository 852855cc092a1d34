#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8" 2>&1 | tail -60 > gpurun_out/r2_pytest_multi8.log; tail -5 gpurun_out/r2_pytest_multi8.log
grep -n "Error\|assert\|REPORT" gpurun_out/multi_worker_world8.log | head -20
