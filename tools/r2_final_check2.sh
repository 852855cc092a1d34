#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu_2gpu_final.log; cat gpurun_out/r2_pytest_gpu_2gpu_final.log
