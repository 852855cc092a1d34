#!/bin/bash
# round 2, GPU call 2: ziggurat normal in the spec -- parity suite, then occupancy variants of the fused Gaussian kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_call2.log; cat gpurun_out/r2_pytest_gpu_call2.log
tools/ab.sh r2b_c5 --workload c5 --steps 20 --warmup 3
tools/ab.sh r2b_c2 --workload c2 --steps 200 --warmup 3
tools/ab.sh r2b_c1 --workload c1 --steps 200 --warmup 3
unset SABC_B200_LIB
python bench.py --workload c3 --steps 50 --no-cpu-baseline --e2e-steps 5 > gpurun_out/r2b_c3.json 2>gpurun_out/r2b_c3.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2b_c3.json")); print("c3", d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"])
PY
