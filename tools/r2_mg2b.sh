#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_2gpu.log; cat gpurun_out/r2_pytest_gpu_2gpu.log
python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c4_n1.json")); print("c4 n1 value %.4g ms/step %.4f e2e %.4g e2e ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"])
PY
