#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_call4.log; cat gpurun_out/r2_pytest_gpu_call4.log
tools/ab.sh r2d_c5 --workload c5 --steps 20 --warmup 3
tools/ab.sh r2d_c2 --workload c2 --steps 200 --warmup 3
unset SABC_B200_LIB
for wl in c3 c4; do python bench.py --workload $wl --steps 50 --no-cpu-baseline --e2e-steps 5 > gpurun_out/r2d_$wl.json 2>gpurun_out/r2d_$wl.err; done
python - <<'PY'
import json
for wl in ("c3","c4"):
    d=json.load(open(f"gpurun_out/r2d_{wl}.json")); print(wl, d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"])
PY
export SABC_B200_LIB=$PWD/build/variants/libsabc_gm5.so
ncu --set full --clock-control none --import-source on -k regex:update_half -s 6 -c 1 -o gpurun_out/prof_r2d_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2d_c5.log 2>&1
tail -2 gpurun_out/ncu_r2d_c5.log
