#!/bin/bash
# round 2, GPU call 1: full GPU suite on the product (incl. the new full-size parity tests), then the prepared experiments
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc
python -m pytest tests -m gpu -q -x --durations=12 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu_call1.log; cat gpurun_out/r2_pytest_gpu_call1.log
for lib in build/variants/libsabc_ACCEPT_FILTER.so build/variants/libsabc_NODE2.so build/variants/libsabc_RK_ALL.so build/variants/libsabc_ptrs2.so; do
  echo "== parity suite on $lib"
  SABC_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_golden.py -m gpu -x -q -k "not full_size and not posterior" 2>&1 | tail -2
done
echo "== ptrs2 decision check"
SABC_B200_LIB=$PWD/build/variants/libsabc_ptrs2.so timeout 600 python tools/exp_ptrs2_check.py 3e9
echo "== A/B c4"
tools/ab.sh r2_c4 --steps 200 --warmup 3
echo "== A/B c5"
tools/ab.sh r2_c5 --workload c5 --steps 20 --warmup 3
echo "== A/B c2"
tools/ab.sh r2_c2 --workload c2 --steps 200 --warmup 3
