#!/bin/bash
# Round-2 starting point for the Gaussian (C1/C2/C5) kernels: three compile-time experiments that are NOT in the product build
# and have never run on a GPU (DEVELOPMENT.md): -DSABC_EXPERIMENTAL_ACCEPT_FILTER (MUFU decision of log U < L),
# -DSABC_EXPERIMENTAL_NODE2 (ECDF index node searched in two dependent rounds), -DSABC_EXPERIMENTAL_RK_ALL (Philox round keys
# from the parameter block in the fused and propose kernels).
#   here:        tools/exp_gauss.sh build
#   on the box:  gpurun -- 'tools/exp_gauss.sh run'
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p build/variants
  for v in ACCEPT_FILTER NODE2 RK_ALL; do
    SABC_LIB_OUT=$PWD/build/variants/libsabc_$v.so SABC_EXTRA_NVCC_FLAGS="-DSABC_EXPERIMENTAL_$v" python simulatedannealingabc.jl_b200/build.py --force &
  done
  wait
  rm -rf build/variants/*.obj
else
  for lib in build/variants/*.so; do
    echo "== parity suite on $lib"
    SABC_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_engine.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
  done
  tools/ab.sh gauss_c5 --workload c5 --steps 20 --warmup 3
  tools/ab.sh gauss_c2 --workload c2 --steps 200 --warmup 3
fi
