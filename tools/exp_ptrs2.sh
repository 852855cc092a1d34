#!/bin/bash
# Round-2 starting point for the SIR kernel (DEVELOPMENT.md): builds the experimental PTRS attempt as a variant library
# (HERE, before gpurun) and, on the GPU box, checks its decisions against the exact attempt and times it against the product.
#   here:        tools/exp_ptrs2.sh build
#   on the box:  gpurun -- 'tools/exp_ptrs2.sh run'
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p build/variants
  SABC_LIB_OUT=$PWD/build/variants/libsabc_ptrs2.so SABC_EXTRA_NVCC_FLAGS="-DSABC_EXPERIMENTAL_PTRS2" python simulatedannealingabc.jl_b200/build.py --force
  rm -rf build/variants/*.obj
else
  SABC_B200_LIB=$PWD/build/variants/libsabc_ptrs2.so python tools/exp_ptrs2_check.py ${2:-1e10}
  SABC_B200_LIB=$PWD/build/variants/libsabc_ptrs2.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  tools/ab.sh ptrs2 --steps 200 --warmup 3
fi
