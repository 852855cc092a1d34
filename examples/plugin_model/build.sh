#!/bin/bash
# builds libsabc_ar1.so next to this script; same flags as simulatedannealingabc.jl_b200/build.py
set -e
cd "$(dirname "$0")"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off -ccbin /usr/bin/g++ -shared -o libsabc_ar1.so ar1_model.cu \
    -L ../../simulatedannealingabc.jl_b200 -l:libsabc_b200.so -Xlinker -rpath -Xlinker '$ORIGIN/../../simulatedannealingabc.jl_b200'
echo "$(pwd)/libsabc_ar1.so"
