#!/bin/sh
# builds and runs the CPU check of the PTRS acceptance filter (see check_ptrs_filter.cpp)
set -e
cd "$(dirname "$0")"
OUT=${TMPDIR:-/tmp}/sabc_check_filter
mkdir -p "$OUT"
CXXFLAGS="-O2 -std=c++17 -ffp-contract=off -mfma -fopenmp -x c++"
g++ $CXXFLAGS -DSABC_NO_PTRS_FILTER -Dsabc=sabc_exact -c check_ptrs_filter.cpp -o "$OUT/exact.o"
g++ $CXXFLAGS -c check_ptrs_filter.cpp -o "$OUT/filtered.o"
g++ -fopenmp "$OUT/exact.o" "$OUT/filtered.o" -o "$OUT/check_ptrs_filter" -lm
exec "$OUT/check_ptrs_filter" "$@"
