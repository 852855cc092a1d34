#!/bin/bash
# usage: tools/gpu_prof.sh <tag> <kernel-regex> [bench args...]   (runs on the GPU box via gpurun)
tag=$1; shift; kre=$1; shift
python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 2200 gpurun_out/bench_$tag.json | head -c 900; echo
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s 6 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
