#!/bin/bash
# one gpurun call: GPU tests, the default bench line, the ncu launch list of the same short command and one full capture
# of the dominant kernel -> gpurun_out/ (copy the summaries to profiles/ with tools/ncu_summary.py afterwards)
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err; tail -c 600 gpurun_out/bench_default_$tag.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"simulate_accept|propose_kernel|stats_kernel" -s 6 -c 3 -o gpurun_out/prof_c4_$tag \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
