#!/bin/bash
# final round-2 evidence on one GPU: default bench line, ncu launch list of the same short command, full captures of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/plain_r2final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r2final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_launches_r2final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"simulate_accept|propose_kernel|stats_kernel" -s 6 -c 3 -o gpurun_out/prof_c4_r2final \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_full_c4_r2final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_half -s 6 -c 1 -o gpurun_out/prof_c5_r2final \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_full_c5_r2final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_half -s 6 -c 1 -o gpurun_out/prof_c2_r2final \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_full_c2_r2final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"simulate_accept" -s 2 -c 1 -o gpurun_out/prof_c3_r2final \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_full_c3_r2final.log 2>&1
tail -1 gpurun_out/ncu_full_c3_r2final.log
python tools/sweep.py > gpurun_out/r2_sweep_n1.jsonl 2> gpurun_out/r2_sweep_n1.err; tail -3 gpurun_out/r2_sweep_n1.jsonl | cut -c1-300
