#!/bin/bash
# final round-2 evidence on one GPU: ncu launch list of the short bench command, full captures of the dominant kernels reduced to
# text summaries ON THE BOX (the .ncu-rep files are too large to bring back), kernel counts for bench.py, C5 sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out profiles
T=/tmp/ncu_r2; mkdir -p $T
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > $T/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r2final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > $T/l.log 2>&1
cap() { tag=$1; kre=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -o $T/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 "$@" > $T/ncu_$tag.log 2>&1
  python tools/ncu_summary.py report $T/prof_$tag.ncu-rep > gpurun_out/r2_${tag}_final_ncu.txt 2>&1
  ncu -i $T/prof_$tag.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:"$kre" --launch-count 1 > $T/src_$tag.csv 2>/dev/null
  python tools/ncu_source_hot.py $T/src_$tag.csv 40 > gpurun_out/r2_${tag}_final_hotlines.txt 2>&1
}
cap c4 "simulate_accept|propose_kernel|stats_kernel" 6 3
ncu -i $T/prof_c4.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:simulate_accept --launch-count 1 > $T/src_c4s.csv 2>/dev/null
python tools/ncu_source_hot.py $T/src_c4s.csv 40 > gpurun_out/r2_c4_final_hotlines.txt 2>&1
cap c5 update_half 6 1 --workload c5
cap c2 update_half 6 1 --workload c2
cap c3 simulate_accept 2 1 --workload c3
python tools/make_kernel_counts.py $T/prof_c4.ncu-rep:625000 $T/prof_c5.ncu-rep:5000000 $T/prof_c2.ncu-rep:50000 $T/prof_c3.ncu-rep:500000 > $T/counts.log 2>&1; cp profiles/r2_kernel_counts.json gpurun_out/r2_kernel_counts.json
python tools/sweep.py > gpurun_out/r2_sweep_n1.jsonl 2> $T/sweep.err; tail -2 gpurun_out/r2_sweep_n1.jsonl | cut -c1-200
ls -la gpurun_out | head -20
