#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in 0 256; do
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sass__inst_executed_local_loads,sass__inst_executed_local_stores --clock-control none -k regex:"simulate|accept_list|propose_kernel|stats_kernel" -s 20 -c 12 --csv --log-file gpurun_out/r2g_launch_f$f.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 --flags $f > /dev/null 2>&1
done
python - <<'PY'
import csv
for f in (0, 256):
    rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r2g_launch_f{f}.csv", errors="replace") if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    i_n, i_m, i_v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    i_id = hdr.index("ID")
    d = {}
    for r in rows:
        d.setdefault((r[i_id], r[i_n][:60]), {})[r[i_m]] = r[i_v]
    print("flags", f)
    for (i, n), m in list(d.items())[:8]:
        print(" ", i, n, {k.split(".")[0][-28:]: v for k, v in m.items()})
PY
