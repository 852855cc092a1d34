#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_engine.py tests/test_golden.py -m gpu -q -x -k "not full_size and not posterior and not very_large" 2>&1 | tail -3
for f in 0 256; do for st in 20 200; do python bench.py --steps $st --no-cpu-baseline --no-extra --e2e-steps 10 --flags $f 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags $f steps $st: value %.4g ms/step %.4f kernel %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['avg_kernel_ms'], d['e2e']['value']))"; done; done
for f in 0 256; do
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"simulate|accept_list" -s 20 -c 4 --csv --log-file gpurun_out/r2h_launch_f$f.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 --flags $f > /dev/null 2>&1
done
python - <<'PY'
import csv
for f in (0, 256):
    rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r2h_launch_f{f}.csv", errors="replace") if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    i_n, i_m, i_v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    i_id = hdr.index("ID")
    d = {}
    for r in rows:
        d.setdefault((r[i_id], r[i_n][:40]), {})[r[i_m]] = r[i_v]
    print("flags", f)
    for (i, n), m in list(d.items())[:4]:
        print(" ", i, n, {k.split(".")[0][-28:]: v for k, v in m.items()})
PY
