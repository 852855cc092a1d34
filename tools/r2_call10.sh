#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
tools/ab.sh r2j_c3 --workload c3 --steps 50 --warmup 3 --no-extra
unset SABC_B200_LIB
for wl in c5 c2 c4 c1; do python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-extra --e2e-steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', '%.4g' % d['value'], d['ms_per_step'], d['roofline']['avg_kernel_ms'])"; done
