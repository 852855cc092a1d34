#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_engine.py tests/test_golden.py tests/test_gpu_hooks.py tests/test_plugin_model.py -m gpu -q -x 2>&1 | tail -6
for f in 0 256; do for st in 20 200; do python bench.py --steps $st --no-cpu-baseline --no-extra --e2e-steps 10 --flags $f 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags $f steps $st: value %.4g ms/step %.4f kernel %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['avg_kernel_ms'], d['e2e']['value']))"; done; done
python bench.py --workload c4 --particles 10000000 --steps 10 --no-cpu-baseline --no-extra --e2e-steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 1e7: value %.4g ms/step %.4f kernel %.4f' % (d['value'], d['ms_per_step'], d['roofline']['avg_kernel_ms']))"
