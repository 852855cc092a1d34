#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_2gpu_b.log; cat gpurun_out/r2_pytest_gpu_2gpu_b.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d.get("roofline", {})
    print(sys.argv[1], "value %.4g ms/step %.4f e2e %.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)), "roof", r.get("bound"), r.get("frac"), "mg", d.get("mg_parity"), "extra", {k: (round(v.get("ms_per_step", 0), 4) if isinstance(v, dict) else v) for k, v in d.get("extra", {}).items()})
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; show gpurun_out/r2_bench_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; show gpurun_out/r2_bench_n2.json
python bench.py --gpus 2 --single-process --steps 100 --no-cpu-baseline > gpurun_out/r2_bench_n2_single_process.json 2> gpurun_out/r2_bench_n2_single_process.err; show gpurun_out/r2_bench_n2_single_process.json
