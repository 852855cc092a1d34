#!/bin/bash
# 8-GPU evidence: multi-GPU parity tests at world 8, bench at 8 and 4 GPUs (process per GPU and single-process handle), C5 sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8" 2>&1 | tail -12 > gpurun_out/r2_pytest_multi8.log; cat gpurun_out/r2_pytest_multi8.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d.get("roofline", {})
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value %.4g ms/step %.4f e2e %.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)), "mg", d.get("mg_parity"), d.get("run", {}).get("host_numa"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; show gpurun_out/r2_bench_n8.json
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --no-extra > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; show gpurun_out/r2_bench_n4.json
timeout 600 python bench.py --gpus 8 --single-process --steps 100 --no-cpu-baseline > gpurun_out/r2_bench_n8_single_process.json 2> gpurun_out/r2_bench_n8_single_process.err; show gpurun_out/r2_bench_n8_single_process.json
timeout 600 $TR --nproc-per-node 8 --master-port 29523 tools/sweep.py --configs c5 --sizes 800000,8000000,80000000 --steps 50 > gpurun_out/r2_sweep_c5_n8.jsonl 2> gpurun_out/r2_sweep_c5_n8.err; cut -c1-140 gpurun_out/r2_sweep_c5_n8.jsonl
timeout 600 $TR --nproc-per-node 4 --master-port 29524 tools/sweep.py --configs c5 --sizes 400000,4000000,40000000 --steps 50 > gpurun_out/r2_sweep_c5_n4.jsonl 2> gpurun_out/r2_sweep_c5_n4.err; cut -c1-140 gpurun_out/r2_sweep_c5_n4.jsonl
python bench.py --steps 200 --no-cpu-baseline --no-extra --e2e-steps 20 > gpurun_out/r2_bench_n1_same_box.json 2>/dev/null; show gpurun_out/r2_bench_n1_same_box.json
python tools/sweep.py --configs c5 --sizes 100000,1000000,10000000 --steps 50 2>/dev/null | cut -c1-140
