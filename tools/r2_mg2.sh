#!/bin/bash
# 2-GPU check: multi-GPU parity tests, then the bench at N=2 (C4 and C5 small sizes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_pytest_multi2.log; cat gpurun_out/r2_pytest_multi2.log
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/r2_mg2_$tag.json 2> gpurun_out/r2_mg2_$tag.err; python - "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_mg2_{sys.argv[1]}.json")); print(sys.argv[1], "value %.4g  ms/step %.4f  e2e %.4g  launch %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("launch")))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open(f"gpurun_out/r2_mg2_{sys.argv[1]}.err").read()[-1500:])
PY
}
run c4 --steps 100 --warmup 3
run c5_1e7 --workload c5 --steps 20 --warmup 3 --e2e-steps 3
run c5_1e6 --workload c5 --particles 1000000 --steps 50 --warmup 3 --e2e-steps 3
run c5_1e5 --workload c5 --particles 100000 --steps 200 --warmup 3 --e2e-steps 3
run c5_1e5_nograph --workload c5 --particles 100000 --steps 200 --warmup 3 --e2e-steps 3 --flags 1
for p in 10000000 1000000 100000; do python bench.py --workload c5 --particles $p --steps 50 --warmup 3 --e2e-steps 3 --no-cpu-baseline --graph 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1gpu c5', $p, d['ms_per_step'])"; done
