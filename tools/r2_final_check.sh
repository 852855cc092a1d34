#!/bin/bash
# last call of the round on one GPU: the whole GPU suite, smoke(), the default bench line and the reference arm (short)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_final.log; cat gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_default_final.json")); r = d["roofline"]
    print("value %.4g ms/step %.4f e2e %.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("roofline", r["bound"], r["achieved"], r["peak"], r["frac"], "lane_eff", r.get("lane_efficiency"), "traffic", r["traffic"], "hbm", r["hbm"]["frac"], "event-pass value", r.get("value_per_gpu_in_the_event_pair_pass"))
    print("clocks", d["clocks"]); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]); print("launch", d["run"]["launch"], "gpu_launches", d["gpu_launches"])
    print("extra", {k: (round(v["ms_per_step"], 4), "%.3g" % v["value"]) for k, v in d["extra"].items() if isinstance(v, dict)})
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/r2_bench_default_final.err").read()[-1500:])
PY
SABC_BENCH_REF_SECONDS=10 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference_final.json 2>/dev/null; cut -c1-400 gpurun_out/r2_bench_reference_final.json
