#!/bin/sh
# A/B timing of experimental builds of the CUDA library (build/variants/*.so, made with SABC_LIB_OUT=... build.py):
# runs the same short bench for the product library and every variant, one JSON line each in gpurun_out/ab_<tag>.jsonl
tag=${1:-ab}; shift
mkdir -p gpurun_out
out=gpurun_out/ab_$tag.jsonl; : > $out
for lib in product build/variants/*.so; do
  [ "$lib" = product ] && unset SABC_B200_LIB || export SABC_B200_LIB=$PWD/$lib
  echo "{\"variant\": \"$lib\"}" >> $out
  python bench.py --no-cpu-baseline --e2e-steps 5 "$@" >> $out 2>> gpurun_out/ab_$tag.err
done
python - "$out" <<'PY'
import json, sys
name = None
for l in open(sys.argv[1]):
    d = json.loads(l)
    if "variant" in d: name = d["variant"]; continue
    r = d.get("roofline", {})
    print(f"{name:45s} {d['value']:.4g} upd/s  {d['ms_per_step']:.3f} ms/step  kernel {r.get('avg_kernel_ms')} ms  e2e {d['e2e']['value']:.4g}")
PY
