#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/ab.sh r2k_c3 --workload c3 --steps 50 --warmup 3 --no-extra
