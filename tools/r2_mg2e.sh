#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "2" 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29531 tools/sweep.py --configs c5,c4 --sizes 200000,2000000,20000000 --steps 100 2>/dev/null | cut -c1-150
