#!/usr/bin/env bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
echo "== cli tests (grm-list)"; timeout 300 python -m pytest tests/test_cli_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "== LD bench"; timeout 300 python tools/ld_bench.py 50000 131072 500 2>&1 | tail -2 | tee gpurun_out/ld_bench.log
echo "== CLI timing"; bash tools/gpu_cli_timing.sh
