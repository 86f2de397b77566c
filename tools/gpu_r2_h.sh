#!/usr/bin/env bash
# Round-2 run H (two GPUs): multi-GPU product tests (KING, GRM, approx PCA) and the 2-GPU bench line.
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== pytest multi-gpu"; ( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_h.log 2>&1; tail -12 gpurun_out/pytest_h.log ) 2>&1 | tee gpurun_out/pytest_h_tail.log
bash tools/gpu_multi_bench.sh 2 3 3
cp gpurun_out/bench_x2.json gpurun_out/bench_x2_r02.json 2>/dev/null
