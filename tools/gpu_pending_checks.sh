#!/usr/bin/env bash
# Everything that was written after the round's GPU budget ran out and still wants one pass on hardware.
# One GPU:   gpurun --timeout 300 -- 'bash tools/gpu_pending_checks.sh'
# Two GPUs:  gpurun --gpus 2 --timeout 300 -- 'bash tools/check_ld_multi_gpu.sh'
mkdir -p gpurun_out
echo "== filters / prune chaining through the device commands"; timeout 170 python tools/check_r2s.py 2>&1 | tail -30
echo "== --r2-unphased"; timeout 120 bash tools/check_r2_unphased.sh 2>&1 | tail -6
echo "== new pytest file"; timeout 300 python -m pytest tests/test_filters_gpu.py -m gpu -q 2>&1 | tail -3
