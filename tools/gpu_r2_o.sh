#!/usr/bin/env bash
# Round-2 run O (one GPU): score modes after the kernel change, --variant-score probe.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest score"; ( time timeout 600 python -m pytest tests/test_score_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_o.log 2>&1; tail -8 gpurun_out/pytest_o.log ) 2>&1 | tee gpurun_out/pytest_o_tail.log
echo "== vscore probe"; timeout 300 python tools/vscore_probe.py 2>&1 | tail -2 | tee gpurun_out/vscore_probe.log
echo "== score probe"; timeout 300 python tools/score_probe.py 2>&1 | tail -1 | tee gpurun_out/score_probe.log
