#!/usr/bin/env bash
# Round-2 run M (one GPU): --variant-score, rel-check and the rest of the GPU suite.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest score/cli first"; ( time timeout 900 python -m pytest tests/test_score_gpu.py tests/test_cli_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_m1.log 2>&1; tail -12 gpurun_out/pytest_m1.log ) 2>&1 | tee gpurun_out/pytest_m_tail.log
echo "== pytest -m gpu (all)"; ( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_m.log 2>&1; tail -6 gpurun_out/pytest_m.log ) 2>&1 | tee -a gpurun_out/pytest_m_tail.log
