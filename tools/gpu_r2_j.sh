#!/usr/bin/env bash
# Round-2 run J (one GPU): re-check after the score-kernel rewrite and the per-chromosome LD staging.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest (score, cli, ld)"; ( time timeout 900 python -m pytest tests/test_score_gpu.py tests/test_cli_gpu.py tests/test_ld_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_j.log 2>&1; tail -8 gpurun_out/pytest_j.log ) 2>&1 | tee gpurun_out/pytest_j_tail.log
echo "== score probe"; timeout 300 python tools/score_probe.py 2>&1 | tail -2 | tee gpurun_out/score_probe.log
echo "== config 4"; ( time timeout 900 python tests/harness/run_configs.py c4 > gpurun_out/config4.json 2> gpurun_out/config4.err ) 2>&1 | tail -3; tail -c 1700 gpurun_out/config4.json; tail -3 gpurun_out/config4.err; rm -rf /tmp/pl2_c4
