#!/usr/bin/env bash
# Round-2 run G (one GPU): the whole GPU test-suite (what the driver runs at round end), smoke, config 4 with real LD
# blocks, and a --score throughput probe.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest -m gpu"; ( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_g.log 2>&1; tail -8 gpurun_out/pytest_g.log ) 2>&1 | tee gpurun_out/pytest_g_tail.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== config 4"; ( time timeout 900 python tests/harness/run_configs.py c4 > gpurun_out/config4.json 2> gpurun_out/config4.err ) 2>&1 | tail -3; tail -c 1500 gpurun_out/config4.json; tail -3 gpurun_out/config4.err; rm -rf /tmp/pl2_c4
echo "== score probe"; timeout 300 python tools/score_probe.py 2>&1 | tail -3 | tee gpurun_out/score_probe.log
