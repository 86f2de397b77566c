#!/usr/bin/env bash
# Round-2 run P (one GPU): config 5's per-GPU share - 200k samples, accumulators beyond HBM, natural multipass,
# rows re-checked by the reference's pair-list path.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== config 5 share"; ( time timeout 1200 python tests/harness/run_configs.py c5 > gpurun_out/config5.json 2> gpurun_out/config5.err ) 2>&1 | tail -3; tail -c 2500 gpurun_out/config5.json; tail -5 gpurun_out/config5.err; rm -rf /tmp/pl2_c5
