#!/usr/bin/env bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0 SKIP_POPC=1 SKIP_GRM=1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:king_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_king_ts python tools/quick_king_bench.py 16384 16384 1 > gpurun_out/ncu_full_ts.log 2>&1; tail -3 gpurun_out/ncu_full_ts.log
ls -la gpurun_out/prof_king_ts.ncu-rep
