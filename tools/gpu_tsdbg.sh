#!/usr/bin/env bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0 SKIP_POPC=1 SKIP_GRM=1
for d in ${TS_MODES:-0 8}; do echo "dbg=$d"; PL2_TS_DEBUG=$d timeout 200 python tools/quick_king_bench.py 16384 ${TS_M:-65536} ${TS_REPS:-2} 2>&1 | grep -E "tensor_ts|row warp 0|row warp 4|col warp|issuer" | sort | uniq -c | sort -rn | head -${TS_LINES:-14}; done | tee gpurun_out/tsdbg.log
