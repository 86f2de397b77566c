#!/usr/bin/env bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== selftest (SS + TS probes)"; timeout 120 python -c "
import plink_ng_b200 as p
ctx = p.GpuContext(0)
try:
    ctx.selftest_umma(True); print('UMMA selftest OK (SS + TS)')
except Exception as e:
    print('UMMA selftest FAILED', e)
" 2>&1 | tail -12 | tee gpurun_out/selftest.log
echo "== king tests"; timeout 600 python -m pytest tests/test_king_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -15 | tee gpurun_out/pytest_king.log
echo "== quick bench"; 
for d in ${TS_MODES:-0}; do echo "dbg=$d"; PL2_TS_DEBUG=$d SKIP_GRM=${SKIP_GRM:-1} SKIP_POPC=1 timeout 300 python tools/quick_king_bench.py 16384 65536 3 2>&1 | tail -6; done | tee gpurun_out/quick_bench.log
