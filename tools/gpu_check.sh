#!/usr/bin/env bash
# Development aid: one gpurun call = selftest + GPU tests + coarse kernel timing; logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== selftest"; timeout 180 python -c "
import plink_ng_b200 as p
ctx = p.GpuContext(0)
try:
    ctx.selftest_umma(True); print('UMMA selftest OK')
except Exception as e:
    print('UMMA selftest FAILED', e)
" 2>&1 | tail -15 | tee gpurun_out/selftest.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== quick bench"; timeout 600 python tools/quick_king_bench.py ${1:-8192} ${2:-65536} 2>&1 | tail -8 | tee gpurun_out/quick_bench.log
