#!/usr/bin/env bash
# Round-2 check A: box CPU facts, smoke (new TMA KING kernel), int8 peak, quick KING/GRM timing, GPU tests.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; echo "nproc $(nproc)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; free -g | head -2; df -h /tmp | tail -1; } > gpurun_out/box.txt 2>&1
cat gpurun_out/box.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== int8 peak"; timeout 120 python tools/int8_peak.py 2>&1 | tail -2 | tee gpurun_out/int8_peak.json
echo "== quick bench"; SKIP_POPC=1 SKIP_SS=1 timeout 300 python tools/quick_king_bench.py 16384 65536 2>&1 | tail -4 | tee gpurun_out/quick_bench.log
echo "== pytest"; ( time timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/pytest_gpu.log
