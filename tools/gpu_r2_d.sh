#!/usr/bin/env bash
# Round-2 evidence run D (one GPU): box facts, smoke, GPU tests, int8 peak, default bench + reference arm,
# phase timings, and the ncu captures / launch list that profiles/r02_* summarise.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; echo "nproc $(nproc)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; free -g | head -2; df -h /tmp | tail -1; } > gpurun_out/box.txt 2>&1
cat gpurun_out/box.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== pytest"; ( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_full.log 2>&1; tail -8 gpurun_out/pytest_full.log ) 2>&1 | tee gpurun_out/pytest_gpu.log
echo "== int8 peak"; timeout 120 python tools/int8_peak.py 2>&1 | tail -1 | tee gpurun_out/int8_peak.json
echo "== bench default"; ( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; tail -c 2500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"; ( time timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3; tail -c 600 gpurun_out/bench_ref.json
echo "== pca timing"; timeout 300 python tools/pca_timing.py 16384 65536 20 2>&1 | tail -10 | tee gpurun_out/pca_timing_tensor.log
PL2_PCA_ALGO=fp64 timeout 300 python tools/pca_timing.py 16384 65536 20 2>&1 | tail -10 | tee gpurun_out/pca_timing_fp64.log
echo "== ld bench"; timeout 300 python tools/ld_bench.py 2>&1 | tail -2 | tee gpurun_out/ld_bench.log
PL2_LD_ALGO=popcount timeout 300 python tools/ld_bench.py 50000 32768 2>&1 | tail -1 | tee -a gpurun_out/ld_bench.log
echo "== quick"; SKIP_POPC=1 SKIP_SS=1 timeout 300 python tools/quick_king_bench.py 16384 65536 2>&1 | tail -3 | tee gpurun_out/quick_bench.log
echo "== ncu king"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:king_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_king_ts env SKIP_POPC=1 SKIP_SS=1 SKIP_GRM=1 python tools/quick_king_bench.py 16384 65536 1 > gpurun_out/ncu_full_king.log 2>&1; tail -2 gpurun_out/ncu_full_king.log
echo "== ncu grm"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:grm_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_grm_ts env SKIP_POPC=1 SKIP_SS=1 python tools/quick_king_bench.py 16384 65536 1 > gpurun_out/ncu_full_grm.log 2>&1; tail -2 gpurun_out/ncu_full_grm.log
echo "== ncu ld"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:ld_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_ld_ts python tools/ld_bench.py 50000 32768 > gpurun_out/ncu_full_ld.log 2>&1; tail -2 gpurun_out/ncu_full_ld.log
echo "== ncu pca"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:pca_x.*_ts_kernel -s 2 -c 2 -f -o gpurun_out/prof_pca_ts python tools/pca_timing.py 16384 65536 20 > gpurun_out/ncu_full_pca.log 2>&1; tail -2 gpurun_out/ncu_full_pca.log
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench.csv python bench.py --samples 16384 --batch-variants 65536 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log
ls -la gpurun_out
