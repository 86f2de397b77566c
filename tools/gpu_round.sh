#!/usr/bin/env bash
# Round check: what the driver runs at round end (tests, smoke, bench) + GRM profile + eigensolver timing.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest -m gpu"; ( time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== eigen timing"; for n in 1024 4096; do timeout 300 python tools/eigen_timing.py $n 4096 2>&1 | tail -1; done | tee gpurun_out/eigen_timing.log
echo "== bench default"; ( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; tail -c 1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3; tail -c 600 gpurun_out/bench_ref.json
echo "== ncu full: grm_ts_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grm_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_grm_ts env SKIP_POPC=1 SKIP_SS=1 python tools/quick_king_bench.py 16384 65536 1 > gpurun_out/ncu_full_grm.log 2>&1; tail -3 gpurun_out/ncu_full_grm.log
