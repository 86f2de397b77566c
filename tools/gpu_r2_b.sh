#!/usr/bin/env bash
# Round-2 check B: triage of the TMA kernels (compute-sanitizer when the plain run fails), then the usual list.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== min king"; timeout 120 python tools/king_ts_min.py 257 300 king 2>&1 | tail -3 | tee gpurun_out/min_king.log
if ! grep -q "king ts == popcount: True" gpurun_out/min_king.log; then
  echo "== sanitizer king"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/king_ts_min.py 257 300 king 2>&1 | grep -v "^$" | head -60 | tee gpurun_out/sanitizer_king.log
fi
echo "== min grm"; timeout 120 python tools/king_ts_min.py 257 300 grm 2>&1 | tail -3 | tee gpurun_out/min_grm.log
if ! grep -q "grm ok" gpurun_out/min_grm.log; then
  echo "== sanitizer grm"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/king_ts_min.py 257 300 grm 2>&1 | grep -v "^$" | head -60 | tee gpurun_out/sanitizer_grm.log
fi
if grep -q "king ts == popcount: True" gpurun_out/min_king.log; then
  echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
  echo "== int8 peak"; timeout 120 python tools/int8_peak.py 2>&1 | tail -2 | tee gpurun_out/int8_peak.json
  echo "== quick bench"; SKIP_POPC=1 SKIP_SS=1 timeout 300 python tools/quick_king_bench.py 16384 65536 2>&1 | tail -4 | tee gpurun_out/quick_bench.log
  echo "== ld bench"; timeout 300 python tools/ld_bench.py 2>&1 | tail -6 | tee gpurun_out/ld_bench.log
  echo "== debug cli"; REPS=12 timeout 600 python tests/harness/debug_king_cli.py 3000 4096 --gpu-memory 640 2>&1 | grep -v "^rc" | tail -40 | tee gpurun_out/debug_cli.log
  echo "== pytest"; ( time timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_full.log 2>&1; tail -25 gpurun_out/pytest_full.log ) 2>&1 | tee gpurun_out/pytest_gpu.log
  grep -n "differs from" -A6 gpurun_out/pytest_full.log | head -40
fi
if grep -q "king ts == popcount: True" gpurun_out/min_king.log; then
  echo "== bench default"; ( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
  echo "== ncu full: king_ts_kernel"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:king_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_king_ts env SKIP_POPC=1 SKIP_SS=1 SKIP_GRM=1 python tools/quick_king_bench.py 16384 65536 1 > gpurun_out/ncu_full_king.log 2>&1; tail -3 gpurun_out/ncu_full_king.log
fi
