#!/usr/bin/env bash
# One gpurun call: full GPU test suite + full-size bench line + ncu launch list + ncu --set full captures.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench (full 100k x 65536/step)"
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3200 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
SMALL="--samples 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_tensor.csv python bench.py $SMALL > gpurun_out/ncu_bench_tensor.log 2>&1
tail -3 gpurun_out/launches_tensor.csv
echo "== ncu full: king_ts_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:king_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_king_ts python bench.py $SMALL > gpurun_out/ncu_full_ts.log 2>&1; tail -2 gpurun_out/ncu_full_ts.log
if [ -n "$PROFILE_GRM" ]; then
echo "== ncu full: grm_ts_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grm_ts_kernel -s 1 -c 1 -f -o gpurun_out/prof_grm_ts python tools/quick_king_bench.py 16384 65536 1 > gpurun_out/ncu_full_grm.log 2>&1; tail -4 gpurun_out/ncu_full_grm.log
fi
ls -la gpurun_out
