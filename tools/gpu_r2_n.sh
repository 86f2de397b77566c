#!/usr/bin/env bash
# Round-2 run N (one GPU): L2 eviction hints on king_ts_kernel - parity, DRAM traffic and step time at the bench shape.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest king + scale"; ( time timeout 900 python -m pytest tests/test_king_gpu.py tests/test_scale_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_n.log 2>&1; tail -5 gpurun_out/pytest_n.log ) 2>&1 | tee gpurun_out/pytest_n_tail.log
echo "== traffic"; timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:king_ts_kernel -s 2 -c 1 --csv --log-file gpurun_out/king_traffic_100k_hint.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/king_traffic_bench.log 2>&1; grep -E "king_ts" gpurun_out/king_traffic_100k_hint.csv | awk -F'","' '{print $13, $14, $15}'
echo "== bench (kernel loop only)"; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_hint.json 2> gpurun_out/bench_hint.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_hint.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], d["clocks"])
PY
