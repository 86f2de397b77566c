#!/usr/bin/env bash
# Round-2 run I (one GPU): config 4 with real LD blocks, DRAM traffic of king_ts_kernel at the bench's own shape,
# reader decode throughput per storage mode, default bench line (with the LD / score secondary legs) + reference arm.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== config 4"; ( time timeout 900 python tests/harness/run_configs.py c4 > gpurun_out/config4.json 2> gpurun_out/config4.err ) 2>&1 | tail -3; tail -c 1500 gpurun_out/config4.json; tail -3 gpurun_out/config4.err; rm -rf /tmp/pl2_c4
echo "== decode probe"; timeout 600 bash tests/harness/decode_probe.sh 2>&1 | tail -12 | tee gpurun_out/decode_probe.log
echo "== king traffic at the bench shape"; timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:king_ts_kernel -s 2 -c 1 --csv --log-file gpurun_out/king_traffic_100k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/king_traffic_bench.log 2>&1; grep -E "king_ts" gpurun_out/king_traffic_100k.csv | awk -F'","' '{print $13, $14, $15}'
echo "== bench default"; ( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["clocks"])
for k, v in d.get("secondary", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("workload", "unit", "kernel")})
PY
tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"; ( time timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3; tail -c 400 gpurun_out/bench_ref.json
