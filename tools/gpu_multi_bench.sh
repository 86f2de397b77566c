#!/usr/bin/env bash
# N-GPU bench only (run under `gpurun --gpus N`): what the driver's scaling step runs, with an outer timeout.
N=${1:-2}; STEPS=${2:-2}; WARM=${3:-1}
mkdir -p gpurun_out
echo "== bench x$N"; ( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup $WARM > gpurun_out/bench_x$N.json 2> gpurun_out/bench_x$N.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_x$N.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("no json:", e)
PY
tail -5 gpurun_out/bench_x$N.err
