#!/usr/bin/env bash
# multi-GPU bench check: torchrun, one rank per GPU
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/multi_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_n$N.json; tail -8 gpurun_out/bench_n$N.err
