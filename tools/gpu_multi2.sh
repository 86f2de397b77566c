#!/usr/bin/env bash
# N-GPU check (run under `gpurun --gpus N`): CLI --gpus tests, then the bench the driver's scaling step runs.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/multi_box.txt
python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max
echo "== min"; timeout 120 python tools/king_ts_min.py 257 300 king 2>&1 | tail -1; timeout 120 python tools/king_ts_min.py 257 300 grm 2>&1 | tail -1
echo "== quick"; SKIP_POPC=1 SKIP_SS=1 timeout 300 python tools/quick_king_bench.py 16384 65536 2>&1 | tail -3 | tee gpurun_out/quick_bench.log
echo "== pytest multi"; timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest_multi.log
echo "== bench x$N"; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_x$N.json 2> gpurun_out/bench_x$N.err ) 2>&1 | tail -3; tail -c 2500 gpurun_out/bench_x$N.json; tail -8 gpurun_out/bench_x$N.err
