#!/usr/bin/env bash
# Round-2 run F (one GPU): re-check of the fixed pieces, then BASELINE configs 2, 4 and 3 at their stated sizes
# through the product with the reference timed beside them (tests/harness/run_configs.py).
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
df -h /tmp /dev/shm | tee gpurun_out/df.txt
echo "== pytest (cli, ld, pca)"; ( time timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_ld_gpu.py tests/test_pca_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_f.log 2>&1; tail -6 gpurun_out/pytest_f.log ) 2>&1 | tee gpurun_out/pytest_f_tail.log
echo "== config 2"; ( time timeout 900 python tests/harness/run_configs.py c2 > gpurun_out/config2.json 2> gpurun_out/config2.err ) 2>&1 | tail -3; tail -c 1800 gpurun_out/config2.json; tail -3 gpurun_out/config2.err; rm -rf /tmp/pl2_c2
echo "== config 4"; ( time timeout 900 python tests/harness/run_configs.py c4 > gpurun_out/config4.json 2> gpurun_out/config4.err ) 2>&1 | tail -3; tail -c 1500 gpurun_out/config4.json; tail -3 gpurun_out/config4.err; rm -rf /tmp/pl2_c4
echo "== config 3"; ( time timeout 1200 python tests/harness/run_configs.py c3 > gpurun_out/config3.json 2> gpurun_out/config3.err ) 2>&1 | tail -3; tail -c 2500 gpurun_out/config3.json; tail -5 gpurun_out/config3.err; rm -rf /tmp/pl2_c3
