#!/usr/bin/env bash
# Round-2 run Q (one GPU): what the driver runs at round end - the GPU suite, smoke, a short default bench.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest -m gpu"; ( time timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_q.log 2>&1; tail -6 gpurun_out/pytest_q.log ) 2>&1 | tee gpurun_out/pytest_q_tail.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== cli timing"; D=$(mktemp -d); python - "$D" <<'PY'
import sys
sys.path.insert(0, ".")
import bench
bench.write_synth_bed(sys.argv[1] + "/g", 16384, 65536)
PY
for rep in 1 2; do PL2_TIMING=1 plink_ng_b200/plink2_b200 --bfile $D/g --make-king-table --king-table-filter 0.3 --out $D/o 2>&1 | grep -E "timing" | tr '\n' ';'; echo; done; rm -rf $D
