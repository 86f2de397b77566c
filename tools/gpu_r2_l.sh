#!/usr/bin/env bash
# Round-2 run L (one GPU): the whole GPU suite after the last host-side additions (score modes, rel-check, Krylov
# threshold), smoke, and the ncu capture of the score kernel.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest -m gpu"; ( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_l.log 2>&1; tail -8 gpurun_out/pytest_l.log ) 2>&1 | tee gpurun_out/pytest_l_tail.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== ncu score"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 1 -c 1 -f -o gpurun_out/prof_score python tools/score_probe.py > gpurun_out/ncu_full_score.log 2>&1; tail -2 gpurun_out/ncu_full_score.log
