#!/usr/bin/env bash
# Round-2 run K (one GPU): exact-PCA Krylov solver tests + timing, then the whole GPU suite.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest pca"; ( time timeout 900 python -m pytest tests/test_pca_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_k_pca.log 2>&1; tail -12 gpurun_out/pytest_k_pca.log ) 2>&1 | tee gpurun_out/pytest_k_tail.log
echo "== eigen timing"; { PL2_EIGEN=jacobi timeout 600 python tools/eigen_timing.py 4096 8192 2>&1 | tail -1 | sed 's/^/jacobi: /'; PL2_TIMING=1 PL2_EIGEN=krylov timeout 600 python tools/eigen_timing.py 4096 8192 2>&1 | tail -2 | sed 's/^/krylov: /'; PL2_TIMING=1 timeout 600 python tools/eigen_timing.py 10000 8192 2>&1 | tail -2 | sed 's/^/krylov: /'; } | tee gpurun_out/eigen_timing.log
echo "== pytest -m gpu (all)"; ( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_k.log 2>&1; tail -6 gpurun_out/pytest_k.log ) 2>&1 | tee -a gpurun_out/pytest_k_tail.log
