#!/bin/bash
# PCA tensor-path check on one GPU: parity tests, then the phase timing for both algorithms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pca_gpu.py -m gpu -q --timeout 400 -x 2>&1 | tail -30 | tee gpurun_out/pytest_pca.log
echo "== timing tensor"
PL2_TIMING=1 timeout 300 python tools/pca_timing.py 16384 65536 20 2>&1 | tail -8 | tee gpurun_out/pca_timing_tensor.log
echo "== timing fp64"
PL2_PCA_ALGO=fp64 PL2_TIMING=1 timeout 300 python tools/pca_timing.py 16384 65536 20 2>&1 | tail -8 | tee gpurun_out/pca_timing_fp64.log
