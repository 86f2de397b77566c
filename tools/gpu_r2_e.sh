#!/usr/bin/env bash
# Round-2 run E (one GPU): new LD forms (chrX/Y/MT, --indep-order 1), PCA residual digit pass + Gram final stage,
# basis comparison (jacobi vs bcgs), phase timings, launch list restricted to the library's kernels.
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
echo "== pytest (ld, pca, cli)"; ( time timeout 1200 python -m pytest tests/test_cli_gpu.py tests/test_ld_gpu.py tests/test_pca_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_e.log 2>&1; tail -15 gpurun_out/pytest_e.log ) 2>&1 | tee gpurun_out/pytest_e_tail.log
echo "== pca basis compare"; timeout 900 python tests/harness/pca_basis_compare.py 2>&1 | tail -12 | tee gpurun_out/pca_basis_compare.log
echo "== pca timing"; for basis in jacobi bcgs; do PL2_PCA_BASIS=$basis timeout 300 python tools/pca_timing.py 16384 65536 20 2>&1 | tail -6 | tee gpurun_out/pca_timing_$basis.log; done
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"king|geno|pad|finalize|filter" -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --samples 16384 --batch-variants 65536 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/bench_under_ncu.log; grep -c king_ts gpurun_out/launches_bench.csv
