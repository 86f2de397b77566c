#!/usr/bin/env bash
mkdir -p gpurun_out /tmp/pp
export CUDA_VISIBLE_DEVICES=0
for i in 1 2 3; do
( time timeout 120 plink_ng_b200/plink2_b200 --bfile tests/golden/a --pca 4 --out /tmp/pp/p$i ) > gpurun_out/pca_cli_$i.log 2>&1; echo "run $i rc=$?"; tail -5 gpurun_out/pca_cli_$i.log
done
( time timeout 120 plink_ng_b200/plink2_b200 --bfile tests/golden/a --make-grm-bin --out /tmp/pp/g ) 2>&1 | tail -6
