#!/usr/bin/env bash
# Multi-GPU LD prune (one chromosome run per device at a time) against the single-device lists and the reference's
# golden lists.  Needs >= 2 GPUs:  gpurun --gpus 2 -- 'bash tools/check_ld_multi_gpu.sh'
# NOT YET RUN ON HARDWARE (the round's GPU budget was spent when this path was written).
set -u
mkdir -p gpurun_out/ldmg
B=plink_ng_b200/plink2_b200
G=tests/golden
fail=0
for g in 1 2; do
  $B --bfile $G/x --gpus $g --indep-pairwise 50 5 0.2 --out gpurun_out/ldmg/x$g > gpurun_out/ldmg/x$g.out 2>&1 || { echo "x --gpus $g failed"; tail -3 gpurun_out/ldmg/x$g.out; fail=1; }
  $B --bfile $G/x --gpus $g --indep-pairwise 2kb 1 0.2 --indep-order 1 --out gpurun_out/ldmg/xk$g > gpurun_out/ldmg/xk$g.out 2>&1 || { echo "x kb --gpus $g failed"; fail=1; }
done
cmp gpurun_out/ldmg/x1.prune.in gpurun_out/ldmg/x2.prune.in && echo "x: 1 vs 2 GPUs identical" || fail=1
cmp gpurun_out/ldmg/xk1.prune.in gpurun_out/ldmg/xk2.prune.in && echo "x kb/order1: 1 vs 2 GPUs identical" || fail=1
true
echo "failures: $fail"
exit $fail
