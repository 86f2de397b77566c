#!/usr/bin/env bash
# --r2-unphased on the device (screen = pl2gpu_ld_band_flags) against the reference's tables.  One GPU, seconds.
# Run on the B200 as part of tools/gpu_quick.sh (profiles/r02_quick_check.txt): identical.  The host half is also pinned
# on the CPU through tests/test_host_orchestration.py (stand-in library).
set -u
mkdir -p gpurun_out/r2u
B=$PWD/plink_ng_b200/plink2_b200
cd tests/golden
fail=0
chk() { name=$1; gold=$2; shift 2; $B "$@" --out ../../gpurun_out/r2u/$name > ../../gpurun_out/r2u/$name.out 2>&1 || { echo "$name: run failed"; tail -3 ../../gpurun_out/r2u/$name.out; fail=1; return; }
  if zcat $gold | cmp -s - ../../gpurun_out/r2u/$name.vcor; then echo "$name: identical"; else echo "$name: MISMATCH"; fail=1; fi; }
chk a_all a_r2.vcor.gz --bfile a --r2-unphased
chk a_win a_r2w.vcor.gz --bfile a --r2-unphased --ld-window 7 --ld-window-r2 0.5
chk x_keep x_r2.vcor.gz --bfile x --not-chr X --keep x_keep1.txt x_keep2.txt --r2-unphased --ld-window-r2 0.3 --ld-window-kb 0.1
# chrX rows are computed on the host (no device screen); these two tables were added after the last GPU slot
chk x_chrx x_r2x.vcor.gz --bfile x --keep x_keep1.txt x_keep2.txt --r2-unphased --ld-window-r2 0.1 --ld-window-kb 0.05
chk x_nf x_r2nf.vcor.gz --bfile x --nonfounders --r2-unphased --ld-window-r2 0.1 --chr X,Y
echo "failures: $fail"; exit $fail
