#!/usr/bin/env bash
# ~10 s on one GPU: the host-side changes made after the last full GPU pass, most important first.
mkdir -p gpurun_out/q
B=$PWD/plink_ng_b200/plink2_b200
O=$PWD/gpurun_out/q
cd tests/golden
{
run() { name=$1; shift; $B "$@" --out $O/$name > $O/$name.out 2>&1; echo "$name rc=$?"; }
same() { if cmp -s "$1" "$2"; then echo "  $3: identical"; else echo "  $3: MISMATCH"; fi; }
run ld_a --bfile a --indep-pairwise 50 5 0.2;                   same $O/ld_a.prune.in a_ld.prune.in "LD prune set A (worker loop, one device)"
run ld_x --bfile x --indep-pairwise 50 5 0.2;                   same $O/ld_x.prune.in x_o2.prune.in "LD prune set X (sex chromosomes)"
run g4 --bfile a --king-cutoff 0.02 --indep-pairwise 50 5 0.2;  same $O/g4.prune.in g_acut.prune.in "--king-cutoff -> LD prune (frozen freqs, guard moved)"
run r2a --bfile a --r2-unphased;                                 zcat a_r2.vcor.gz | cmp -s - $O/r2a.vcor && echo "  --r2-unphased set A: identical" || echo "  --r2-unphased set A: MISMATCH"
run r2w --bfile a --r2-unphased --ld-window 7 --ld-window-r2 0.5; zcat a_r2w.vcor.gz | cmp -s - $O/r2w.vcor && echo "  --r2-unphased window: identical" || echo "  --r2-unphased window: MISMATCH"
run r2x --bfile x --not-chr X --keep x_keep1.txt x_keep2.txt --r2-unphased --ld-window-r2 0.3 --ld-window-kb 0.1; zcat x_r2.vcor.gz | cmp -s - $O/r2x.vcor && echo "  --r2-unphased set X: identical" || echo "  --r2-unphased set X: MISMATCH"
} 2>&1 | tee $O/../quick_check.txt
