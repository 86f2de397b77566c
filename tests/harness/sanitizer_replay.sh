#!/usr/bin/env bash
# Builds the host program with ASan + UBSan (and once with TSan) and replays the host-only scenarios and, through the
# stand-in device library (mock_pl2gpu.cc), the device-driver scenarios of tests/test_host_orchestration.py.
# CPU-only; run from the repo root:  bash tests/harness/sanitizer_replay.sh      (expects no sanitizer output)
set -u
SRC="host/dataset.cc host/pca.cc host/pgen_reader.cc host/plink2_b200.cc host/sfmt.cc host/text_util.cc host/filters.cc host/ped_import.cc"
W=$(mktemp -d); R=$PWD
( cd plink_ng_b200/csrc && g++ -O1 -g -std=c++17 -mbmi2 -ffp-contract=off -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o $W/asan $SRC -L.. -lpl2gpu -Wl,-rpath,$R/plink_ng_b200 \
  && g++ -O1 -g -std=c++17 -mbmi2 -ffp-contract=off -pthread -fsanitize=thread -o $W/tsan $SRC -L.. -lpl2gpu -Wl,-rpath,$R/plink_ng_b200 ) || exit 1
g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o $W/mock.so tests/harness/mock_pl2gpu.cc
g++ -O1 -g -std=c++17 -ffp-contract=off -shared -fPIC -fsanitize=thread -o $W/mock_tsan.so tests/harness/mock_pl2gpu.cc
export ASAN_OPTIONS=detect_leaks=0
LA=$(g++ -print-file-name=libasan.so); LT=$(g++ -print-file-name=libtsan.so)
cd tests/golden
flt() { grep -i "runtime error\|AddressSanitizer\|ThreadSanitizer" ; }
$W/asan --bfile x --keep x_keep1.txt x_keep2.txt --remove x_remove.txt --extract x_extract.txt --exclude x_exclude.txt --mind 0.035 --geno 0.02 --maf 0.05 --make-bed --make-pgen --write-snplist --out $W/a1 2>&1 | flt
$W/asan --pedmap p --make-bed --out $W/a2 2>&1 | flt
$W/asan --pfile ma --max-alleles 2 --threads 3 --make-bed --out $W/a3 2>&1 | flt
$W/asan --bed x.bed --pvar x.pvar.zst --fam x.fam --king-cutoff-table a_kingp.kin0.gz 0.02 --make-bed --out $W/a4 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" PL2_MOCK_DEVICES=3 $W/asan --bed a.bed --bim a_chr6.bim --fam a.fam --gpus 3 --indep-pairwise 50 5 0.2 --out $W/b1 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile a --king-cutoff 0.02 --indep-pairwise 50 5 0.2 --out $W/b2 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile x --not-chr X --keep x_keep1.txt x_keep2.txt --r2-unphased --ld-window-r2 0.3 --ld-window-kb 0.1 --out $W/b3 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile x --nonfounders --freq counts --out $W/b4 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile a --gpu-memory 4 --make-king-table counts cols=+ibs1,+ibs --make-king bin4 triangle --out $W/b5 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile a --make-grm-sparse 0.02 --pca 4 --out $W/b6 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile a --king-cutoff 0.02 --score a_score.txt header dominant list-variants cols=+scoresums,+denom --out $W/b7 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile a --variant-score a_vscore_weights.txt cols=+altfreq zs --out $W/b8 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile s --make-king-table rel-check --out $W/b9 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile a --make-king-table counts cols=+ibs1 --king-table-subset a_sub2.txt --out $W/b10 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile x --keep x_keep2.txt --mind 0.05 --make-rel square zs --make-grm-list --out $W/b11 2>&1 | flt
LD_PRELOAD="$LA $W/mock.so" $W/asan --bfile x --nonfounders --r2-unphased --ld-window-r2 0.1 --out $W/b12 2>&1 | flt
LD_PRELOAD="$LT $W/mock_tsan.so" $W/tsan --bfile x --r2-unphased --ld-window-r2 0.05 --threads 5 --out $W/t2 2>&1 | flt
LD_PRELOAD="$LT $W/mock_tsan.so" PL2_MOCK_DEVICES=3 $W/tsan --bed a.bed --bim a_chr6.bim --fam a.fam --gpus 3 --threads 6 --indep-pairwise 50 5 0.2 --out $W/t1 2>&1 | flt
cmp $W/b1.prune.in a_chr6.prune.in && cmp $W/t1.prune.in a_chr6.prune.in && cmp $W/b2.prune.in g_acut.prune.in && echo "sanitizer replay: outputs as expected, no reports above"
