#!/usr/bin/env bash
# Host decode throughput of the reader (PgenReader::GetBlock, multi-threaded) on the same 16,384 x 65,536 genotypes
# stored as .bed (mode 0x01), fixed-width .pgen (0x02) and the reference's default compressed .pgen (0x10:
# difflist / LD-compressed records).  Decoded size = 268 MB of 2-bit rows in every case.
set -e
D=$(mktemp -d); P=oracle/_ref/plink2
python - "$D" <<'PY'
import sys; sys.path.insert(0, "."); import bench
bench.write_synth_bed(sys.argv[1] + "/g", 16384, 65536)
PY
$P --bfile $D/g --make-pgen --threads 16 --out $D/m10 > /dev/null
$P --bfile $D/g --make-pgen format=2 --threads 16 --out $D/m02 > /dev/null
ls -la $D/g.bed $D/m10.pgen $D/m02.pgen | awk '{print $5, $9}'
for src in "--bfile $D/g" "--pfile $D/m02" "--pfile $D/m10"; do
  for rep in 1 2; do PL2_TIMING=1 plink_ng_b200/plink2_b200 $src --make-king-table --king-table-filter 0.3 --out $D/o 2>&1 | grep -E "decode \(PgrGet\)" | sed "s|^|$src : |"; done
done
rm -rf $D
