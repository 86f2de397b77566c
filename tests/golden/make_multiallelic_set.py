#!/usr/bin/env python
"""Writes ma.vcf: 200 samples x 80 variants on chr1, every fourth variant triallelic, and the variant right behind each
triallelic one a near-copy of its main track - so that the reference's --make-pgen LD-compresses a BIALLELIC record
against a MULTIALLELIC base.  make_golden.sh imports it with the reference (ma.pgen/.pvar/.psam) and writes the
`--max-alleles 2 --make-bed` result (ma_bi.*).
usage: make_multiallelic_set.py   (run in tests/golden)"""
import random

rnd = random.Random(4)
n = 200
hdr = ("##fileformat=VCFv4.2\n##contig=<ID=1>\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"GT\">\n"
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join("s%d" % i for i in range(n)) + "\n")
rows = []
prev = None
for v in range(80):
    multi = v % 4 == 1
    alts = "C,T" if multi else "C"
    if prev is not None and v % 4 == 2:
        gts = [g.replace("2", "1") for g in prev]
        for k in rnd.sample(range(n), 2):
            gts[k] = "0/0"
    else:
        gts = []
        for s in range(n):
            if rnd.random() < 0.03:
                gts.append("./.")
                continue
            a = [rnd.choice([0, 0, 0, 1, 2] if multi else [0, 0, 1]) for _ in range(2)]
            gts.append("%d/%d" % tuple(sorted(a)))
    prev = gts
    rows.append("1\t%d\tv%d\tA\t%s\t.\t.\t.\tGT\t%s\n" % (100 + v, v, alts, "\t".join(gts)))
open("ma.vcf", "w").write(hdr + "".join(rows))
