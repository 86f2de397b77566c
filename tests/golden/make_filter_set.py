#!/usr/bin/env python
"""Writes the ID lists behind the filter fixtures (--keep / --remove / --keep-fam / --remove-fam / --extract /
--exclude) from sets X and S.  Deliberately awkward on purpose: a headerless FID IID list with an unknown ID and a
repeated ID, an `#IID` list (matches because set X has FID 0 throughout), a bare-IID removal list, an --extract file
with many tokens per line and an unknown variant, FID lists for set S, and an `#IID` list for set S that must match
nobody there (its samples carry real FIDs).
usage: make_filter_set.py   (run in tests/golden)"""
import random

rnd = random.Random(11)
fam = [l.split() for l in open("x.fam")]
bim = [l.split() for l in open("x.bim")]
ks = rnd.sample(fam, len(fam) * 2 // 3)
h = len(ks) // 2
open("x_keep1.txt", "w").write("".join("%s %s\n" % (f[0], f[1]) for f in ks[:h]) + "0 nosuch\n" + "%s %s\n" % (ks[0][0], ks[0][1]))
open("x_keep2.txt", "w").write("#IID\n" + "".join("%s\n" % f[1] for f in ks[h:]))
open("x_remove.txt", "w").write("".join("%s\n" % f[1] for f in ks[h - 10:h + 10]))
ex = rnd.sample(bim, 500)
open("x_extract.txt", "w").write(" ".join(b[1] for b in ex[:300]) + "\n" + "\n".join(b[1] for b in ex[300:]) + "\nnosuchvar\n")
open("x_exclude.txt", "w").write("\n".join(b[1] for b in ex[100:180]) + "\n")
fam = [l.split() for l in open("s.fam")]
fids = sorted(set(f[0] for f in fam))
open("s_keepfam.txt", "w").write("\n".join(fids[:3]) + "\n")
open("s_removefam.txt", "w").write(fids[1] + "\n")
ks = rnd.sample(fam, 150)
open("s_keep_iid.txt", "w").write("#IID\n" + "".join("%s\n" % f[1] for f in ks))
# set X's .bim with awkward allele codes for --snps-only [just-acgt]: indels, symbolic alleles, lower case, missing codes
rnd = random.Random(3)
rows = [l.rstrip("\n").split("\t") for l in open("x.bim")]
codes = ["A", "C", "G", "T", "a", "c", "N", "AT", "GCC", ".", "0", "*", "I", "D", "<DEL>", "t"]
for r in rows:
    if rnd.random() < 0.4:
        r[4] = rnd.choice(codes)
    if rnd.random() < 0.25:
        r[5] = rnd.choice(codes)
open("x_alleles.bim", "w").write("".join("\t".join(r) + "\n" for r in rows))
