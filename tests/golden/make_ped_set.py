#!/usr/bin/env python
"""Writes the legacy text filesets behind the --ped/--map import fixture: p.ped + p.map (regular layout: two allele
tokens per call, multi-character alleles, missing calls, a comment line, 4-column .map with centimorgans, two variants
with a negative bp that the import drops, allele frequencies on both sides of 0.5 so that REF/ALT get swapped for
about half the variants) and pc.ped + pc.map (compound-genotypes layout, 3-column .map).
usage: make_ped_set.py   (run in tests/golden)"""
import random

rnd = random.Random(17)
n, m = 40, 60
alleles = []
for v in range(m):
    a, b = rnd.sample(["A", "C", "G", "T"], 2)
    if v % 9 == 4:
        a = a + rnd.choice(["T", "GG", "CAT"])
    if v % 13 == 6:
        b = "<DEL>"
    alleles.append((a, b))
freq = [rnd.uniform(0.1, 0.9) for _ in range(m)]
chrom = sorted(rnd.choice(["1", "2", "X", "MT"]) for _ in range(m))
chrom.sort(key=lambda c: {"1": 1, "2": 2, "X": 23, "MT": 26}[c])


def call(v, compound=False):
    if rnd.random() < 0.06:
        return ("0", "0")
    a, b = alleles[v] if not compound else ("A", "G")
    g = [a if rnd.random() < freq[v] else b for _ in range(2)]
    return (g[0], g[1])


with open("p.ped", "w") as f:
    f.write("# comment line before the first sample\n")
    for k in range(n):
        row = ["F%d" % (k // 4), "s%d" % k, "0" if k % 5 else "s%d" % (k - 1 if k else 0), "0", str(k % 3), str(1 + k % 2) if k % 7 else "-9"]
        for v in range(m):
            row += list(call(v))
        f.write((" " if k % 2 else "\t").join(row) + "\n")
with open("p.map", "w") as f:
    for v in range(m):
        bp = 1000 + 37 * v
        f.write("%s\trs%d\t%s\t%d\n" % (chrom[v], v, "0" if v % 4 else "%.3f" % (v * 0.011), -bp if v in (7, 33) else bp))
with open("pc.ped", "w") as f:
    for k in range(n):
        row = ["0", "c%d" % k, "0", "0", str(1 + k % 2), "-9"]
        for v in range(m):
            row.append("".join(call(v, True)))
        f.write(" ".join(row) + "\n")
with open("pc.map", "w") as f:
    for v in range(m):
        f.write("%s rs%d %d\n" % (chrom[v], v, 500 + 11 * v))
