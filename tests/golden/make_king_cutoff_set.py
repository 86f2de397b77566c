#!/usr/bin/env python
"""Builds the matrix-prefix input of the `--king-cutoff <prefix> <threshold>` fixture from the reference's own fp64
triangle over set A: 80 of the 100 IDs in shuffled order, three unknown IDs mixed in (the reference drops those lines
before it numbers the matrix rows), and the 80-row fp64 triangle in that order.
usage: make_king_cutoff_set.py <a.king.bin fp64 triangle> <a.king.id> <out prefix>"""
import random
import sys

import numpy as np

tri_path, id_path, out = sys.argv[1:4]
ids = [l.split() for l in open(id_path).read().splitlines()[1:]]
n = len(ids)
full = np.fromfile(tri_path, dtype=np.float64)
assert full.size == n * (n - 1) // 2
m = np.zeros((n, n))
k = 0
for i in range(1, n):
    m[i, :i] = full[k:k + i]
    k += i
m = m + m.T
order = list(range(n))
random.Random(5).shuffle(order)
sel = order[:80]
with open(out + ".king.id", "w") as f:
    f.write("#FID\tIID\n")
    for r, s in enumerate(sel):
        f.write("%s\t%s\n" % tuple(ids[s]))
        if r in (5, 40, 79):
            f.write("0\tnosuch%d\n" % r)
np.array([m[sel[i], sel[j]] for i in range(1, 80) for j in range(i)], dtype=np.float64).tofile(out + ".king.bin")
