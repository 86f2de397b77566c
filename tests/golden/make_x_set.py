"""Set X for the sex-chromosome forms of --indep-pairwise: 120 samples x 800 variants of the reference's --dummy
generator re-labelled as chromosomes 1 / X / Y / XY / MT, with males, females, unknown-sex samples and four
non-founders.  Usage: python make_x_set.py <dummy prefix> <out prefix>  (called by make_golden.sh)."""
import random
import sys

src, dst = sys.argv[1], sys.argv[2]
random.seed(5)
out = []
for k, ln in enumerate(open(src + ".bim").read().split("\n")[:-1]):
    f = ln.split("\t")
    f[0] = "1" if k < 300 else "X" if k < 550 else "Y" if k < 650 else "XY" if k < 700 else "MT"
    out.append("\t".join(f))
open(dst + ".bim", "w").write("\n".join(out) + "\n")
out = []
for k, ln in enumerate(open(src + ".fam").read().split("\n")[:-1]):
    f = ln.split("\t")
    f[4] = random.choice(["1", "1", "2", "2", "2", "0"]) if k % 17 else "0"
    if k in (5, 40, 77, 101):  # non-founders
        f[2], f[3] = "per0", "per1"
    out.append("\t".join(f))
open(dst + ".fam", "w").write("\n".join(out) + "\n")
open(dst + ".bed", "wb").write(open(src + ".bed", "rb").read())
