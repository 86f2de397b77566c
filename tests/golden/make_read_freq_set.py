"""--read-freq fixture: a.afreq with a third of the frequencies perturbed, some lines dropped, some REF/ALT pairs
swapped (frequency of the file's ALT = the dataset's REF), a mismatching allele code, OBS_CT = 0 and nan entries."""
import random
import sys

src, dst = sys.argv[1], sys.argv[2]
random.seed(3)
lines = open(src).read().split("\n")
out = [lines[0]]
for k, ln in enumerate(lines[1:]):
    if not ln:
        continue
    f = ln.split("\t")
    col = {name: i for i, name in enumerate(lines[0].lstrip("#").split("\t"))}
    r = k % 11
    if r == 0:
        continue  # absent from the file -> frequency computed from the data
    if r in (1, 2, 3):
        f[col["ALT_FREQS"]] = "%.6g" % min(0.98, max(0.02, float(f[col["ALT_FREQS"]]) * random.uniform(0.5, 1.5)))
    elif r == 4:
        f[col["REF"]], f[col["ALT"]] = f[col["ALT"]], f[col["REF"]]
        f[col["ALT_FREQS"]] = "%.6g" % random.uniform(0.05, 0.95)
    elif r == 5 and k % 3 == 0:
        f[col["ALT"]] = "Q"  # allele code not in the dataset -> skipped
    elif r == 6 and k % 3 == 0:
        f[col["OBS_CT"]] = "0"
    elif r == 7 and k % 3 == 0:
        f[col["ALT_FREQS"]] = "nan"
    out.append("\t".join(f))
open(dst, "w").write("\n".join(out) + "\n")
