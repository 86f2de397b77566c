"""--score fixture for set A: ~400 (variant, allele, weight) lines in shuffled order - REF and ALT alleles, a few unknown
IDs and foreign allele codes, a header line."""
import random
import sys

bim = [ln.split() for ln in open(sys.argv[1])]
random.seed(12)
rows = []
for k, b in enumerate(bim):
    if k % 5 in (0, 3):
        allele = b[4] if k % 2 else b[5]
        rows.append((b[1], allele, "%.5g" % random.gauss(0, 0.1)))
    elif k % 97 == 1:
        rows.append((b[1], "Q", "0.5"))      # allele code not in the dataset
rows += [("nosuch%d" % k, "A", "0.1") for k in range(5)]
random.shuffle(rows)
with open(sys.argv[2], "w") as f:
    f.write("SNP\tA1\tBETA\n")
    f.write("".join("\t".join(r) + "\n" for r in rows))
