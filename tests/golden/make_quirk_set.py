"""Set Q: 150 samples x 140 variants built to exercise the reference's rare-variant KING pre-scan
(CalcKingSparseThread, 2.0/plink2_matrix_calc.cc:904-1250): half the variants are common (HWE, 5 % missing),
half carry at most 4 non-common genotypes (het / other homozygote / missing, either allele common), so several
pairs hit the (other homozygote, missing) case in which the reference's NSNP exceeds the dense count by one.
Writes q.bed / q.bim / q.fam next to this script (PLINK 1 .bed coding: 0 hom-A1(ALT), 1 missing, 2 het, 3 hom-A2(REF))."""
import os
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20260923)
n, m_common, m_rare = 150, 70, 70
g = np.zeros((m_common + m_rare, n), dtype=np.uint8)  # PgrGet codes: 0 hom-REF, 1 het, 2 hom-ALT, 3 missing
for v in range(m_common):
    f = rng.uniform(0.1, 0.9)
    g[v] = (rng.random(n) < f).astype(np.uint8) + (rng.random(n) < f).astype(np.uint8)
    g[v][rng.random(n) < 0.05] = 3
for v in range(m_common, m_common + m_rare):
    common = 0 if v % 3 else 2
    g[v] = common
    k = rng.integers(1, 5)  # 1..4 rare genotypes (max_sparse_ct = 150 / 33 = 4)
    idx = rng.choice(n, size=k, replace=False)
    g[v][idx] = rng.choice([1, 2 - common, 3, 3], size=k)
order = rng.permutation(m_common + m_rare)  # interleave common and rare variants
g = g[order]
lut = np.array([3, 2, 0, 1], dtype=np.uint8)  # PgrGet code -> .bed code
b = lut[g]
pad = (-n) % 4
b = np.concatenate([b, np.zeros((b.shape[0], pad), dtype=np.uint8)], axis=1).reshape(b.shape[0], -1, 4)
by = (b[:, :, 0] | (b[:, :, 1] << 2) | (b[:, :, 2] << 4) | (b[:, :, 3] << 6)).astype(np.uint8)
with open(os.path.join(here, "q.bed"), "wb") as f:
    f.write(bytes([0x6C, 0x1B, 0x01]))
    f.write(by.tobytes())
with open(os.path.join(here, "q.bim"), "w") as f:
    f.write("".join(f"1\tq{k}\t0\t{k + 1}\tA\tG\n" for k in range(g.shape[0])))
with open(os.path.join(here, "q.fam"), "w") as f:
    f.write("".join(f"fam{k % 7}\tid{k}\t0\t0\t{1 + k % 2}\t-9\n" for k in range(n)))
