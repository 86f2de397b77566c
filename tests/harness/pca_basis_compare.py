"""Development aid: `--pca approx` core against the numpy restatement for the two ways of building the orthonormal
basis of the Krylov matrix (PL2_PCA_BASIS=jacobi|bcgs) and the two final stages (PL2_PCA_FINAL=gram|jacobi)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import plink_ng_b200 as p
from oracle import plink_oracle as orc
from plink_ng_b200.host import pack_genotypes, pca_approx
from test_pca_gpu import _structured_geno

cases = [(1100, 9000, 20, 21, 8, 0.12, 3), (4000, 12000, 20, 5, 8, 0.1, 4)]
with p.GpuContext(0) as ctx:
    for n, m, k, seed, pops, fst, gseed in cases:
        geno = _structured_geno(m, n, seed=seed, pops=pops, fst=fst)
        g1 = np.random.default_rng(gseed).standard_normal((n, 2 * k))
        want_vals, want_vecs = orc.pca_approx(geno, k, g1)
        packed = pack_genotypes(geno)
        for basis in ("jacobi", "bcgs"):
            for final in ("gram", "jacobi"):
                os.environ["PL2_PCA_BASIS"] = basis
                os.environ["PL2_PCA_FINAL"] = final
                vals, vecs = pca_approx(ctx, packed, n, k, g1)
                rel = np.abs(vals - want_vals) / want_vals
                sg = np.sign(np.sum(vecs * want_vecs, axis=1, keepdims=True))
                verr = np.abs(vecs * sg - want_vecs).max(axis=1) / np.abs(want_vecs).max()
                top = pops - 1
                print(f"n={n} m={m} k={k} basis={basis} final={final}: eigenvalue rel err structure PCs {rel[:top].max():.2e}, trailing {rel[top:].max():.2e}; "
                      f"eigenvector err (rel. to largest component) structure {verr[:top].max():.2e}, trailing {verr[top:].max():.2e}", flush=True)
