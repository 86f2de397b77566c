"""Development aid: time the exact-PCA eigensolve (one-sided Jacobi, jacobi.cuh) and check it against numpy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plink_ng_b200 as p
from plink_ng_b200.host import GrmJob, pack_genotypes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = 10
rng = np.random.default_rng(5)
pops = 6
f0 = rng.uniform(0.05, 0.95, m)
fp = np.clip(f0[None, :] + 0.12 * rng.standard_normal((pops, m)), 0.02, 0.98)
lab = rng.integers(0, pops, n)
geno = (rng.random((m, n)) < fp[lab].T).astype(np.uint8) + (rng.random((m, n)) < fp[lab].T).astype(np.uint8)
with p.GpuContext(0) as ctx, GrmJob(ctx, n) as job:
    job.add_variants(pack_genotypes(geno))
    g = job.rows()
    full = np.tril(g) + np.tril(g, -1).T
    t0 = time.perf_counter()
    vals, vecs = job.eigen_topk(k)
    dt = time.perf_counter() - t0
    l0 = ctx.launch_count()
w, v = np.linalg.eigh(full)
w = w[::-1][:k]; v = v[:, ::-1][:, :k].T
err_v = max(min(np.abs(vecs[i] - v[i]).max(), np.abs(vecs[i] + v[i]).max()) for i in range(k))
print(f"eigen_topk n={n}: {dt:.3f} s  eigval rel err {np.max(np.abs(vals - w) / np.abs(w)):.2e}  eigvec max err (sign-aligned) {err_v:.2e}  launches so far {l0}")
