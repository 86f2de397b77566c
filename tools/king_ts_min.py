"""Smallest possible run of the default KING kernel (for compute-sanitizer / quick triage)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plink_ng_b200 as p
from plink_ng_b200.host import KING_ALGO_TENSOR_TS, KING_ALGO_POPCOUNT, KingJob, GrmJob, pack_genotypes
n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 257, int(sys.argv[2]) if len(sys.argv) > 2 else 300
what = sys.argv[3] if len(sys.argv) > 3 else "king"
rng = np.random.default_rng(0)
geno = rng.integers(0, 4, size=(m, n), dtype=np.uint8)
with p.GpuContext(0) as ctx:
    if what == "king":
        res = []
        for algo in (KING_ALGO_TENSOR_TS, KING_ALGO_POPCOUNT):
            with KingJob(ctx, n, 0, n, algo) as job:
                job.add_variants(pack_genotypes(geno))
                job.add_variants(pack_genotypes(geno[: m // 2]))
                res.append(job.counts())
        print("king ts == popcount:", np.array_equal(res[0], res[1]))
    else:
        with GrmJob(ctx, n) as job:
            job.add_variants(pack_genotypes(geno))
            g = job.rows()
        print("grm ok", float(np.abs(g).max()))
