"""Coarse device-side timing of the two KING kernels (development aid, not bench.py)."""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plink_ng_b200 as p
from plink_ng_b200.host import KING_ALGO_POPCOUNT, KING_ALGO_TENSOR, KING_ALGO_TENSOR_TS, KingJob

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
words = (n + 31) // 32
g = torch.randint(0, 256, (m, words * 8), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
pairs = n * (n - 1) // 2
with p.GpuContext(0) as ctx:
    for name, algo in (("tensor", KING_ALGO_TENSOR), ("tensor_ts", KING_ALGO_TENSOR_TS), ("popcount", KING_ALGO_POPCOUNT))[: (2 if os.environ.get("SKIP_POPC") else 3)][(1 if os.environ.get("SKIP_SS") else 0):]:
        with KingJob(ctx, n, 0, n, algo) as job:
            job.add_variants_device(g.data_ptr(), words * 8, m)  # warm-up
            ctx.synchronize()
            ctx.event_record(0)
            for _ in range(reps):
                job.add_variants_device(g.data_ptr(), words * 8, m)
            ctx.event_record(1)
            ms = ctx.event_elapsed_ms(0, 1) / reps
            print(f"{name:9s} N={n} M={m}: {ms:9.3f} ms/batch  {pairs * m / ms / 1e9:10.2f} G pair*SNP/ms->{pairs * m / (ms * 1e-3):.3e} pair*SNP/s  int8-equiv {5 * 2 * pairs * m / (ms * 1e-3) / 1e12:8.1f} TOP/s", flush=True)

# GRM kernel timing (same shape)
if os.environ.get("SKIP_GRM"):
    raise SystemExit(0)
from plink_ng_b200.host import GrmJob
rf = np.random.default_rng(0).uniform(0.05, 0.95, size=m)
with p.GpuContext(0) as ctx, GrmJob(ctx, n) as job:
    job.add_variants_device(g.data_ptr(), words * 8, m, ref_freqs=rf)
    ctx.synchronize()
    ctx.event_record(0)
    for _ in range(reps):
        job.add_variants_device(g.data_ptr(), words * 8, m, ref_freqs=rf)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1) / reps
    tri = n * (n + 1) // 2
    print(f"grm       N={n} M={m}: {ms:9.3f} ms/batch  int8 {11 * 2 * tri * m / (ms * 1e-3) / 1e12:8.1f} TOP/s (11 products: 10 digit planes + obs)", flush=True)
