"""Development aid / measurement: --indep-pairwise pair-decision kernel (pl2gpu_ld_band_flags) at a C4-like shape.
Prints pairs/s, sample-pairs/s and the POPC-pipe fraction (7 popc32 per pair and 32 founders; XU pipe = 16 lanes/clk/SM)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import plink_ng_b200 as p
from plink_ng_b200.capi import check, lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
window = int(sys.argv[3]) if len(sys.argv) > 3 else 500
band = window - 1
dev = torch.device("cuda", 0)
g = bench.synth_genovecs(torch, n, 0, m, dev)
row_bytes = g.shape[1]
torch.cuda.synchronize()
flags_t = torch.zeros((m, band), dtype=torch.uint8).pin_memory()  # pinned: the D2H of the decision bytes is not the pageable-copy bottleneck
flags = flags_t.numpy()
with p.GpuContext(0) as ctx:
    for rep in range(2):
        t0 = time.perf_counter()
        check(lib.pl2gpu_ld_band_flags(ctx.handle, C.c_void_p(g.data_ptr()), row_bytes, n, m, 1, band, 0.2 * (1 + 2.0 ** -44), flags.ctypes.data), "pl2gpu_ld_band_flags")
        dt = time.perf_counter() - t0
    pairs = m * band - band * (band + 1) // 2
    words = (n + 31) // 32
    popc = 7 * pairs * words
    peak = 148 * 16 * 1.965e9
    tops = 12 * pairs * n / dt / 1e12  # 6 int8 products x 2 ops per pair and founder (tensor kernel)
    print(f"ld_band_flags founders={n} variants={m} window={window} algo={os.environ.get('PL2_LD_ALGO', 'tensor')}: {dt * 1e3:.1f} ms  {pairs / dt:.3e} pairs/s  {pairs * n / dt:.3e} sample-pairs/s  "
          f"int8-equivalent {tops:.0f} TOP/s; popcount-equivalent {popc / dt:.3e} popc32/s = {100 * popc / dt / peak:.1f}% of the XU pipe ({peak:.2e}/s); flagged {flags.mean() * 100:.2f}% of pairs; "
          f"whole call incl. D2D staging and D2H of {flags.nbytes / 1e6:.0f} MB of decisions (pinned)")
