"""Throughput of the --score accumulation kernel (pl2gpu_score_add_variants, inputs resident on the device)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import plink_ng_b200 as p
from plink_ng_b200.capi import check, lib

n, m = 100000, 131072
g = bench.synth_genovecs(torch, n, 0, m, torch.device("cuda", 0))
w4 = np.random.default_rng(0).normal(size=(m, 4))
d4 = np.full(m, 0 | (1 << 2) | (2 << 4), dtype=np.uint8)
with p.GpuContext(0) as ctx:
    h = C.c_void_p()
    check(lib.pl2gpu_score_begin(ctx.handle, n, C.byref(h)), "begin")
    for rep in range(3):
        ctx.synchronize()
        ctx.event_record(0)
        check(lib.pl2gpu_score_add_variants(h, C.c_void_p(g.data_ptr()), g.shape[1], m, 1, w4.ctypes.data, d4.ctypes.data), "add")
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1)
    lib.pl2gpu_score_end(h)
    gb = n * m / 4 / 1e9
    print(f"score accumulation {n} samples x {m} entries: {ms:.1f} ms per call = {gb / (ms * 1e-3):.0f} GB/s of 2-bit genotypes ({n * m / (ms * 1e-3):.3e} sample-entries/s), incl. the D2D staging copy")
