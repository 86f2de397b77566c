"""Throughput of --variant-score on the tensor tile path (pl2gpu_pca_vscore over a resident block)."""
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

n, m, cols = 100000, 131072, 40
g = bench.synth_genovecs(torch, n, 0, m, torch.device("cuda", 0))
w = np.random.default_rng(0).normal(size=(n, cols))
with p.GpuContext(0) as ctx:
    h = C.c_void_p()
    check(lib.pl2gpu_pca_begin_shard(ctx.handle, n, m, 1, C.byref(h)), "begin")
    check(lib.pl2gpu_pca_add_variants(h, C.c_void_p(g.data_ptr()), g.shape[1], m, 1, None), "add")
    out = np.empty((m, cols))
    for rep in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        check(lib.pl2gpu_pca_vscore(h, w.ctypes.data, cols, out.ctypes.data), "vscore")
        dt = time.perf_counter() - t0
    lib.pl2gpu_pca_end(h)
    print(f"variant scores {n} samples x {m} variants x {cols} weight columns: {dt * 1e3:.1f} ms per call (incl. weight upload, digit planes, two precision passes, {out.nbytes / 1e6:.0f} MB of results to the host) = "
          f"{n * m * cols / dt:.3e} sample-variant-columns/s, {n * m / 4 / dt / 1e9:.0f} GB/s of 2-bit genotypes")
