"""Phase timing of `--pca approx` core (PL2_TIMING=1 makes pl2gpu_pca_run print its phases)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PL2_TIMING"] = "1"
import numpy as np, torch
import bench, plink_ng_b200 as p
from plink_ng_b200.capi import check, lib
n, m, k = (int(x) for x in (sys.argv[1:4] + ["16384", "65536", "20"][len(sys.argv) - 1:]))
g = bench.synth_genovecs(torch, n, 0, m, torch.device("cuda", 0))
g1 = np.random.default_rng(1).standard_normal((n, 2 * k))
with p.GpuContext(0) as ctx:
    h = C.c_void_p()
    check(lib.pl2gpu_pca_begin(ctx.handle, n, m, k, C.byref(h)), "begin")
    check(lib.pl2gpu_pca_add_variants(h, C.c_void_p(g.data_ptr()), g.shape[1], m, 1, None), "add")
    vals, vecs = np.empty(k), np.empty((k, n))
    t0 = time.perf_counter()
    check(lib.pl2gpu_pca_run(h, g1.ctypes.data, vals.ctypes.data, vecs.ctypes.data), "run")
    print(f"pca_run {n} x {m} k={k}: {time.perf_counter() - t0:.2f} s; eigenvalues {vals[:3]}")
    lib.pl2gpu_pca_end(h)
