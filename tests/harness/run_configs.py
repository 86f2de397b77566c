"""BASELINE.json configs 2 and 4 at their stated sizes THROUGH THE PRODUCT (plink2_b200, process start to exit),
with the unmodified reference binary timed beside them on the same files (config 2: on a variant subset of the
same file, scaled by variant count - the full CPU run would take ~40 min on this box's 16-CPU quota; config 4:
the whole file).  Writes one JSON object per config to stdout / gpurun_out/configs_r02.json.

  python tests/harness/run_configs.py c2 [samples variants ref_variants]     (defaults 50000 500000 50000)
  python tests/harness/run_configs.py c4 [founders variants]                 (defaults 50000 1000000, 22 chromosomes)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench

REF = os.path.join(ROOT, "oracle", "_ref", "plink2")
BIN = os.path.join(ROOT, "plink_ng_b200", "plink2_b200")


def write_pgen(prefix, n, m, chrom_ct=1, ld_copy=0.0):
    """Fixed-width (mode 0x02) .pgen + .pvar + .psam of the bench generator's genotypes; with ld_copy > 0 the variants
    come in LD blocks of 8 (so --indep-pairwise has work)."""
    dev = "cuda"
    bpv = (n + 3) // 4
    t0 = time.perf_counter()
    with open(prefix + ".pgen", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x02]) + int(m).to_bytes(4, "little") + int(n).to_bytes(4, "little") + bytes([0x40]))
        for s0 in range(0, m, 8192):
            s1 = min(m, s0 + 8192)
            by = bench.synth_genovecs(torch, n, s0, s1, dev)[:, :bpv].contiguous()
            if ld_copy > 0:
                # LD blocks of 8 consecutive variants: members copy their block's first variant byte-wise (4 samples at a
                # time) with probability ld_copy -> r^2 ~ ld_copy^2 against the leader, ~ ld_copy^4 between members
                g = torch.Generator(device=dev)
                g.manual_seed(7919 + s0)
                mask = torch.rand((s1 - s0, bpv), generator=g, device=dev) < ld_copy
                lead = by[(torch.arange(s1 - s0, device=dev) // 8) * 8]
                by = torch.where(mask, lead, by)
            by.cpu().numpy().tofile(f)
    per_chr = -(-m // chrom_ct)
    with open(prefix + ".pvar", "w") as f:
        f.write("#CHROM\tPOS\tID\tREF\tALT\n")
        f.write("".join(f"{1 + k // per_chr}\t{1 + k % per_chr}\tsnp{k}\tA\tG\n" for k in range(m)))
    with open(prefix + ".psam", "w") as f:
        f.write("#IID\tSEX\n")
        f.write("".join(f"per{k}\t2\n" for k in range(n)))
    return time.perf_counter() - t0


def run(cmd, env=None):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    if r.returncode:
        raise RuntimeError(" ".join(cmd) + "\n" + r.stdout[-800:] + r.stderr[-800:])
    return dt, r


def c2(n=50000, m=500000, m_ref=50000):
    d = "/tmp/pl2_c2"
    os.makedirs(d, exist_ok=True)
    pre = os.path.join(d, "c2")
    gen_s = write_pgen(pre, n, m)
    cores = bench.effective_cores()
    pairs = n * (n - 1) // 2
    env = dict(os.environ, PL2_TIMING="1")
    # table filter that keeps ~2e5 of the 1.25e9 rows: kinship distribution of a 4,096-sample corner over the first
    # m_ref variants (through the library), tail quantile; for the full run the spread shrinks with sqrt(variants)
    import plink_ng_b200 as p
    from plink_ng_b200.host import KingJob
    mm = np.memmap(pre + ".pgen", dtype=np.uint8, mode="r", offset=12).reshape(m, (n + 3) // 4)
    corner = np.ascontiguousarray(mm[:m_ref, :1024]).view("<u8")
    with p.GpuContext(0) as cx, KingJob(cx, 4096) as job:
        job.add_variants(corner)
        kin = job.kinship()
    med, q = float(np.median(kin)), float(np.quantile(kin, 1 - 2e5 / pairs))
    thr_sub, thr_full = repr(q), repr(med + (q - med) / (m / m_ref) ** 0.5)
    del mm, corner, kin
    flags_full = ["--make-king-table", "counts", "--king-table-filter", thr_full]
    flags = ["--make-king-table", "counts", "--king-table-filter", thr_sub]
    t_ours, r = run([BIN, "--pfile", pre] + flags_full + ["--out", pre + "_b200"], env)
    phases = [ln for ln in r.stderr.split("\n") if ln.startswith("[timing]")]
    rows = sum(1 for _ in open(pre + "_b200.kin0")) - 1
    # reference on the first m_ref variants of the same file; ours on the same subset for the byte-for-byte comparison
    sub = ["--chr", "1", "--from-bp", "1", "--to-bp", str(m_ref)]
    t_ref, _ = run([REF, "--pfile", pre] + sub + flags + ["--threads", str(cores["threads_used"]), "--memory", "120000", "--out", pre + "_ref"])
    # ours cannot subset by position; write the subset file pair instead (same bytes: first m_ref records)
    with open(pre + ".pgen", "rb") as f, open(pre + "_sub.pgen", "wb") as g:
        hdr = bytearray(f.read(12))
        hdr[3:7] = int(m_ref).to_bytes(4, "little")
        g.write(hdr)
        g.write(f.read(m_ref * ((n + 3) // 4)))
    with open(pre + ".pvar") as f, open(pre + "_sub.pvar", "w") as g:
        for k, ln in enumerate(f):
            if k > m_ref:
                break
            g.write(ln)
    os.link(pre + ".psam", pre + "_sub.psam") if not os.path.exists(pre + "_sub.psam") else None
    t_ours_sub, _ = run([BIN, "--pfile", pre + "_sub"] + flags + ["--out", pre + "_b200sub"], env)
    same = open(pre + "_ref.kin0", "rb").read() == open(pre + "_b200sub.kin0", "rb").read()
    rows_sub = sum(1 for _ in open(pre + "_ref.kin0")) - 1
    out = {"config": "C2: --make-king-table, 50k samples x 500k SNPs, 1 B200", "samples": n, "variants": m, "input": "mode 0x02 .pgen, %.2f GB, generated in %.1f s" % (os.path.getsize(pre + ".pgen") / 1e9, gen_s),
           "b200_seconds_process": t_ours, "b200_pair_snp_per_s": pairs * m / t_ours, "b200_table_rows": rows, "b200_phases": phases,
           "reference": {"variants": m_ref, "seconds_process": t_ref, "pair_snp_per_s": pairs * m_ref / t_ref, "threads": cores["threads_used"], "host_cpus": cores,
                         "seconds_extrapolated_to_full": t_ref * m / m_ref, "note": "same file, first m_ref variants (--from-bp/--to-bp); scaled by variant count"},
           "same_subset": {"b200_seconds": t_ours_sub, "kin0_rows": rows_sub, "kin0_identical_to_reference": bool(same)},
           "speedup_process_vs_extrapolated_reference": (t_ref * m / m_ref) / t_ours}
    return out


def c4(n=50000, m=1000000):
    d = "/tmp/pl2_c4"
    os.makedirs(d, exist_ok=True)
    pre = os.path.join(d, "c4")
    gen_s = write_pgen(pre, n, m, chrom_ct=22, ld_copy=0.9)
    cores = bench.effective_cores()
    env = dict(os.environ, PL2_TIMING="1")
    flags = ["--indep-pairwise", "500", "50", "0.2"]
    t_ours, r = run([BIN, "--pfile", pre] + flags + ["--out", pre + "_b200"], env)
    phases = [ln for ln in r.stderr.split("\n") if ln.startswith("[timing]")]
    t_ref, _ = run([REF, "--pfile", pre] + flags + ["--threads", str(cores["threads_used"]), "--memory", "120000", "--out", pre + "_ref"])
    same_in = open(pre + "_ref.prune.in", "rb").read() == open(pre + "_b200.prune.in", "rb").read()
    same_out = open(pre + "_ref.prune.out", "rb").read() == open(pre + "_b200.prune.out", "rb").read()
    kept = sum(1 for _ in open(pre + "_ref.prune.in"))
    return {"config": "C4: --indep-pairwise 500 50 0.2, 50k founders x 1M SNPs (22 chromosomes)", "founders": n, "variants": m,
            "input": "mode 0x02 .pgen, %.2f GB, generated in %.1f s" % (os.path.getsize(pre + ".pgen") / 1e9, gen_s),
            "b200_seconds_process": t_ours, "b200_phases": phases, "reference_seconds_process": t_ref, "reference_threads": cores["threads_used"], "host_cpus": cores,
            "kept_variants": kept, "prune_in_identical": bool(same_in), "prune_out_identical": bool(same_out), "speedup_process": t_ref / t_ours}


def c3(n=100000, m=1000000, n_ref=8192, m_ref=32768):
    """Config 3's job at its stated size on ONE GPU (the 8-GPU row-block form is `--gpus 8`): --make-grm-bin, then
    --pca 20 approx (exact --pca is out of the reference's own reach at 100k samples: dsyevr on an 80 GB matrix).
    The reference (LAPACK build, threaded OpenBLAS) is timed on an n_ref x m_ref corner of the same file and
    scaled by pairs x variants (GRM) / samples x variants (PCA passes)."""
    d = os.environ.get("PL2_CFG_DIR", "/tmp/pl2_c3")
    os.makedirs(d, exist_ok=True)
    pre = os.path.join(d, "c3")
    gen_s = write_pgen(pre, n, m)
    cores = bench.effective_cores()
    env = dict(os.environ, PL2_TIMING="1")
    out = {"config": "C3: --make-grm-bin + --pca 20 (approx), 100k samples x 1M SNPs; single B200 here", "samples": n, "variants": m,
           "input": "mode 0x02 .pgen, %.2f GB, generated in %.1f s" % (os.path.getsize(pre + ".pgen") / 1e9, gen_s), "host_cpus": cores}
    t_grm, r = run([BIN, "--pfile", pre, "--make-grm-bin", "--out", pre + "_b200"], env)
    out["grm"] = {"b200_seconds_process": t_grm, "pair_snp_per_s": n * (n + 1) / 2 * m / t_grm, "phases": [ln for ln in r.stderr.split("\n") if ln.startswith("[timing]")],
                  "grm_bin_bytes": os.path.getsize(pre + "_b200.grm.bin"), "grm_N_bin_bytes": os.path.getsize(pre + "_b200.grm.N.bin")}
    # spot check of the written matrix: the n_ref-sample corner against the same job on the corner's own file
    corner = np.fromfile(pre + "_b200.grm.bin", dtype=np.float32, count=n_ref * (n_ref + 1) // 2)
    for suffix in (".grm.bin", ".grm.N.bin"):
        os.remove(pre + "_b200" + suffix)
    t_pca, r = run([BIN, "--pfile", pre, "--pca", "20", "approx", "--seed", "1", "--out", pre + "_b200"], env)
    out["pca"] = {"b200_seconds_process": t_pca, "phases": [ln for ln in r.stderr.split("\n") if ln.startswith("[timing]")], "eigenval_head": open(pre + "_b200.eigenval").read().split()[:5]}
    # reference on a corner of the same file: first n_ref samples, first m_ref variants
    with open(pre + "_keep.txt", "w") as f:
        f.write("".join(f"per{k}\n" for k in range(n_ref)))
    sub = ["--keep", pre + "_keep.txt", "--chr", "1", "--from-bp", "1", "--to-bp", str(m_ref), "--threads", str(cores["threads_used"]), "--memory", "120000"]
    lib_env = dict(os.environ)  # the LAPACK build finds the venv's OpenBLAS through its rpath
    t_ref_grm, _ = run([REF + "_lapack", "--pfile", pre] + sub + ["--make-grm-bin", "--out", pre + "_refgrm"], lib_env)
    t_ref_pca, _ = run([REF + "_lapack", "--pfile", pre] + sub + ["--pca", "20", "approx", "--seed", "1", "--out", pre + "_refpca"], lib_env)
    pairs_ref = n_ref * (n_ref + 1) / 2
    out["reference"] = {"binary": "oracle/_ref/plink2_lapack (threaded OpenBLAS)", "samples": n_ref, "variants": m_ref, "threads": cores["threads_used"],
                        "grm_seconds_process": t_ref_grm, "grm_pair_snp_per_s": pairs_ref * m_ref / t_ref_grm, "grm_seconds_extrapolated_to_full": t_ref_grm * (n * (n + 1) / 2 * m) / (pairs_ref * m_ref),
                        "pca_seconds_process": t_ref_pca, "pca_seconds_extrapolated_to_full": t_ref_pca * (n * m) / (n_ref * m_ref)}
    # ours on the SAME corner (own files) for the parity of the written GRM: corner of the big matrix == small job up to the allele frequencies
    # (the big job's frequencies come from all 100k samples, so this is a loose sanity bound, not a parity test - parity lives in tests/test_scale_gpu.py)
    ref_grm = np.fromfile(pre + "_refgrm.grm.bin", dtype=np.float32)
    out["grm_corner_vs_reference_corner_max_abs_diff"] = float(np.abs(corner - ref_grm).max())
    out["speedup_grm_process_vs_extrapolated_reference"] = out["reference"]["grm_seconds_extrapolated_to_full"] / t_grm
    out["speedup_pca_process_vs_extrapolated_reference"] = out["reference"]["pca_seconds_extrapolated_to_full"] / t_pca
    return out


def c5(n=200000, m=131072, pairs_checked=20000):
    """Config 5's per-GPU situation on ONE GPU: a KING job whose accumulators (20 bytes x N(N-1)/2 = 400 GB at 200k
    samples) do not fit the device, so the host program makes several passes over the file on its own (no --gpu-memory).
    Parity at this size: the rows our dense multipass run reports above a kinship threshold are recomputed by the
    REFERENCE through its pair-list path (--king-table-subset on our table, same file) and compared byte for byte."""
    d = os.environ.get("PL2_CFG_DIR", "/tmp/pl2_c5")
    os.makedirs(d, exist_ok=True)
    pre = os.path.join(d, "c5")
    gen_s = write_pgen(pre, n, m)
    cores = bench.effective_cores()
    env = dict(os.environ, PL2_TIMING="1")
    import plink_ng_b200 as p
    from plink_ng_b200.host import KingJob
    mm = np.memmap(pre + ".pgen", dtype=np.uint8, mode="r", offset=12).reshape(m, (n + 3) // 4)
    corner = np.ascontiguousarray(mm[:, :1024]).view("<u8")
    with p.GpuContext(0) as cx, KingJob(cx, 4096) as job:
        job.add_variants(corner)
        kin = job.kinship()
    pairs = n * (n - 1) // 2
    # ~pairs_checked rows expected above the threshold under a normal approximation of the kinship distribution
    from statistics import NormalDist
    z = NormalDist().inv_cdf(1.0 - pairs_checked / pairs)
    thr = repr(float(np.median(kin) + z * kin.std()))
    del mm, corner, kin
    flags = ["--make-king-table", "counts", "--king-table-filter", thr]
    # NSNP as the plain dense count: the reference's pair-list path has no rare-variant pre-scan, so its NSNP lacks the
    # +1 quirk of its own dense path that the host program otherwise reproduces
    t_ours, r = run([BIN, "--pfile", pre] + flags + ["--out", pre + "_b200"], dict(env, PL2_KING_DENSE_NSNP="1"))
    phases = [ln for ln in r.stderr.split("\n") if ln.startswith("[timing]")]
    passes = max([int(ln.split("pass ")[1].split("/")[1].split(":")[0]) for ln in (r.stdout + r.stderr).replace("\r", "\n").split("\n") if "--make-king-table pass " in ln] or [1])
    rows = sum(1 for _ in open(pre + "_b200.kin0")) - 1
    t_ref, _ = run([REF, "--pfile", pre, "--make-king-table", "counts", "--king-table-subset", pre + "_b200.kin0", "--threads", str(cores["threads_used"]), "--memory", "120000", "--out", pre + "_ref"])
    same = open(pre + "_ref.kin0", "rb").read() == open(pre + "_b200.kin0", "rb").read()
    return {"config": "C5 share: --make-king-table, 200k samples x 131,072 SNPs on one B200 (accumulators 400 GB > HBM: natural multipass)", "samples": n, "variants": m,
            "input": "mode 0x02 .pgen, %.2f GB, generated in %.1f s" % (os.path.getsize(pre + ".pgen") / 1e9, gen_s), "passes": passes, "b200_seconds_process": t_ours, "b200_pair_snp_per_s": pairs * m / t_ours,
            "b200_phases": phases[-12:], "kinship_threshold": thr, "table_rows": rows,
            "reference_pair_list": {"command": "--make-king-table counts --king-table-subset <our table>", "seconds_process": t_ref, "threads": cores["threads_used"], "rows_identical_to_ours": bool(same)}, "host_cpus": cores}


if __name__ == "__main__":
    which = sys.argv[1]
    args = [int(x) for x in sys.argv[2:]]
    res = {"c2": c2, "c3": c3, "c4": c4, "c5": c5}[which](*args)
    print(json.dumps(res))
