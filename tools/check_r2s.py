#!/usr/bin/env python
"""Round-2 check S (one GPU, ~1 minute): the filter view and the relatedness-prune chaining through the DEVICE
commands, compared with files the reference wrote for the same command lines (tests/golden/g_*, make_golden.sh).
No pytest / torch import, so it fits a very short GPU slot.  Writes gpurun_out/r2s_check.txt; exit code = failures."""
import gzip
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "plink_ng_b200", "plink2_b200")
OUT = os.path.join(ROOT, "gpurun_out", "r2s")
os.makedirs(OUT, exist_ok=True)
ENV = dict(os.environ, CUDA_VISIBLE_DEVICES="0")
F = ["--keep", "x_keep1.txt", "x_keep2.txt", "--remove", "x_remove.txt", "--extract", "x_extract.txt", "--exclude", "x_exclude.txt"]
report = []
fails = 0


def run(name, args):
    t0 = time.time()
    r = subprocess.run([BIN] + args + ["--out", os.path.join(OUT, name)], capture_output=True, text=True, env=ENV, cwd=GD, timeout=120)
    report.append("%s rc=%d %.1fs" % (name, r.returncode, time.time() - t0))
    if r.returncode:
        report.append((r.stdout + r.stderr)[-600:])
    return r.returncode == 0


def gold(name):
    p = os.path.join(GD, name)
    return gzip.open(p, "rb").read() if name.endswith(".gz") else open(p, "rb").read()


def check(label, ok, detail=""):
    global fails
    report.append("  %-46s %s %s" % (label, "ok" if ok else "MISMATCH", detail))
    fails += 0 if ok else 1


def same_bytes(label, name, ext, gold_name):
    got = open(os.path.join(OUT, name + ext), "rb").read()
    check(label, got == gold(gold_name))


def table_close(label, name, ext, gold_name, tol):
    got = open(os.path.join(OUT, name + ext)).read().split("\n")
    want = gold(gold_name).decode().split("\n")
    if len(got) != len(want):
        return check(label, False, "line count %d vs %d" % (len(got), len(want)))
    worst = 0.0
    for a, b in zip(got, want):
        ta, tb = a.split(), b.split()
        if len(ta) != len(tb):
            return check(label, False, "token count")
        for x, y in zip(ta, tb):
            if x == y:
                continue
            try:
                if abs(float(x) - float(y)) <= 1e-9:  # a sum that is exactly 0 in exact arithmetic prints as 1e-16 noise
                    continue
                worst = max(worst, abs(float(x) - float(y)) / max(1e-300, abs(float(y))))
            except ValueError:
                return check(label, False, "%s vs %s" % (x, y))
    check(label, worst <= tol, "worst rel %.2e" % worst)


try:
    if run("g1", ["--bfile", "x"] + F + ["--make-king-table"]):
        same_bytes("filters -> KING table", "g1", ".kin0", "g_xfilt.kin0.gz")
    if run("g2", ["--bfile", "x"] + F + ["--indep-pairwise", "50", "5", "0.2"]):
        same_bytes("filters -> LD prune (chrX, founders of view)", "g2", ".prune.in", "g_xfilt.prune.in")
    if run("g3", ["--bfile", "x"] + F + ["--freq"]):
        table_close("filters -> --freq", "g3", ".afreq", "g_xfilt.afreq", 1e-5)
    if run("g4", ["--bfile", "a", "--king-cutoff", "0.02", "--indep-pairwise", "50", "5", "0.2"]):
        same_bytes("--king-cutoff -> LD prune (frozen freqs)", "g4", ".prune.in", "g_acut.prune.in")
        same_bytes("--king-cutoff list", "g4", ".king.cutoff.in.id", "a_cut.king.cutoff.in.id")
    if run("g5", ["--bfile", "a", "--king-cutoff", "0.02", "--make-grm-bin"]):
        a = np.fromfile(os.path.join(OUT, "g5.grm.bin"), dtype=np.float32)
        b = np.frombuffer(gold("g_acut.grm.bin"), dtype=np.float32)
        ok = a.size == b.size and bool(np.allclose(a, b, rtol=2e-7, atol=1e-10))
        check("--king-cutoff -> GRM (frozen freqs)", ok, "n=%d" % a.size)
    sub = os.path.join(OUT, "in.kin0")
    open(sub, "wb").write(gold("a_kingp.kin0.gz"))
    if run("g6", ["--bfile", "a", "--king-cutoff-table", sub, "0.02", "--pca", "3"]):
        a = np.loadtxt(os.path.join(OUT, "g6.eigenval"))
        b = np.loadtxt(os.path.join(GD, "g_akct.eigenval"))
        check("--king-cutoff-table -> PCA eigenvalues", bool(np.allclose(a, b, rtol=2e-5)), str(a))
        va = np.loadtxt(os.path.join(OUT, "g6.eigenvec"), skiprows=1, usecols=(2, 3, 4))
        vb = np.loadtxt(os.path.join(GD, "g_akct.eigenvec"), skiprows=1, usecols=(2, 3, 4))
        sgn = np.sign((va * vb).sum(axis=0))
        check("--king-cutoff-table -> PCA eigenvectors", va.shape == vb.shape and bool(np.allclose(va * sgn, vb, atol=2e-5)))
    if run("g7", ["--pgen", "a_mode10.pgen", "--pvar", "a.pvar", "--psam", "a.psam", "--remove", "x_remove.txt", "--exclude", "x_exclude.txt", "--make-king-table"]):
        same_bytes(".pgen 0x10 + filters -> KING table", "g7", ".kin0", "g_afilt.kin0.gz")
    if run("g8", ["--bfile", "a", "--king-cutoff", "0.02", "--score", "a_score.txt", "header", "cols=+scoresums,+denom"]):
        table_close("--king-cutoff -> --score (frozen freqs)", "g8", ".sscore", "g_acut.sscore", 2e-6)
    # regressions for the FID-0 rule of IID-only files on an all-FID-0 dataset (behaviour must not have changed)
    if run("g9", ["--bfile", "a", "--make-king-table", "counts", "cols=+ibs1", "--king-table-subset", "a_sub2.txt"]):
        same_bytes("IID-only --king-table-subset (set A)", "g9", ".kin0", "a_kingsub2.kin0")
    if run("g10", ["--bfile", "a", "--variant-score", "a_vscore_weights.txt"]):
        table_close("--variant-score weights file (set A)", "g10", ".vscore", "a_vs.vscore", 2e-6)
except Exception as e:  # keep whatever was learned
    report.append("EXCEPTION %r" % (e,))
    fails += 1
report.append("failures: %d" % fails)
open(os.path.join(ROOT, "gpurun_out", "r2s_check.txt"), "w").write("\n".join(report) + "\n")
print("\n".join(report))
sys.exit(fails)
