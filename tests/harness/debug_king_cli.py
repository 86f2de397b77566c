"""Triage: reference vs plink2_b200 on --dummy data, reports WHERE the kinship matrix / table differ."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF, BIN = os.path.join(ROOT, "oracle/_ref/plink2"), os.path.join(ROOT, "plink_ng_b200/plink2_b200")
n, m = int(sys.argv[1]), int(sys.argv[2])
extra = sys.argv[3:]
d = tempfile.mkdtemp(prefix="dbg_")
pre = os.path.join(d, "d")
subprocess.run([REF, "--dummy", str(n), str(m), "0.02", "--seed", "9", "--threads", "4", "--make-bed", "--out", pre], check=True, capture_output=True)
flags = ["--make-king", "bin", "triangle", "--make-king-table", "counts", "--king-table-filter", "-0.05"]
subprocess.run([REF, "--bfile", pre] + flags + ["--threads", "16", "--out", pre + "_ref"], check=True, capture_output=True)
for rep in range(int(os.environ.get('REPS', '3'))):
    r = subprocess.run([BIN, "--bfile", pre] + flags + extra + ["--out", pre + "_b"], capture_output=True, text=True)
    print("rc", r.returncode, [ln for ln in r.stdout.split("\n") if "passes" in ln or "Error" in ln])
    a = np.fromfile(pre + "_ref.king.bin", dtype=np.float64)
    b = np.fromfile(pre + "_b.king.bin", dtype=np.float64)
    bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b)))) if a.size == b.size else None
    print(f"rep {rep}: sizes {a.size} {b.size}; differing entries: {None if bad is None else bad.size}")
    if bad is not None and bad.size:
        rows = np.floor((1 + np.sqrt(1 + 8 * bad.astype(np.float64))) / 2).astype(np.int64)
        cols = bad - rows * (rows - 1) // 2
        print("  rows min/max", rows.min(), rows.max(), "cols min/max", cols.min(), cols.max(), "row tiles", np.unique(rows // 128)[:20], "col tiles(80)", np.unique(cols // 80)[:30])
        print("  sample:", [(int(r_), int(c_), float(a[i]), float(b[i])) for i, r_, c_ in list(zip(bad, rows, cols))[:5]])
    ta, tb = open(pre + "_ref.kin0").read().split("\n"), open(pre + "_b.kin0").read().split("\n")
    nd = sum(1 for x, y in zip(ta, tb) if x != y)
    print(f"  kin0 lines {len(ta)} {len(tb)} differing {nd}")
    if nd:
        k = next(i for i, (x, y) in enumerate(zip(ta, tb)) if x != y)
        print("   ", ta[k], "|", tb[k])
