"""GPU: the sample / variant filter view and the relatedness-prune chaining through the device commands.  Every
expected file was written by the reference binary for the SAME command line (tests/golden/g_*, make_golden.sh); the
host-side halves (filter lists, view decode, frozen-frequency semantics) are pinned on the CPU in test_host_program.py
and test_oracle_golden.py."""
import gzip
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("PL2_TEST_BIN", os.path.join(ROOT, "plink_ng_b200", "plink2_b200"))
ENV = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
X_FILTERS = ["--keep", "x_keep1.txt", "x_keep2.txt", "--remove", "x_remove.txt", "--extract", "x_extract.txt", "--exclude", "x_exclude.txt"]


def _run(golden_dir, tmp_path, args, name="o"):
    out = str(tmp_path / name)
    r = subprocess.run([BIN] + args + ["--out", out], capture_output=True, text=True, env=ENV, cwd=golden_dir)
    assert r.returncode == 0, r.stdout + r.stderr
    return out


def _gold(golden_dir, name):
    p = os.path.join(golden_dir, name)
    return gzip.open(p, "rb").read() if name.endswith(".gz") else open(p, "rb").read()


def _tables_close(got_path, want_bytes, rtol, atol):
    got = open(got_path).read().split("\n")
    want = want_bytes.decode().split("\n")
    assert len(got) == len(want)
    for a, b in zip(got, want):
        ta, tb = a.split("\t"), b.split("\t")
        assert len(ta) == len(tb), (a, b)
        for x, y in zip(ta, tb):
            if x != y:
                assert np.isclose(float(x), float(y), rtol=rtol, atol=atol), (a, b)


def test_filtered_view_king_table_byte_identical(golden_dir, tmp_path):
    """--keep x2 / --remove / --extract / --exclude on set X (60 of 120 samples, 420 of 800 variants), then the KING table."""
    out = _run(golden_dir, tmp_path, ["--bfile", "x"] + X_FILTERS + ["--make-king-table"])
    assert open(out + ".kin0", "rb").read() == _gold(golden_dir, "g_xfilt.kin0.gz")


def test_filtered_view_ld_prune_and_freq(golden_dir, tmp_path):
    """LD prune and --freq decode the FOUNDERS of the filtered view (58 of its 60 samples), chrX / chrY / MT included."""
    out = _run(golden_dir, tmp_path, ["--bfile", "x"] + X_FILTERS + ["--indep-pairwise", "50", "5", "0.2"])
    assert open(out + ".prune.in", "rb").read() == _gold(golden_dir, "g_xfilt.prune.in")
    out = _run(golden_dir, tmp_path, ["--bfile", "x"] + X_FILTERS + ["--freq"], name="f")
    _tables_close(out + ".afreq", _gold(golden_dir, "g_xfilt.afreq"), 1e-5, 1e-9)


def test_filtered_view_of_ld_compressed_pgen(golden_dir, tmp_path):
    out = _run(golden_dir, tmp_path, ["--pgen", "a_mode10.pgen", "--pvar", "a.pvar", "--psam", "a.psam", "--remove", "x_remove.txt", "--exclude", "x_exclude.txt", "--make-king-table"])
    assert open(out + ".kin0", "rb").read() == _gold(golden_dir, "g_afilt.kin0.gz")


def test_king_cutoff_feeds_grm_and_score_with_frozen_frequencies(golden_dir, tmp_path):
    """`--king-cutoff 0.02` then `--make-grm-bin` / `--score` in one run: 49 survivors, allele frequencies from before
    the prune (the reference's order of operations, plink2.cc:2280-2304 then :2580)."""
    out = _run(golden_dir, tmp_path, ["--bfile", "a", "--king-cutoff", "0.02", "--make-grm-bin"])
    assert open(out + ".king.cutoff.in.id", "rb").read() == _gold(golden_dir, "a_cut.king.cutoff.in.id")
    got = np.fromfile(out + ".grm.bin", dtype=np.float32)
    want = np.frombuffer(_gold(golden_dir, "g_acut.grm.bin"), dtype=np.float32)
    assert got.shape == want.shape == (49 * 50 // 2,)
    assert np.allclose(got, want, rtol=2e-7, atol=1e-10)  # fp64 sums in another order, rounded to fp32
    # frequencies re-estimated from the survivors would be off by ~1e-3
    out = _run(golden_dir, tmp_path, ["--bfile", "a", "--king-cutoff", "0.02", "--score", "a_score.txt", "header", "cols=+scoresums,+denom"], name="s")
    _tables_close(out + ".sscore", _gold(golden_dir, "g_acut.sscore"), 2e-5, 2e-9)


def test_king_cutoff_table_feeds_pca(golden_dir, tmp_path):
    """File-driven prune (host) then exact PCA of the 50 survivors (device), frequencies frozen before the prune."""
    sub = tmp_path / "in.kin0"
    sub.write_bytes(_gold(golden_dir, "a_kingp.kin0.gz"))
    out = _run(golden_dir, tmp_path, ["--bfile", "a", "--king-cutoff-table", str(sub), "0.02", "--pca", "3"])
    vals = np.loadtxt(out + ".eigenval")
    assert np.allclose(vals, np.loadtxt(os.path.join(golden_dir, "g_akct.eigenval")), rtol=2e-5)
    got = np.loadtxt(out + ".eigenvec", skiprows=1, usecols=(2, 3, 4))
    want = np.loadtxt(os.path.join(golden_dir, "g_akct.eigenvec"), skiprows=1, usecols=(2, 3, 4))
    assert got.shape == want.shape == (50, 3)
    sign = np.sign((got * want).sum(axis=0))
    assert np.allclose(got * sign, want, atol=2e-5)


def test_king_cutoff_feeds_ld_prune(golden_dir, tmp_path):
    """`--king-cutoff 0.02 --indep-pairwise 50 5 0.2`: 49 survivors (below the 50-founder guard, which applies to the
    pre-prune dataset), frequencies frozen before the prune - the reference's chained keep-list."""
    out = _run(golden_dir, tmp_path, ["--bfile", "a", "--king-cutoff", "0.02", "--indep-pairwise", "50", "5", "0.2"])
    assert open(out + ".prune.in", "rb").read() == _gold(golden_dir, "g_acut.prune.in")
    assert open(out + ".king.cutoff.in.id", "rb").read() == _gold(golden_dir, "a_cut.king.cutoff.in.id")


@pytest.mark.parametrize("args,gold", [(["--bfile", "a", "--r2-unphased"], "a_r2.vcor.gz"),
                                       (["--bfile", "a", "--r2-unphased", "--ld-window", "7", "--ld-window-r2", "0.5"], "a_r2w.vcor.gz"),
                                       (["--bfile", "x", "--not-chr", "X", "--keep", "x_keep1.txt", "x_keep2.txt", "--r2-unphased", "--ld-window-r2", "0.3", "--ld-window-kb", "0.1"], "x_r2.vcor.gz")])
def test_r2_unphased_tables_byte_identical(golden_dir, tmp_path, args, gold):
    """--r2-unphased: pair band screened on the device (pl2gpu_ld_band_flags), flagged pairs finished on the host."""
    out = _run(golden_dir, tmp_path, args)
    assert open(out + ".vcor", "rb").read() == _gold(golden_dir, gold)
