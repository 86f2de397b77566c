"""CPU: host-side pieces of plink2_b200 that need no GPU - the pgenlib reader surface (all three
storage modes incl. difflist / LD-compressed records) and byte-identical number formatting."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import plink_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "plink_ng_b200", "plink2_b200")


def _dump(pgen, psam, pvar, tmp_path):
    out = tmp_path / "geno.bin"
    r = subprocess.run([BIN, "--debug-dump-geno", pgen, psam, pvar, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.fromfile(out, dtype=np.uint8)


@pytest.mark.parametrize("pgen,psam,pvar", [("a.bed", "a.fam", "a.bim"), ("a_mode02.pgen", "a.psam", "a.pvar"), ("a_mode10.pgen", "a.psam", "a.pvar")])
def test_reader_decodes_all_storage_modes(golden_dir, tmp_path, pgen, psam, pvar):
    want = orc.read_bed(os.path.join(golden_dir, "a.bed"), 100)
    got = _dump(os.path.join(golden_dir, pgen), os.path.join(golden_dir, psam), os.path.join(golden_dir, pvar), tmp_path)
    assert np.array_equal(got.reshape(want.shape), want)


def test_mode10_fixture_really_uses_compressed_records(golden_dir):
    raw = open(os.path.join(golden_dir, "a_mode10.pgen"), "rb").read()
    assert raw[:3] == b"\x6c\x1b\x10"
    assert len(raw) < 25012  # smaller than the fixed-width file => difflist / LD records present


def test_dtoa_g_matches_reference_text(golden_dir, tmp_path):
    # KINSHIP column of the reference's .kin0 and every entry of its square .king matrix
    geno = orc.read_bed(os.path.join(golden_dir, "a.bed"), 100)
    kin = orc.king_kinship(orc.king_counts(geno))
    extra = np.array([0.0, 1.0, -1.0, 0.5, 123456.7, 1234567.0, 9.9999949e-5, 1e-5, 3.25e-7, 1e300, -2.5e-300, np.inf, -np.inf, np.nan, 0.1, 0.01, 0.001, 0.0001, 99999.95, 999999.5, 0.9999995, 2.0 / 3, 1e15, 1e16])
    vals = np.concatenate([kin, extra])
    fin, fout = tmp_path / "d.bin", tmp_path / "d.txt"
    vals.astype("<f8").tofile(fin)
    r = subprocess.run([BIN, "--debug-dtoa", str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = open(fout).read().split("\n")[:-1]
    ref_rows = [ln.split("\t") for ln in gzip.open(os.path.join(golden_dir, "a_king.kin0.gz"), "rt").read().split("\n")[1:-1]]
    assert [row[-1] for row in ref_rows] == got[: len(kin)]
    # spot checks against C's %g (same 6-significant-digit contract away from exact ties)
    for x, s in zip(extra, got[len(kin):]):
        if np.isnan(x):
            assert s == "nan"
        elif np.isinf(x):
            assert s == ("inf" if x > 0 else "-inf")
        else:
            assert float(s) == pytest.approx(x, rel=6e-6, abs=0), (x, s)
    assert got[len(kin)] == "0" and got[len(kin) + 3] == "0.5" and got[len(kin) + 7] == "1e-05"
