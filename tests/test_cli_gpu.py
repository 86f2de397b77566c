"""GPU: the plink2 command-line face.  plink2_b200 is run with the reference's own flags on the
golden inputs and its output FILES are compared with the files the reference binary wrote
(tests/golden/, see make_golden.sh): byte-identical for KING tables / matrices / ID lists /
prune lists / observation counts, fp32-identical-or-1ulp for GRM payloads."""
import gzip
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "plink_ng_b200", "plink2_b200")
ENV = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])  # one device: faster CUDA init per process


def run(golden_dir, tmp_path, *flags, inp="bfile"):
    out = str(tmp_path / "o")
    src = ["--bfile", os.path.join(golden_dir, "a")] if inp == "bfile" else ["--pgen", os.path.join(golden_dir, inp), "--pvar", os.path.join(golden_dir, "a.pvar"), "--psam", os.path.join(golden_dir, "a.psam")]
    r = subprocess.run([BIN] + src + list(flags) + ["--out", out], capture_output=True, text=True, env=ENV)
    assert r.returncode == 0, r.stdout + r.stderr
    return out


def gz(golden_dir, name):
    return gzip.open(os.path.join(golden_dir, name), "rb").read()


def test_make_king_table_counts_byte_identical(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--make-king-table", "counts", "cols=+ibs1,+ibs", "--make-king", "bin4", "triangle")
    assert open(out + ".kin0", "rb").read() == gz(golden_dir, "a_king.kin0.gz")
    assert open(out + ".king.bin", "rb").read() == open(os.path.join(golden_dir, "a_king.king.bin"), "rb").read()


def test_make_king_table_proportions_byte_identical(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--make-king-table")
    assert open(out + ".kin0", "rb").read() == gz(golden_dir, "a_kingp.kin0.gz")


def test_make_king_square_text_and_ids(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--make-king", "square")
    assert open(out + ".king", "rb").read() == gz(golden_dir, "a_kingsq.king.gz")
    assert open(out + ".king.id", "rb").read() == open(os.path.join(golden_dir, "a_kingsq.king.id"), "rb").read()


def test_make_king_table_parallel_piece(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--make-king-table", "counts", "--parallel", "2", "3")
    assert open(out + ".kin0.2", "rb").read() == gz(golden_dir, "a_kingpar.kin0.2.gz")


def test_king_table_filter_byte_identical(golden_dir, tmp_path):
    """--king-table-filter runs on the device (pl2gpu_king_get_filtered); same rows, same order, same log count."""
    out = run(golden_dir, tmp_path, "--make-king-table", "counts", "--king-table-filter", "0.02")
    assert open(out + ".kin0", "rb").read() == open(os.path.join(golden_dir, "a_kingfilt.kin0"), "rb").read()
    log = open(out + ".log").read()
    assert "--king-table-filter: 662 relationships reported (4288 filtered out)." in log


def test_king_table_rare_variant_prescan_nsnp_byte_identical(golden_dir, tmp_path):
    """Set Q: the reference's rare-variant pre-scan makes NSNP (and the proportion columns divided by it) one higher for
    (other homozygote, missing) pairs; the host program reproduces that (SparseNsnpFix), counts and proportions."""
    for flags, gold in ((["--make-king-table", "counts", "cols=+ibs1,+ibs"], "q_king.kin0"), (["--make-king-table", "--king-table-filter", "-0.2"], "q_kingp.kin0")):
        out = str(tmp_path / "q")
        r = subprocess.run([BIN, "--bfile", os.path.join(golden_dir, "q")] + flags + ["--out", out], capture_output=True, text=True, env=ENV)
        assert r.returncode == 0, r.stdout + r.stderr
        assert open(out + ".kin0", "rb").read() == open(os.path.join(golden_dir, gold), "rb").read(), gold


def test_king_table_subset_byte_identical(golden_dir, tmp_path):
    """--king-table-subset (pair-list kernel): the reference's own .kin0 as the pair list + kinship threshold, and a
    hand-written IID-only list with swapped orientation and an unknown ID."""
    sub = tmp_path / "in.kin0"
    sub.write_bytes(gz(golden_dir, "a_kingp.kin0.gz"))
    out = run(golden_dir, tmp_path, "--make-king-table", "counts", "--king-table-subset", str(sub), "-0.05")
    assert open(out + ".kin0", "rb").read() == gz(golden_dir, "a_kingsub.kin0.gz")
    out = run(golden_dir, tmp_path, "--make-king-table", "counts", "cols=+ibs1", "--king-table-subset", os.path.join(golden_dir, "a_sub2.txt"))
    assert open(out + ".kin0", "rb").read() == open(os.path.join(golden_dir, "a_kingsub2.kin0"), "rb").read()


def test_king_cutoff_lists(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--king-cutoff", "0.02")
    for ext in (".king.cutoff.in.id", ".king.cutoff.out.id"):
        assert open(out + ext, "rb").read() == open(os.path.join(golden_dir, "a_cut" + ext), "rb").read(), ext


@pytest.mark.parametrize("inp", ["bfile", "a_mode10.pgen", "a_mode02.pgen"])
def test_make_grm_bin_files(golden_dir, tmp_path, inp):
    out = run(golden_dir, tmp_path, "--make-grm-bin", inp=inp)
    got = np.fromfile(out + ".grm.bin", dtype=np.float32)
    ref = np.fromfile(os.path.join(golden_dir, "a_grm.grm.bin"), dtype=np.float32)
    # fp64 values agree to ~1e-10 absolute; after the cast to fp32 that is at most the
    # last bit for entries >= 1e-3 and a 5e-10 absolute difference for the near-zero ones
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= 1.2e-7 * np.abs(ref) + 5e-10)
    assert (got != ref).mean() < 0.05
    assert open(out + ".grm.N.bin", "rb").read() == open(os.path.join(golden_dir, "a_grm.grm.N.bin"), "rb").read()
    if inp == "bfile":
        assert open(out + ".grm.id", "rb").read() == open(os.path.join(golden_dir, "a_grm.grm.id"), "rb").read()


def test_make_grm_list_text(golden_dir, tmp_path):
    """`.grm`: sample indices and observation counts byte-identical, values at dtoa_g's 6 significant digits."""
    out = run(golden_dir, tmp_path, "--make-grm-list")
    got = open(out + ".grm").read().split("\n")
    ref = gz(golden_dir, "a_grml.grm.gz").decode().split("\n")
    assert len(got) == len(ref) == 100 * 101 // 2 + 1
    differing = 0
    for a, b in zip(got, ref):
        if a != b:
            fa, fb = a.split("\t"), b.split("\t")
            assert fa[:3] == fb[:3]
            assert abs(float(fa[3]) - float(fb[3])) <= 1.2e-5 * abs(float(fb[3])) + 5e-10
            differing += 1
    assert differing < 0.01 * len(ref)
    assert open(out + ".grm.id", "rb").read() == open(os.path.join(golden_dir, "a_grm.grm.id"), "rb").read()


def test_make_grm_sparse_text(golden_dir, tmp_path):
    """`.grm.sp`: same (row, column) entries as the reference except values within 1e-9 of the cutoff, values at
    dtoa_g_p8's 8 significant digits (our GRM is within ~1e-10 of the reference's fp64)."""
    out = run(golden_dir, tmp_path, "--make-grm-sparse", "0.02")
    got = {tuple(ln.split("\t")[:2]): ln.split("\t")[2] for ln in open(out + ".grm.sp").read().split("\n") if ln}
    ref = {tuple(ln.split("\t")[:2]): ln.split("\t")[2] for ln in open(os.path.join(golden_dir, "a_grmsp.grm.sp")).read().split("\n") if ln}
    for key in set(got) ^ set(ref):
        v = float(got.get(key, ref.get(key)))
        assert abs(v - 0.02) < 1e-8, key
    differing = 0
    for key in set(got) & set(ref):
        if got[key] != ref[key]:
            assert abs(float(got[key]) - float(ref[key])) <= 2e-9 + 1.5e-8 * abs(float(ref[key]))
            differing += 1
    assert len(ref) == 1380 and differing < 0.05 * len(ref)
    assert open(out + ".grm.id", "rb").read() == open(os.path.join(golden_dir, "a_grm.grm.id"), "rb").read()


def test_make_rel_variants(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--make-rel", "cov", "bin4", "triangle")
    got = np.fromfile(out + ".rel.bin", dtype=np.float32)
    ref = np.fromfile(os.path.join(golden_dir, "a_relcov.rel.bin"), dtype=np.float32)
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= 1.2e-7 * np.abs(ref) + 5e-10)
    out = run(golden_dir, tmp_path, "--make-rel", "square")
    got_txt = open(out + ".rel").read().split("\n")
    ref_txt = gz(golden_dir, "a_rel.rel.gz").decode().split("\n")
    assert len(got_txt) == len(ref_txt)
    diff = 0
    for a, b in zip(got_txt, ref_txt):
        if a != b:
            fa, fb = np.array(a.split("\t"), dtype=float), np.array(b.split("\t"), dtype=float)
            assert np.allclose(fa, fb, rtol=1.2e-5, atol=1e-9)  # one unit in the sixth significant digit
            diff += int((np.array(a.split("\t")) != np.array(b.split("\t"))).sum())
    assert diff <= 100  # of 10,000 six-significant-digit fields: only values within ~1e-10 of a rounding tie differ


@pytest.mark.parametrize("flags,name", [(("50", "5", "0.2"), "a_ld"), (("100", "1", "0.1"), "a_ld2"), (("20kb", "0.3"), "a_ldkb")])
def test_indep_pairwise_lists_byte_identical(golden_dir, tmp_path, flags, name):
    out = run(golden_dir, tmp_path, "--indep-pairwise", *flags)
    assert open(out + ".prune.in", "rb").read() == open(os.path.join(golden_dir, name + ".prune.in"), "rb").read()
    if name == "a_ld":
        assert open(out + ".prune.out", "rb").read() == open(os.path.join(golden_dir, "a_ld.prune.out"), "rb").read()


@pytest.mark.parametrize("inp,name", [("bfile", "a.afreq"), ("a_mode02.pgen", "a_pvar.afreq"), ("a_mode10.pgen", "a_pvar.afreq")])
def test_freq_byte_identical(golden_dir, tmp_path, inp, name):
    """--freq: genotype counts on the device (pl2gpu_geno_counts), .afreq text as the reference writes it."""
    out = run(golden_dir, tmp_path, "--freq", inp=inp)
    assert open(out + ".afreq", "rb").read() == open(os.path.join(golden_dir, name), "rb").read()


def test_freq_sex_chromosomes_byte_identical(golden_dir, tmp_path):
    """--freq on set X: chrX (nonmales twice, a male het half an ALT), chrY (nonfemale founders), MT (haploid counts)."""
    out = str(tmp_path / "xf")
    r = subprocess.run([BIN, "--bfile", os.path.join(golden_dir, "x"), "--freq", "--out", out], capture_output=True, text=True, env=ENV)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(out + ".afreq", "rb").read() == open(os.path.join(golden_dir, "x.afreq"), "rb").read()


def test_indep_preferred_list_byte_identical(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--indep-pairwise", "50", "5", "0.1", "--indep-preferred", os.path.join(golden_dir, "a_pref.txt"))
    assert open(out + ".prune.in", "rb").read() == open(os.path.join(golden_dir, "a_ldpref.prune.in"), "rb").read()


@pytest.mark.parametrize("flags,name", [(("50", "5", "0.2"), "x_o2"), (("50", "5", "0.2", "--indep-order", "1"), "x_o1"), (("30kb", "0.3"), "x_kb")])
def test_indep_pairwise_sex_chromosomes_byte_identical(golden_dir, tmp_path, flags, name):
    """Set X (chr 1 / X / Y / XY / MT, males, females, unknown sex, non-founders): chrX male hets -> missing and
    nonmales at weight 2, chrY nonfemale founders only, MT hets -> missing (2.0/plink2_ld.cc:1356-1389, :982-998)."""
    out = str(tmp_path / "x")
    r = subprocess.run([BIN, "--bfile", os.path.join(golden_dir, "x"), "--indep-pairwise", *flags, "--out", out], capture_output=True, text=True, env=ENV)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(out + ".prune.in", "rb").read() == open(os.path.join(golden_dir, name + ".prune.in"), "rb").read()


def test_indep_order_1_byte_identical(golden_dir, tmp_path):
    out = run(golden_dir, tmp_path, "--indep-pairwise", "50", "5", "0.2", "--indep-order", "1")
    assert open(out + ".prune.in", "rb").read() == open(os.path.join(golden_dir, "a_ldo1.prune.in"), "rb").read()


def _zstd_decompress(path):
    """All frames of a .zst file through the system libzstd (ctypes; the python zstandard module is not in this image)."""
    import ctypes as C

    z = C.CDLL("libzstd.so.1")
    z.ZSTD_findFrameCompressedSize.restype = C.c_size_t
    z.ZSTD_findFrameCompressedSize.argtypes = [C.c_void_p, C.c_size_t]
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    raw = open(path, "rb").read()
    buf = C.create_string_buffer(raw, len(raw))
    base = C.addressof(buf)
    out, off = b"", 0
    while off < len(raw):
        fsz = z.ZSTD_findFrameCompressedSize(base + off, len(raw) - off)
        assert not z.ZSTD_isError(fsz)
        n = z.ZSTD_getFrameContentSize(base + off, fsz)
        dst = C.create_string_buffer(max(int(n), 1))
        got = z.ZSTD_decompress(dst, n, base + off, fsz)
        assert got == n
        out += dst.raw[:n]
        off += fsz
    return out


def test_zs_outputs_decompress_to_the_reference_files(golden_dir, tmp_path):
    """'zs' modifiers: <name>.zst holds exactly the bytes of the uncompressed report (file names per SetKingTableFname,
    2.0/plink2_matrix_calc.cc:1576-1609)."""
    out = run(golden_dir, tmp_path, "--make-king-table", "zs", "--make-king", "square", "zs")
    assert _zstd_decompress(out + ".kin0.zst") == gz(golden_dir, "a_kingp.kin0.gz")
    assert _zstd_decompress(out + ".king.zst") == gz(golden_dir, "a_kingsq.king.gz")
    assert not os.path.exists(out + ".kin0")
    out = run(golden_dir, tmp_path, "--freq", "zs")
    assert _zstd_decompress(out + ".afreq.zst") == open(os.path.join(golden_dir, "a.afreq"), "rb").read()


def test_read_freq_grm_and_prune_list(golden_dir, tmp_path):
    """--read-freq (PLINK 2 --freq report, partial / perturbed / allele-swapped entries): GRM within fp32 rounding of
    the reference's, prune list byte-identical."""
    rf = os.path.join(golden_dir, "a_rf.afreq")
    out = run(golden_dir, tmp_path, "--read-freq", rf, "--make-grm-bin")
    got = np.fromfile(out + ".grm.bin", dtype=np.float32)
    want = np.fromfile(os.path.join(golden_dir, "a_rf.grm.bin"), dtype=np.float32)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7) and not np.allclose(got, np.fromfile(os.path.join(golden_dir, "a_grm.grm.bin"), dtype=np.float32), rtol=1e-5, atol=1e-7)
    out = run(golden_dir, tmp_path, "--read-freq", rf, "--indep-pairwise", "50", "5", "0.2")
    assert open(out + ".prune.in", "rb").read() == open(os.path.join(golden_dir, "a_rfld.prune.in"), "rb").read()


@pytest.mark.parametrize("prefix,flags,golden", [("r", ("rel-check", "counts"), "r_relcheck.kin0"), ("s", ("rel-check",), "s_relcheck.kin0.gz")])
def test_rel_check_table_byte_identical(golden_dir, tmp_path, prefix, flags, golden):
    """--make-king-table rel-check: same-FID pairs in PLINK 2's natural sort order of the IDs (leading zeros, mixed case,
    digit runs; 300 random IDs over 5 FIDs in set S), counts through the pair-list kernel."""
    out = str(tmp_path / "rc")
    r = subprocess.run([BIN, "--bfile", os.path.join(golden_dir, prefix), "--make-king-table", *flags, "--out", out], capture_output=True, text=True, env=ENV)
    assert r.returncode == 0, r.stdout + r.stderr
    want = gz(golden_dir, golden) if golden.endswith(".gz") else open(os.path.join(golden_dir, golden), "rb").read()
    assert open(out + ".kin0", "rb").read() == want


def test_toy_fixture_configs0(golden_dir, tmp_path):
    out = str(tmp_path / "toy")
    r = subprocess.run([BIN, "--bfile", os.path.join(golden_dir, "toy"), "--make-king", "square", "--make-king-table", "counts", "cols=+ibs1,+ibs", "--out", out], capture_output=True, text=True, env=ENV)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(out + ".king", "rb").read() == open(os.path.join(golden_dir, "toy_king.king"), "rb").read()
    assert open(out + ".kin0", "rb").read() == open(os.path.join(golden_dir, "toy_king.kin0"), "rb").read()


def test_unsupported_flag_fails_loudly(golden_dir, tmp_path):
    r = subprocess.run([BIN, "--bfile", os.path.join(golden_dir, "a"), "--glm", "--out", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode == 8 and "unsupported" in r.stdout
