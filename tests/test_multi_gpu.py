"""GPU (needs >= 2 devices; skipped on a 1-GPU box): the multi-GPU product path - `plink2_b200 --gpus 2` splits the
N x N triangle into tile-aligned row slabs, every device uploads half of each decoded block and NCCL (inside
libpl2gpu) all-gathers the column tile.  Outputs must be byte-identical to the single-GPU run and to the reference."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "plink_ng_b200", "plink2_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "plink2")
REF_LAPACK = os.path.join(ROOT, "oracle", "_ref", "plink2_lapack")


def _device_count():
    import plink_ng_b200 as p

    return p.lib.pl2gpu_device_count()


def sh(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, " ".join(cmd) + "\n" + r.stdout[-2000:] + r.stderr[-2000:]
    return r


@pytest.fixture(scope="module")
def dummy(tmp_path_factory):
    if _device_count() < 2:
        pytest.skip("needs two CUDA devices")
    d = tmp_path_factory.mktemp("mg")
    pre = str(d / "d")
    sh([REF, "--dummy", "3000", "20000", "0.02", "--seed", "11", "--threads", "4", "--make-bed", "--out", pre])
    return pre


def test_king_two_gpus_byte_identical(dummy, tmp_path):
    flags = ["--make-king", "bin", "triangle", "--make-king-table", "counts", "--king-table-filter", "-0.03"]
    ref, one, two = str(tmp_path / "ref"), str(tmp_path / "one"), str(tmp_path / "two")
    sh([REF, "--bfile", dummy] + flags + ["--threads", "8", "--out", ref])
    sh([BIN, "--bfile", dummy] + flags + ["--out", one])
    r = sh([BIN, "--bfile", dummy] + flags + ["--gpus", "2", "--out", two])
    assert "reduced to 1" not in r.stdout
    for ext in (".king.bin", ".kin0", ".king.id"):
        a = open(ref + ext, "rb").read()
        assert a == open(one + ext, "rb").read(), ext
        assert a == open(two + ext, "rb").read(), ext


def test_grm_two_gpus_matches_single(dummy, tmp_path):
    one, two, ref = str(tmp_path / "one"), str(tmp_path / "two"), str(tmp_path / "ref")
    sh([REF_LAPACK, "--bfile", dummy, "--make-grm-bin", "--threads", "8", "--out", ref])
    sh([BIN, "--bfile", dummy, "--make-grm-bin", "--out", one])
    sh([BIN, "--bfile", dummy, "--make-grm-bin", "--gpus", "2", "--out", two])
    for ext in (".grm.bin", ".grm.N.bin", ".grm.id"):
        assert open(one + ext, "rb").read() == open(two + ext, "rb").read(), ext
    a = np.fromfile(ref + ".grm.bin", dtype=np.float32).astype(np.float64)
    b = np.fromfile(two + ".grm.bin", dtype=np.float32).astype(np.float64)
    assert np.allclose(b, a, rtol=1e-5, atol=2e-8)
    assert open(ref + ".grm.N.bin", "rb").read() == open(two + ".grm.N.bin", "rb").read()
