"""GPU parity: GRM through the C-ABI vs the fp64 oracle (which is pinned to the reference's
.grm.bin / .grm.N.bin / .rel.bin in tests/test_oracle_golden.py).

Tolerance (north_star: 1e-5 relative): |G - G_ref| <= 1e-5 * |G_ref| + 1e-10.  The absolute floor
(1e-10 on a matrix whose diagonal is ~1) only matters for off-diagonal entries that cancel below
1e-5; the int8 path's own error is the 2^-(F+1) rounding of the per-variant tables (40 significant
bits relative to the largest table entry), far below the fp32 precision of the reference's files.
Observation counts are exact integers."""
import os

import numpy as np
import pytest

from plink_ng_b200.host import GRM_COV, GRM_MEANIMPUTE, GrmJob, pack_genotypes, parallel_bounds
from oracle import plink_oracle as orc

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-10


def _geno(m, n, seed, miss=0.03, lo=0.02):
    rng = np.random.default_rng(seed)
    freq = rng.uniform(lo, 1 - lo, size=(m, 1))
    g = (rng.random((m, n)) < freq).astype(np.uint8) + (rng.random((m, n)) < freq).astype(np.uint8)
    if miss:
        g[rng.random((m, n)) < miss] = 3
    return g


def _lower(mat, r0, r1):
    return np.concatenate([mat[j, : j + 1] for j in range(r0, r1)])


def _got_lower(rows, r0, r1):
    return np.concatenate([rows[j - r0, : j + 1] for j in range(r0, r1)])


@pytest.mark.parametrize("n,m,flags", [(3, 40, 0), (97, 300, 0), (130, 1000, 0), (385, 700, 0), (200, 600, GRM_MEANIMPUTE), (200, 600, GRM_COV), (150, 500, GRM_COV | GRM_MEANIMPUTE)])
def test_grm_matches_oracle(gpu_ctx, n, m, flags):
    geno = _geno(m, n, seed=n + m + flags)
    geno = geno[(np.stack(orc.genotype_counts(geno), 1)[:, :3] > 0).sum(1) > 1]  # drop monomorphic rows
    want, obs = orc.grm(geno, meanimpute=bool(flags & GRM_MEANIMPUTE), cov=bool(flags & GRM_COV))
    with GrmJob(gpu_ctx, n, 0, n, flags) as job:
        job.add_variants(pack_genotypes(geno))
        got, got_obs = job.rows(with_obs=True)
    a, b = _got_lower(got, 0, n), _lower(want, 0, n)
    assert np.all(np.abs(a - b) <= RTOL * np.abs(b) + ATOL), float(np.max(np.abs(a - b)))
    assert float(np.max(np.abs(a - b))) < 1e-9  # what the 40-bit int8 path actually delivers at these sizes
    if obs is not None:
        assert np.array_equal(_got_lower(got_obs, 0, n), _lower(obs, 0, n).astype(np.float32))


def test_grm_no_missing_uses_variant_count(gpu_ctx):
    n, m = 120, 400
    geno = _geno(m, n, seed=5, miss=0.0, lo=0.1)
    want, obs = orc.grm(geno)
    assert obs is None
    with GrmJob(gpu_ctx, n) as job:
        job.add_variants(pack_genotypes(geno))
        got = job.rows()
    a, b = _got_lower(got, 0, n), _lower(want, 0, n)
    assert np.all(np.abs(a - b) <= RTOL * np.abs(b) + ATOL)


def test_grm_batches_row_ranges_and_given_freqs(gpu_ctx):
    n, m = 300, 1200
    geno = _geno(m, n, seed=21)
    gv = pack_genotypes(geno)
    rf = orc.ref_allele_freqs(geno)
    want, obs = orc.grm(geno, ref_freq=rf)
    r0, r1 = parallel_bounds(n, 0, 1, 3)
    with GrmJob(gpu_ctx, n, r0, r1) as job:
        for s in range(0, m, 500):
            job.add_variants(gv[s : s + 500], ref_freqs=rf[s : s + 500])
        got, got_obs = job.rows(with_obs=True)
    a, b = _got_lower(got, r0, r1), _lower(want, r0, r1)
    assert np.all(np.abs(a - b) <= RTOL * np.abs(b) + ATOL)
    assert np.array_equal(_got_lower(got_obs, r0, r1), _lower(obs, r0, r1).astype(np.float32))


def test_grm_rare_variants_precision(gpu_ctx):
    # singletons/doubletons: the per-variant weights span several orders of magnitude
    n, m = 400, 800
    rng = np.random.default_rng(3)
    geno = _geno(m, n, seed=8, miss=0.01)
    for v in range(0, m, 7):
        geno[v] = 0
        geno[v, rng.integers(0, n, size=rng.integers(1, 3))] = 1
    want, _ = orc.grm(geno)
    with GrmJob(gpu_ctx, n) as job:
        job.add_variants(pack_genotypes(geno))
        got = job.rows()
    a, b = _got_lower(got, 0, n), _lower(want, 0, n)
    # singleton weights are ~800x the common-variant ones, so the shared fixed-point scale costs ~10 bits
    assert np.all(np.abs(a - b) <= RTOL * np.abs(b) + 2e-9), float(np.max(np.abs(a - b)))


def test_grm_degenerate_frequency_is_an_error(gpu_ctx):
    import plink_ng_b200 as p

    geno = _geno(50, 40, seed=2, miss=0.0)
    rf = orc.ref_allele_freqs(geno)
    rf[3] = 1.0  # claims monomorphic REF while genotypes are polymorphic
    with GrmJob(gpu_ctx, 40) as job, pytest.raises(p.Pl2Error):
        job.add_variants(pack_genotypes(geno), ref_freqs=rf)


def test_grm_golden_reference_files(gpu_ctx, golden_dir):
    geno = orc.read_bed(os.path.join(golden_dir, "a.bed"), 100)
    gv = pack_genotypes(geno)
    for fname, flags in (("a_grm.grm.bin", 0), ("a_grmmi.grm.bin", GRM_MEANIMPUTE), ("a_relcov.rel.bin", GRM_COV)):
        with GrmJob(gpu_ctx, 100, 0, 100, flags) as job:
            job.add_variants(gv)
            got, got_obs = job.rows(with_obs=True)
        ref = np.fromfile(os.path.join(golden_dir, fname), dtype=np.float32)
        assert np.allclose(_got_lower(got, 0, 100).astype(np.float32), ref, rtol=2e-6, atol=2e-7), fname
        if flags == 0:
            refn = np.fromfile(os.path.join(golden_dir, "a_grm.grm.N.bin"), dtype=np.float32)
            assert np.array_equal(_got_lower(got_obs, 0, 100), refn)
