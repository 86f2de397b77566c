"""GPU parity: KING counts through the C-ABI (both kernels) vs the oracle, bit-exact."""
import gzip
import os

import numpy as np
import pytest

import plink_ng_b200 as p
from plink_ng_b200.host import KING_ALGO_POPCOUNT, KING_ALGO_TENSOR, KING_ALGO_TENSOR_TS, KingJob, KingPairJob, pack_genotypes, parallel_bounds
from oracle import plink_oracle as orc

pytestmark = pytest.mark.gpu

ALGOS = [pytest.param(KING_ALGO_POPCOUNT, id="popcount"), pytest.param(KING_ALGO_TENSOR, id="tensor"), pytest.param(KING_ALGO_TENSOR_TS, id="tensor_ts")]


def _random_geno(m, n, seed, miss=0.03):
    rng = np.random.default_rng(seed)
    freq = rng.uniform(0.02, 0.98, size=(m, 1))
    g = (rng.random((m, n)) < freq).astype(np.uint8) + (rng.random((m, n)) < freq).astype(np.uint8)
    g[rng.random((m, n)) < miss] = 3
    return g


def test_umma_operand_layout(gpu_ctx):
    gpu_ctx.selftest_umma(verbose=True)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n,m", [(2, 1), (5, 37), (33, 64), (97, 300), (128, 256), (129, 257), (200, 1000), (385, 513), (700, 2100)])
def test_king_counts_match_oracle(gpu_ctx, algo, n, m):
    geno = _random_geno(m, n, seed=n * 1000 + m)
    with KingJob(gpu_ctx, n, 0, n, algo) as job:
        job.add_variants(pack_genotypes(geno))
        got = job.counts()
        kin = job.kinship()
    want = orc.king_counts(geno)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    wk = orc.king_kinship(want)
    assert np.array_equal(np.isnan(kin), np.isnan(wk))
    ok = ~np.isnan(wk)
    assert np.array_equal(kin[ok], wk[ok])  # same integer numerators/denominators, one IEEE divide


@pytest.mark.parametrize("algo", ALGOS)
def test_king_batches_accumulate_and_row_ranges(gpu_ctx, algo):
    n, m = 300, 1500
    geno = _random_geno(m, n, seed=77)
    geno[10] = 3  # all-missing variant
    geno[:, 7] = 3  # all-missing sample
    gv = pack_genotypes(geno)
    want = orc.king_counts(geno)
    r0, r1 = parallel_bounds(n, 1, 1, 3)
    with KingJob(gpu_ctx, n, r0, r1, algo) as job:
        for s in range(0, m, 400):  # ragged batches
            job.add_variants(gv[s : s + 400])
        got = job.counts()
        sub = job.counts(r0 + 5, r1 - 3)
    tri = lambda r: r * (r - 1) // 2  # noqa: E731
    assert np.array_equal(got, want[tri(r0) : tri(r1)])
    assert np.array_equal(sub, want[tri(r0 + 5) : tri(r1 - 3)])


def test_king_tensor_equals_popcount_medium(gpu_ctx):
    n, m = 1500, 20000
    gv = pack_genotypes(_random_geno(m, n, seed=3, miss=0.01))
    res = []
    for algo in (KING_ALGO_POPCOUNT, KING_ALGO_TENSOR, KING_ALGO_TENSOR_TS):
        with KingJob(gpu_ctx, n, 0, n, algo) as job:
            job.add_variants(gv)
            res.append(job.counts())
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
    # size-independent property: every pair's five categories partition the jointly non-missing variants
    c = res[1].astype(np.int64)
    nsnp = c[:, 1] + c[:, 2] + c[:, 3] + c[:, 4]
    assert nsnp.max() <= m and (c[:, 0] <= c[:, 4]).all()


@pytest.mark.parametrize("n,m,pair_ct", [(2, 5, 1), (97, 300, 50), (385, 1300, 2000), (1000, 70000, 300)])
def test_king_pair_list_matches_oracle(gpu_ctx, n, m, pair_ct):
    """--king-table-subset kernel: listed (first, second) pairs in any order / orientation, "1" = first listed."""
    geno = _random_geno(m, n, seed=n + m)
    rng = np.random.default_rng(pair_ct)
    first = rng.integers(0, n, size=pair_ct)
    second = (first + rng.integers(1, n, size=pair_ct)) % n
    pairs = np.stack([first, second], axis=1).astype(np.uint32)
    with KingPairJob(gpu_ctx, n, pairs) as job:
        gv = pack_genotypes(geno)
        half = max(1, m // 2)
        job.add_variants(gv[:half])  # two batches accumulate
        if m > half:
            job.add_variants(gv[half:])
        got = job.counts()
    want = orc.king_counts_pairs(geno, pairs)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("algo", ALGOS)
def test_king_filter_on_device_matches_host_filter(gpu_ctx, algo):
    """pl2gpu_king_get_filtered == filtering the full table on the host (same fp64 kinship, table order, NaN kept),
    including the buffer-overflow retry and a row sub-range."""
    n, m = 333, 900
    geno = _random_geno(m, n, seed=21)
    geno[:, 5] = 0  # two all-hom-REF samples: no het, no opposite homozygotes -> 0/0 = NaN kinship for the pair (9, 5)
    geno[:, 9] = 0
    with KingJob(gpu_ctx, n, 0, n, algo) as job:
        job.add_variants(pack_genotypes(geno))
        counts, kin = job.counts(), job.kinship()
        ii = np.array([(j, i) for j in range(1, n) for i in range(j)], dtype=np.uint32)
        for thr, r0, r1, cap in ((0.02, 0, n, 1 << 16), (-0.1, 0, n, 7), (0.0, 100, 250, 1 << 16)):
            tri = lambda r: r * (r - 1) // 2  # noqa: E731
            sl = slice(tri(r0), tri(r1))
            keep = ~(kin[sl] < thr)
            p, c, k = job.filtered(thr, cap, r0, r1)
            assert np.array_equal(p, ii[sl][keep]) and np.array_equal(c, counts[sl][keep])
            assert np.array_equal(k, kin[sl][keep], equal_nan=True)
            assert np.isnan(k).any() or thr != -0.1


def test_king_large_block_marginal_identities(gpu_ctx):
    """Size-independent check at a size the pairwise oracle cannot reach (2e8 pairs): summed over all pairs, every
    KING count is a per-variant closed form of the genotype counts n0 (hom-REF), n1 (het), n2 (hom-ALT):
      sum HETHET = sum_v C(n1, 2)            sum IBS0   = sum_v n0 n2
      sum (HET1_HOM2 + HET2_HOM1) = sum_v n1 (n0 + n2)      sum HOMHOM = sum_v C(n0 + n2, 2)
    (a checksum of checksums over the whole N x N result of the default TS tensor kernel, several batches)."""
    n, m = 20000, 8192
    rng = np.random.default_rng(77)
    tot = np.zeros(5, dtype=object)
    want = np.zeros(4, dtype=object)
    with KingJob(gpu_ctx, n) as job:
        for b0 in range(0, m, 2048):
            freq = rng.uniform(0.05, 0.95, size=(2048, 1))
            g = (rng.random((2048, n)) < freq).astype(np.uint8) + (rng.random((2048, n)) < freq).astype(np.uint8)
            g[rng.random((2048, n)) < 0.02] = 3
            job.add_variants(pack_genotypes(g))
            n0, n1, n2 = [(g == c).sum(axis=1).astype(np.int64) for c in (0, 1, 2)]
            hom = n0 + n2
            want += np.array([int((n1 * (n1 - 1) // 2).sum()), int((n0 * n2).sum()), int((n1 * hom).sum()), int((hom * (hom - 1) // 2).sum())], dtype=object)
        for r0 in range(0, n, 2500):
            c = job.counts(r0, min(n, r0 + 2500)).astype(np.int64)  # {IBS0, HETHET, HET2HOM1, HET1HOM2, HOMHOM}
            tot += np.array([int(x) for x in c.sum(axis=0)], dtype=object)
    assert int(tot[1]) == want[0] and int(tot[0]) == want[1] and int(tot[2]) + int(tot[3]) == want[2] and int(tot[4]) == want[3]


@pytest.mark.parametrize("algo", ALGOS)
def test_king_golden_reference_table(gpu_ctx, golden_dir, tmp_path, algo):
    geno = orc.read_bed(os.path.join(golden_dir, "a.bed"), 100)
    with KingJob(gpu_ctx, 100, 0, 100, algo) as job:
        job.add_variants(pack_genotypes(geno))
        counts = job.counts()
        kin = job.kinship()
    f = tmp_path / "k.kin0"
    f.write_bytes(gzip.open(os.path.join(golden_dir, "a_king.kin0.gz")).read())
    _, ints, _ = orc.read_kin0_counts(str(f))
    nsnp, hethet, ibs0, het1hom2, het2hom1, hamming = orc.king_table_columns(counts)
    for name, col in (("NSNP", nsnp), ("HETHET", hethet), ("IBS0", ibs0), ("HET1_HOM2", het1hom2), ("HET2_HOM1", het2hom1), ("IBS", hamming)):
        assert np.array_equal(col, ints[name]), name
    ref = np.fromfile(os.path.join(golden_dir, "a_king.king.bin"), dtype=np.float32)
    assert np.array_equal(kin.astype(np.float32), ref)
