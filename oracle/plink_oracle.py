"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the reference's pairwise-genotype math.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (plink_ng_b200/, libpl2gpu.so, plink2_b200) never does.

Every function cites the reference (chrchang/plink-ng @ 22df0611, paths relative to
/root/reference) it restates.  Parity is PINNED: tests/test_oracle_golden.py checks these
restatements against outputs of the reference binary itself (oracle/_ref/plink2, built by
oracle/build_ref.sh) committed under tests/golden/ together with the generating script.

Genotype matrices are uint8 [variants, samples] with the PgrGet codes 0 = hom-REF, 1 = het,
2 = hom-ALT, 3 = missing (2.0/include/pgenlib_read.h:537; pgen_spec/pgen_spec.tex:439-441).
"""
import numpy as np

SMALL_EPSILON = 2.0 ** -44  # kSmallEpsilon, 2.0/include/plink2_base.h


# --------------------------------------------------------------------------------------------- I/O
def read_bed(path: str, sample_ct: int) -> np.ndarray:
    """PLINK 1 variant-major .bed -> codes.  On disk 0 = hom-A1(ALT), 1 = missing, 2 = het,
    3 = hom-A2(REF) (pgen_spec.tex:436-438); remapped like 2.0/include/pgenlib_read.cc:2890-2891."""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[0] == 0x6C and raw[1] == 0x1B and raw[2] == 0x01, "not a variant-major .bed"
    bpv = (sample_ct + 3) // 4
    body = raw[3:].reshape(-1, bpv)
    codes = np.stack([(body >> s) & 3 for s in (0, 2, 4, 6)], axis=-1).reshape(body.shape[0], -1)[:, :sample_ct]
    remap = np.array([2, 3, 1, 0], dtype=np.uint8)
    return remap[codes]


def read_fixed_pgen(path: str) -> np.ndarray:
    """Mode 0x02 fixed-width .pgen (pgen_spec.tex:139-141): 12-byte header then ceil(N/4)-byte
    records already in PgrGet coding."""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[0] == 0x6C and raw[1] == 0x1B and raw[2] == 0x02, "not a mode-0x02 .pgen"
    m = int(raw[3:7].view("<u4")[0])
    n = int(raw[7:11].view("<u4")[0])
    bpv = (n + 3) // 4
    body = raw[12 : 12 + m * bpv].reshape(m, bpv)
    codes = np.stack([(body >> s) & 3 for s in (0, 2, 4, 6)], axis=-1).reshape(m, -1)[:, :n]
    return np.ascontiguousarray(codes)


# -------------------------------------------------------------------------------------------- KING
def split_hom_ref2het(geno: np.ndarray):
    """SplitHomRef2hetUnsafeW (2.0/include/pgenlib_misc.cc:1797-1866): hom = geno in {0,2},
    ref2het = geno in {0,1}, missing -> (0,0)."""
    hom = (geno == 0) | (geno == 2)
    ref2het = (geno == 0) | (geno == 1)
    return hom, ref2het


def king_count_matrices(geno: np.ndarray):
    """IncrKingHomhom's five per-pair sums (2.0/plink2_matrix_calc.cc:1309-1322) for ALL ordered
    pairs, as dense [N,N] int64 matrices indexed [second(larger idx), first(smaller idx)]:
      ibs0     = sum pc((r1^r2) & h1 & h2)     hethet   = sum pc(het1 & het2)
      het2hom1 = sum pc(hom1 & het2)           het1hom2 = sum pc(hom2 & het1)    homhom = sum pc(h1&h2)
    with het = ref2het & ~hom.  Exact: 0/1 indicator products summed in float64 (< 2^53)."""
    hom, r2h = split_hom_ref2het(geno)
    het = r2h & ~hom
    homref = hom & r2h
    homalt = hom & ~r2h
    f = lambda x: np.ascontiguousarray(x.T, dtype=np.float32 if geno.shape[0] < (1 << 24) else np.float64)  # noqa: E731
    H, T, A, B = f(hom), f(het), f(homref), f(homalt)
    to_i = lambda x: np.rint(x).astype(np.int64)  # noqa: E731
    ab = A @ B.T
    ibs0 = to_i(ab + ab.T)            # [2,1]: (r1^r2)&h1&h2 is symmetric
    hethet = to_i(T @ T.T)
    t_h = to_i(T @ H.T)               # [x,y] = het_x . hom_y
    het2hom1 = t_h                    # [second, first] = het_second . hom_first
    het1hom2 = t_h.T.copy()           # [second, first] = hom_second . het_first
    homhom = to_i(H @ H.T)
    return ibs0, hethet, het2hom1, het1hom2, homhom


def king_counts(geno: np.ndarray, row_start: int = 0, row_end: int = None) -> np.ndarray:
    """uint32 king_counts[pair][5] = {IBS0, HETHET, HET2HOM1, HET1HOM2, HOMHOM}
    (2.0/plink2_matrix_calc.cc:864-868) in CalcKingDenseThread's pair order: for second in
    [row_start,row_end): for first in [0, second) (:1545-1547)."""
    n = geno.shape[1]
    row_end = n if row_end is None else row_end
    mats = king_count_matrices(geno)
    rows = []
    for j in range(row_start, row_end):
        if j:
            rows.append(np.stack([m[j, :j] for m in mats], axis=1))
    if not rows:
        return np.zeros((0, 5), dtype=np.uint32)
    return np.concatenate(rows, axis=0).astype(np.uint32)


def king_counts_pairs(geno: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """IncrKingSubsetHomhom (2.0/plink2_matrix_calc.cc:2533-2575): the same five sums for an explicit list
    of (first, second) sample pairs, "1" = first listed.  uint32 [pair][5] = {IBS0, HETHET, HET2HOM1,
    HET1HOM2, HOMHOM}."""
    hom, r2h = split_hom_ref2het(geno)
    het = r2h & ~hom
    out = np.zeros((len(pairs), 5), dtype=np.uint32)
    for p, (a, b) in enumerate(np.asarray(pairs, dtype=np.int64)):
        hom1, hom2, het1, het2 = hom[:, a], hom[:, b], het[:, a], het[:, b]
        hh = hom1 & hom2
        out[p] = (np.count_nonzero((r2h[:, a] ^ r2h[:, b]) & hh), np.count_nonzero(het1 & het2), np.count_nonzero(hom1 & het2), np.count_nonzero(hom2 & het1), np.count_nonzero(hh))
    return out


def king_counts_bruteforce(geno: np.ndarray) -> np.ndarray:
    """Pure-Python loop over pairs and variants (tiny inputs only): the literal per-genotype table
    behind IncrKingHomhom, used to cross-check the vectorised restatement."""
    m, n = geno.shape
    out = []
    for j in range(n):
        for i in range(j):
            ibs0 = hethet = het2hom1 = het1hom2 = homhom = 0
            for v in range(m):
                g1, g2 = int(geno[v, i]), int(geno[v, j])
                if g1 == 3 or g2 == 3:
                    continue
                hom1, hom2 = g1 != 1, g2 != 1
                if hom1 and hom2:
                    homhom += 1
                    if g1 != g2:
                        ibs0 += 1
                elif (not hom1) and (not hom2):
                    hethet += 1
                elif hom1:
                    het2hom1 += 1
                else:
                    het1hom2 += 1
            out.append((ibs0, hethet, het2hom1, het1hom2, homhom))
    return np.array(out, dtype=np.uint32).reshape(-1, 5)


def king_kinship(counts: np.ndarray) -> np.ndarray:
    """ComputeKinship (2.0/plink2_matrix_calc.cc:1566-1573) with zero singleton terms."""
    c = counts.astype(np.int64)
    ibs0, hethet, het2hom1, het1hom2 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    smaller_het = hethet + np.minimum(het1hom2, het2hom1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return 0.5 - (4 * ibs0 + het1hom2 + het2hom1).astype(np.float64) / (4 * smaller_het).astype(np.float64)


def king_table_columns(counts: np.ndarray):
    """Integer columns of `--make-king-table counts cols=+ibs1,+ibs` (:2292-2356):
    NSNP, HETHET, IBS0, HET1_HOM2, HET2_HOM1, IBS(hamming)."""
    c = counts.astype(np.int64)
    ibs0, hethet, het2hom1, het1hom2, homhom = (c[:, k] for k in range(5))
    nsnp = het1hom2 + het2hom1 + homhom + hethet
    hamming = 2 * ibs0 + het1hom2 + het2hom1
    return nsnp, hethet, ibs0, het1hom2, het2hom1, hamming


def king_sparse_nsnp_extra(geno: np.ndarray, row_end: int = None) -> np.ndarray:
    """What the reference's rare-variant pre-scan adds to NSNP beyond the dense count, per pair in king_counts
    order.  CalcKingSparseThread (2.0/plink2_matrix_calc.cc:904-1250) pre-scans a variant when hom-REF (else
    hom-ALT) covers all but row_end // 33 of the pass's samples (KingMaxSparseCt :1654, AVX2 build; test order
    :985-1001).  Its pair corrections reproduce dense counting except for (other homozygote, missing) pairs,
    which get HOMHOM + 1 (:1086-1096, :1129-1139): the table's NSNP = HET1_HOM2 + HET2_HOM1 + HOMHOM + HETHET
    (:2315-2318) is then one higher per such variant.  Verified against the reference binary on the
    4,096 x 65,536 --dummy set (179 of 254,515 table rows differ from the dense count, all explained by this)."""
    m, n = geno.shape
    s_ct = n if row_end is None else row_end
    g = geno[:, :s_ct]
    n0, n2 = (g == 0).sum(axis=1), (g == 2).sum(axis=1)
    min_common = s_ct - s_ct // 33
    c0 = n0 >= min_common
    c2 = (~c0) & (n2 >= min_common)
    sp = np.flatnonzero(c0 | c2)
    gs = g[sp]
    other = np.where(c0[sp][:, None], gs == 2, gs == 0).astype(np.int64)
    miss = (gs == 3).astype(np.int64)
    om = other.T @ miss
    om = om + om.T  # [a, b]: pre-scanned variants where one of a, b is the other homozygote and its partner is missing
    rows = [om[j, :j] for j in range(1, s_ct)]
    return np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)


def read_kin0_counts(path: str):
    """Parse a reference `.kin0` written with `counts cols=+ibs1,+ibs` into (ids, int columns, kinship)."""
    with open(path) as f:
        header = f.readline().rstrip("\n").lstrip("#").split("\t")
        rows = [ln.rstrip("\n").split("\t") for ln in f]
    col = {h: k for k, h in enumerate(header)}
    ints = {h: np.array([int(r[col[h]]) for r in rows], dtype=np.int64) for h in ("NSNP", "HETHET", "IBS0", "HET1_HOM2", "HET2_HOM1", "IBS") if h in col}
    kin = np.array([float(r[col["KINSHIP"]]) for r in rows]) if "KINSHIP" in col else None
    ids = [(r[col["IID1"]], r[col["IID2"]]) for r in rows]
    return ids, ints, kin


# -------------------------------------------------------------------------------- allele frequencies
def genotype_counts(geno: np.ndarray):
    """Per-variant (hom-REF, het, hom-ALT, missing) counts - GenoarrCountFreqsUnsafe
    (2.0/include/pgenlib_misc.cc:702)."""
    return tuple((geno == c).sum(axis=1).astype(np.int64) for c in range(4))


def ref_allele_freqs(geno: np.ndarray) -> np.ndarray:
    """ComputeAlleleFreqs (2.0/plink2_filter.cc:2113-2151), biallelic hard calls, no pseudocount:
    ref_freq = ref_count * (1.0 / total_count) (multiply by reciprocal, :2144-2148); 0.5 when the
    variant has no non-missing founder call (:2138-2142)."""
    n0, n1, n2, _ = genotype_counts(geno)
    ref_ct = (2 * n0 + n1).astype(np.float64)
    tot = (2 * (n0 + n1 + n2)).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = ref_ct * (1.0 / tot)
    return np.where(tot == 0, 0.5, f)


def major_allele_freqs(ref_freq: np.ndarray) -> np.ndarray:
    """GetMajIdx / GetAlleleFreq (2.0/plink2_common.h:559-595): REF is major iff ref_freq >= 0.5;
    the ALT frequency is 1.0 - ref_freq."""
    return np.where(ref_freq >= 0.5, ref_freq, np.maximum(1.0 - ref_freq, 0.0))


def read_freq_overrides(path: str, ids, ref_alleles, alt_alleles) -> np.ndarray:
    """`--read-freq` on a PLINK 2 --freq report (ReadAlleleFreqs, 2.0/plink2_filter.cc:2242-3300, the
    kfReadFreqColsetAltFreqs branch for biallelic variants, no pseudocount): per dataset variant the loaded REF
    frequency, NaN where the file has no usable entry (unknown ID, allele codes that do not match, nan; OBS_CT is only
    consulted for count columns, :3170-3175, so a frequency line with OBS_CT 0 is still loaded).
    REF/ALT listed the other way round: the listed ALT frequency is the dataset's REF frequency (:3187-3192)."""
    idx = {v: k for k, v in enumerate(ids)}
    out = np.full(len(ids), np.nan)
    with open(path) as f:
        hdr = f.readline().rstrip("\n").lstrip("#").split("\t")
        col = {name: i for i, name in enumerate(hdr)}
        for ln in f:
            t = ln.rstrip("\n").split("\t")
            k = idx.get(t[col["ID"]])
            if k is None:
                continue
            try:
                af = float(t[col["ALT_FREQS"]])
            except ValueError:
                continue
            if af != af:
                continue
            fr, fa = t[col["REF"]], t[col["ALT"]]
            if fr == ref_alleles[k] and fa == alt_alleles[k]:
                out[k] = 1.0 - af
            elif fr == alt_alleles[k] and fa == ref_alleles[k]:
                out[k] = af
    return out


# ------------------------------------------------------------------------------------------- --score
def score_file_entries(path: str, ids, ref_alleles, alt_alleles, header: bool = True, id_col: int = 0, allele_col: int = 1, coef_col: int = 2):
    """The (variant index, named-allele index 0 = REF / 1 = ALT, coefficient) triples ScoreReport keeps, in file order
    (2.0/plink2_matrix_calc.cc:7712-7790): unknown variant IDs and allele codes that are neither REF nor ALT are
    skipped (and counted for the warning)."""
    idx = {v: k for k, v in enumerate(ids)}
    out, missing_id, missing_allele = [], 0, 0
    with open(path) as f:
        if header:
            f.readline()
        for ln in f:
            t = ln.split()
            if not t:
                continue
            k = idx.get(t[id_col])
            if k is None:
                missing_id += 1
                continue
            a = t[allele_col]
            if a == ref_alleles[k]:
                out.append((k, 0, float(t[coef_col])))
            elif a == alt_alleles[k]:
                out.append((k, 1, float(t[coef_col])))
            else:
                missing_allele += 1
    return out, missing_id, missing_allele


def score_report(geno: np.ndarray, entries, ref_freq: np.ndarray, no_mean_imputation: bool = False, mode: str = ""):
    """ScoreReport's default report for diploid variants (2.0/plink2_matrix_calc.cc:6892-9270): per sample
    ALLELE_CT = 2 x nonmissing scored variants (:8581), DENOM (= 2 x scored variants, or ALLELE_CT with
    'no-mean-imputation', :8586-8588), NAMED_ALLELE_DOSAGE_SUM over nonmissing calls, score sum = sum of coefficient x
    named-allele dosage with a missing call replaced by 2 x the named allele's frequency (:6605-6607) unless
    no-mean-imputation, and the average = sum x (1 / DENOM) (:8397).  mode "center": dosage - 2 f; "variance-standardize":
    (dosage - 2 f) / sqrt(2 f (1 - f)) (slope 0 when the variance is not above 2^-44, :8005-8033).  NOTE: under both
    modes the reference still adds the UNCENTRED mean 2 f x slope for a missing call (missing_effect has no intercept,
    :6756-6762) - reproduced here because the goldens say so."""
    n = geno.shape[1]
    ssum = np.zeros(n)
    dos = np.zeros(n, dtype=np.int64)
    miss = np.zeros(n, dtype=np.int64)
    for v, aidx, coef in entries:
        g = geno[v]
        named = np.where(g == 3, 0, g if aidx == 1 else 2 - g).astype(np.int64)
        named = np.where(g == 3, 0, named)
        f_named = (1.0 - ref_freq[v]) if aidx == 1 else ref_freq[v]
        slope, icpt = 1.0, 0.0
        if mode in ("center", "variance-standardize"):
            if mode == "variance-standardize":
                var = 2.0 * f_named * (1.0 - f_named)
                slope = 1.0 / np.sqrt(var) if var > SMALL_EPSILON else 0.0
            icpt = (-2.0 * f_named) * slope
        miss_dosage = 2.0 * f_named
        if mode in ("dominant", "recessive"):  # copies -> min(copies, 1) / max(copies - 1, 0); a missing call -> ONE x f (:6747-6762)
            named = np.minimum(named, 1) if mode == "dominant" else np.maximum(named - 1, 0)
            miss_dosage = f_named
        d = named.astype(np.float64) * slope + icpt
        d = np.where(g == 3, 0.0 if no_mean_imputation else miss_dosage * slope, d)
        ssum += coef * d
        dos += named
        miss += g == 3
    denom_full = 2 * len(entries)
    nallele = denom_full - 2 * miss
    denom = nallele if no_mean_imputation else np.full(n, denom_full)
    return nallele, denom, dos, ssum, ssum * (1.0 / denom)


def variant_scores(geno: np.ndarray, weights: np.ndarray, ref_freq: np.ndarray) -> np.ndarray:
    """VscoreReport (2.0/plink2_matrix_calc.cc:9274): per variant the dot product of the sample weights
    [samples, cols] with the ALT dosages, a missing call replaced by 2 x the ALT frequency (the dataset's allele
    frequencies, not the scored subset's).  Samples outside the score file carry weight 0."""
    d = np.where(geno == 3, (2.0 * (1.0 - ref_freq))[:, None], geno.astype(np.float64))
    return d @ weights


# --------------------------------------------------------------------------------------------- GRM
def centered_varmaj(geno: np.ndarray, ref_freq: np.ndarray, variance_standardize: bool = True) -> np.ndarray:
    """ExpandCenteredVarmaj + PopulateRescaledDosage (2.0/plink2_matrix_calc.cc:3839-3886,
    2.0/plink2_common.cc:323-341): per variant a 4-entry table {intercept, intercept+slope,
    intercept+2*slope, 0.0(missing)} with slope = inv_stdev, intercept = -2*alt_freq*inv_stdev.
    Zero-variance variants contribute all zeros (the consistency checks at :3844-3868 only decide
    whether the reference errors out)."""
    alt = 1.0 - ref_freq
    if variance_standardize:
        variance = 2 * ref_freq * alt
        ok = variance > SMALL_EPSILON
        with np.errstate(divide="ignore", invalid="ignore"):
            inv_stdev = np.where(ok, 1.0 / np.sqrt(variance), 0.0)
    else:
        ok = np.ones_like(ref_freq, dtype=bool)
        inv_stdev = np.ones_like(ref_freq)
    slope = inv_stdev
    intercept = -2 * alt * inv_stdev
    table = np.stack([intercept, intercept + slope, intercept + 2 * slope, np.zeros_like(slope)], axis=1)
    table[~ok] = 0.0
    return np.take_along_axis(table, geno.astype(np.int64), axis=1)  # [variants, samples]


def grm(geno: np.ndarray, ref_freq: np.ndarray = None, meanimpute: bool = False, cov: bool = False):
    """CalcGrm (2.0/plink2_matrix_calc.cc:4555-4788): grm = Z^T Z in fp64, then each entry divided by
    its own observation count (M - miss_i - miss_j + bothmiss_ij; diagonal M - miss_i), or multiplied
    by 1/M with `meanimpute` or when no variant has a missing call (:4756-4788).
    Returns (G [N,N] float64 full symmetric, obs_counts [N,N] int64 or None)."""
    if ref_freq is None:
        ref_freq = ref_allele_freqs(geno)
    z = centered_varmaj(geno, ref_freq, not cov)
    g = z.T @ z
    m = geno.shape[0]
    miss = (geno == 3)
    if meanimpute or not miss.any():
        return g * (1.0 / float(m)), None
    # CalcMissingMatrix (:4404-4553): only variants with >= 1 missing call are scanned; totals identical
    mf = miss.astype(np.float64)
    missing_cts = miss.sum(axis=0).astype(np.int64)
    both = np.rint(mf.T @ mf).astype(np.int64)
    obs = m - missing_cts[:, None] - missing_cts[None, :] + both
    np.fill_diagonal(obs, m - missing_cts)
    with np.errstate(divide="ignore", invalid="ignore"):
        return g / obs.astype(np.float64), obs


def lower_triangle_with_diag(mat: np.ndarray) -> np.ndarray:
    """Row-major lower triangle including the diagonal (.grm.bin order, :4963-4971)."""
    n = mat.shape[0]
    return np.concatenate([mat[j, : j + 1] for j in range(n)])


# -------------------------------------------------------------------------------- --indep-pairwise
def ld_pair_components(x: np.ndarray, nm: np.ndarray, a: int, bs: np.ndarray, w: np.ndarray = None):
    """ComputeIndepPairwiseR2Components (2.0/plink2_ld.cc:699-723) for `second` = a against every
    `first` in bs: all six integers restricted to samples non-missing in both variants.
    x in {+1 (code 0), 0 (het), -1 (code 2), 0 (missing)}; nm = non-missing indicator.
    w: per-sample integer weights - chrX adds the nonmale-only sums twice to the all-founder sums
    (:982-998, :1064-1078, "--ld-xchr 3"), i.e. weight 1 for males and 2 for everyone else."""
    xa, na = x[a], nm[a]
    xa2 = xa * xa
    if w is not None:
        xa, na, xa2 = xa * w, na * w, xa2 * w
    xb, nb = x[bs], nm[bs]
    nm_ct = nb @ na
    dot = xb @ xa
    s_b = xb @ na
    q_b = (xb * xb) @ na
    s_a = nb @ xa
    q_a = nb @ xa2
    r = lambda v: np.rint(v).astype(np.int64)  # noqa: E731
    return r(nm_ct), r(s_b), r(q_b), r(s_a), r(q_a), r(dot)


def ld_prune_subcontig(geno: np.ndarray, maj_freq: np.ndarray, bps, window: int, step: int, r2_thresh: float, w: np.ndarray = None, order: int = 2) -> np.ndarray:
    """IndepPairwiseThread (2.0/plink2_ld.cc:862-1109) for one subcontig: default branch (:1039-1100) or, with
    order = 1, the PLINK 1.x pruning order of `--indep-order 1` (:931-1037).  `geno` [L, founders]; returns removed[L] bool.  Window bookkeeping follows
    LdPruneNextSubcontig (:605-633) and LdPruneNextWindow (:635-689); `bps` is None for
    variant-count windows.  The major-allele inversion of PgrGetInv1 (:1357) does not change any
    decision (cov^2 and both variances are invariant under negating a variable), so plain codes are
    used."""
    L = geno.shape[0]
    thr = r2_thresh * (1 + SMALL_EPSILON)  # :1255
    x = np.where(geno == 0, 1.0, np.where(geno == 2, -1.0, 0.0)).astype(np.float32)
    nm = (geno != 3).astype(np.float32)
    wv = np.ones(geno.shape[1], dtype=np.int64) if w is None else np.asarray(w, dtype=np.int64)
    nm_ct_v = (geno != 3).astype(np.int64) @ wv
    plus_v = (geno == 0).astype(np.int64) @ wv
    minus_v = (geno == 2).astype(np.int64) @ wv
    wf = None if w is None else wv.astype(np.float32)
    mono = ((plus_v == 0) & (minus_v == 0)) | (plus_v == nm_ct_v) | (minus_v == nm_ct_v)  # :902
    removed = np.zeros(L, dtype=bool)
    if L < 2:
        return removed
    start = 0
    if bps is not None:
        bp_thresh = int(bps[0]) + window
        first_len = 1
        idx = 0
        while True:
            idx += 1
            if not (bps[idx] <= bp_thresh):
                break
            first_len += 1
            if not (first_len < L):
                break
        next_end = first_len
    else:
        next_end = min(L, window)
    win = []          # tvidx per window position
    win_removed = []  # cur_window_removed bit per window position
    winpos_split = 0
    first_unchecked = {}  # --indep-order 1: first_unchecked_tvidx per live variant (:919-921)

    def over_threshold(a, firsts):
        nm_ct, s_b, q_b, s_a, q_a, dot = ld_pair_components(x, nm, a, firsts, wf)
        cov12 = (dot * nm_ct - s_b * s_a).astype(np.float64)
        var1 = (q_b * nm_ct - s_b * s_b).astype(np.float64)  # first
        var2 = (q_a * nm_ct - s_a * s_a).astype(np.float64)  # second
        return cov12 * cov12 > thr * var1 * var2

    for cur in range(L):
        win.append(cur)
        if mono[cur]:
            win_removed.append(True)
            removed[cur] = True
        else:
            win_removed.append(False)
            first_unchecked[cur] = cur + 1
        if cur + 1 != next_end:
            continue
        cur_tvidx = cur + 1
        if order == 1:
            # PLINK 1.x order (:931-1037): sweep firsts in ascending order, each against the not-yet-checked
            # seconds after it; repeat the sweep while it removes something
            removed_ct = sum(win_removed)
            while True:
                prev_ct = removed_ct
                first_winpos = -1
                while True:
                    first_winpos += 1
                    while first_winpos < len(win) and win_removed[first_winpos]:
                        first_winpos += 1
                    if first_winpos >= len(win):
                        break
                    b = win[first_winpos]
                    fu = first_unchecked[b]
                    if fu == cur_tvidx:
                        continue
                    live = [p for p in range(first_winpos + 1, len(win)) if not win_removed[p]]  # snapshot, like BitIter0
                    live = [p for p in live if win[p] >= fu]
                    if not live:
                        first_unchecked[b] = cur_tvidx
                        continue
                    seconds = np.array([win[p] for p in live], dtype=np.int64)
                    # pair (first b, second a): same sextuple, var1 = first's
                    over = np.array([over_threshold(int(a2), np.array([b]))[0] for a2 in seconds])
                    hit = np.flatnonzero(over)
                    if hit.size == 0:
                        first_unchecked[b] = cur_tvidx
                        continue
                    h = int(hit[0])
                    a = int(seconds[h])
                    if maj_freq[b] > maj_freq[a] * (1 + SMALL_EPSILON):
                        win_removed[first_winpos] = True
                        removed[b] = True
                    else:
                        win_removed[live[h]] = True
                        removed[a] = True
                        first_unchecked[b] = win[live[h + 1]] if h + 1 < len(live) else cur_tvidx
                removed_ct = sum(win_removed)
                if not removed_ct > prev_ct:
                    break
        second_stop = winpos_split if winpos_split else 1
        for second_winpos in (range(len(win) - 1, second_stop - 1, -1) if order != 1 else ()):
            a = win[second_winpos]
            firsts = np.array(win[:second_winpos], dtype=np.int64)
            if firsts.size == 0:
                continue
            nm_ct, s_b, q_b, s_a, q_a, dot = ld_pair_components(x, nm, a, firsts, wf)
            cov12 = (dot * nm_ct - s_b * s_a).astype(np.float64)
            var1 = (q_b * nm_ct - s_b * s_b).astype(np.float64)  # first
            var2 = (q_a * nm_ct - s_a * s_a).astype(np.float64)  # second
            over = cov12 * cov12 > thr * var1 * var2
            for first_winpos in range(second_winpos - 1, -1, -1):
                if win_removed[first_winpos]:
                    continue
                if over[first_winpos]:
                    b = win[first_winpos]
                    if maj_freq[b] <= maj_freq[a] * (1 + SMALL_EPSILON):
                        win_removed[second_winpos] = True
                        removed[a] = True
                        break
                    win_removed[first_winpos] = True
                    removed[b] = True
        # LdPruneNextWindow
        if next_end == L:
            break
        if bps is not None:
            min_bp = int(bps[next_end]) - window
            nstart = start
            while True:
                nstart += 1
                sbp = int(bps[nstart])
                if not (sbp < min_bp):
                    break
            end_thresh = sbp + window
            e = next_end
            while True:
                e += 1
                if e == L:
                    break
                if not (bps[e] <= end_thresh):
                    break
            start, next_end = nstart, e
        else:
            start += step
            next_end = min(start + window, L)
        keep = [k for k in range(len(win)) if (not win_removed[k]) and win[k] >= start]
        win = [win[k] for k in keep]
        win_removed = [False] * len(win)
        winpos_split = len(win)
    return removed


def _chr_class(c) -> str:
    """'x' / 'y' / 'hap' (MT) / 'dip' from a chromosome code or name (human: haploid_mask = X, Y, MT,
    2.0/plink2_common.cc:1979)."""
    t = str(c).upper()
    if t.startswith("CHR"):
        t = t[3:]
    return {"X": "x", "23": "x", "Y": "y", "24": "y", "MT": "hap", "M": "hap", "26": "hap"}.get(t, "dip")


def ld_ref_freqs_by_class(geno: np.ndarray, chrom: np.ndarray, sex: np.ndarray = None) -> np.ndarray:
    """REF allele frequencies as LoadAlleleAndGenoCountsThread + ComputeAlleleFreqs produce them for founders
    (2.0/plink2_data.cc:2420-2690, 2.0/plink2_filter.cc:2113-2151): autosomes and MT = hard-call ratio; chrY =
    the same ratio over nonfemale founders (:2458-2483); chrX = nonmales count twice, males once, a male het is
    half an ALT (:2642, :2685-2688).  sex: 1 male, 2 female, 0 unknown."""
    m, n = geno.shape
    sex = np.zeros(n, dtype=np.int64) if sex is None else np.asarray(sex)
    male, nonfemale = sex == 1, sex != 2
    out = ref_allele_freqs(geno)
    cls = np.array([_chr_class(c) for c in chrom])
    ysel = cls == "y"
    if ysel.any():
        out[ysel] = ref_allele_freqs(geno[ysel][:, nonfemale]) if nonfemale.any() else 0.5
    xsel = cls == "x"
    if xsel.any():
        gx = geno[xsel]
        n0, n1, n2, n3 = genotype_counts(gx)
        m0, m1, m2, m3 = genotype_counts(gx[:, male])
        alt1 = 4 * n2 + 2 * n1 - 2 * m2 - m1
        wobs = (2 * (n - n3) - int(male.sum()) + m3) * 2
        with np.errstate(divide="ignore", invalid="ignore"):
            f = (wobs - alt1).astype(np.float64) * (1.0 / wobs.astype(np.float64))
        out[xsel] = np.where(wobs == 0, 0.5, f)
    return out


def ld_class_block(g: np.ndarray, cls: str, sex: np.ndarray):
    """The genotype block IndepPairwise's loader builds for one chromosome class (2.0/plink2_ld.cc:1356-1389) and the
    per-sample weights of the pair sums (:982-998): MT = hets -> missing; chrY = nonfemale founders, hets -> missing;
    chrX = male hets -> missing, males weight 1, everyone else weight 2."""
    male, nonfemale = sex == 1, sex != 2
    if cls == "hap":
        return np.where(g == 1, 3, g), None
    if cls == "y":
        return np.where(g == 1, 3, g)[:, nonfemale], None
    if cls == "x":
        return np.where((g == 1) & male[None, :], 3, g), np.where(male, 1, 2)
    return g, None


def ld_walk_inputs(geno: np.ndarray, chrom: np.ndarray, r2_thresh: float, band: int, sex: np.ndarray = None, preferred: np.ndarray = None):
    """Everything the greedy walk looks at, per variant in file order: major-allele frequency (minus 1 for preferred
    variants), the load-time monomorphic mark (:902) and the pair decisions flags[v, d - 1] for second = v,
    first = v - d, 1 <= d <= band, within the variant's chromosome (the layout of pl2gpu_ld_band_flags)."""
    m, n = geno.shape
    sex = np.zeros(n, dtype=np.int64) if sex is None else np.asarray(sex)
    majf = major_allele_freqs(ld_ref_freqs_by_class(geno, chrom, sex))
    if preferred is not None:
        majf = np.where(np.asarray(preferred, dtype=bool), majf - 1.0, majf)
    thr = r2_thresh * (1 + SMALL_EPSILON)
    mono = np.zeros(m, dtype=np.uint8)
    flags = np.zeros((m, band), dtype=np.uint8)
    idx_all = np.arange(m)
    for c in dict.fromkeys(chrom.tolist()):
        idx = idx_all[chrom == c]
        g, w = ld_class_block(geno[idx], _chr_class(c), sex)
        wv = np.ones(g.shape[1], dtype=np.int64) if w is None else np.asarray(w, dtype=np.int64)
        nm_ct_v = (g != 3).astype(np.int64) @ wv
        plus_v = (g == 0).astype(np.int64) @ wv
        minus_v = (g == 2).astype(np.int64) @ wv
        mono[idx] = ((plus_v == 0) & (minus_v == 0)) | (plus_v == nm_ct_v) | (minus_v == nm_ct_v)
        x = np.where(g == 0, 1.0, np.where(g == 2, -1.0, 0.0)).astype(np.float32)
        nm = (g != 3).astype(np.float32)
        wf = None if w is None else wv.astype(np.float32)
        for a in range(1, idx.size):
            firsts = np.arange(max(0, a - band), a)
            nm_ct, s_b, q_b, s_a, q_a, dot = ld_pair_components(x, nm, a, firsts, wf)
            cov12 = (dot * nm_ct - s_b * s_a).astype(np.float64)
            var1 = (q_b * nm_ct - s_b * s_b).astype(np.float64)
            var2 = (q_a * nm_ct - s_a * s_a).astype(np.float64)
            flags[idx[a], a - firsts - 1] = cov12 * cov12 > thr * var1 * var2
    return majf, mono, flags


def ld_prune(geno: np.ndarray, chrom: np.ndarray, bps: np.ndarray, window: int, step: int, r2_thresh: float, window_is_bp: bool = False, ref_freq: np.ndarray = None, preferred: np.ndarray = None,
             sex: np.ndarray = None, order: int = 2) -> np.ndarray:
    """LdPrune -> IndepPairwise (2.0/plink2_ld.cc:2530-2724): chr0 variants are dropped up front
    (:2542, reported in neither list), every chromosome (bp windows: every run of variants whose
    gaps are <= window, LdPruneSubcontigSplitAll :2165-2268) with >= 2 variants is an independent
    job; variants in singleton subcontigs are never examined (kept).  Returns removed[M] bool
    (chr0 variants: False).  `geno` holds founders only; `sex` (1 male / 2 female / 0 unknown per founder)
    matters on chrX (males: hets -> missing, weight 1; nonmales weight 2; :1371-1376, :982-998), chrY (nonfemale
    founders only, hets -> missing, :1385-1389) and MT (hets -> missing, :1362-1364).  order = 1: `--indep-order 1`."""
    m, n = geno.shape
    sex = np.zeros(n, dtype=np.int64) if sex is None else np.asarray(sex)
    if ref_freq is None:
        ref_freq = ld_ref_freqs_by_class(geno, chrom, sex)
    majf = major_allele_freqs(ref_freq)
    if preferred is not None:  # --indep-preferred: listed variants win every victim comparison (:916-918)
        majf = np.where(np.asarray(preferred, dtype=bool), majf - 1.0, majf)
    removed = np.zeros(m, dtype=bool)
    idx_all = np.arange(m)
    for c in [c for c in dict.fromkeys(chrom.tolist()) if c not in ("0", 0)]:
        idx = idx_all[chrom == c]
        if idx.size < 2:
            continue
        cls = _chr_class(c)
        groups = []
        if window_is_bp:
            # split where variant_bp - window > previous bp (:2199-2211)
            cur = [idx[0]]
            for k in idx[1:]:
                if int(bps[k]) >= window and int(bps[k]) - window > int(bps[cur[-1]]):
                    groups.append(cur)
                    cur = []
                cur.append(k)
            groups.append(cur)
        else:
            groups = [list(idx)]
        for gidx in groups:
            if len(gidx) < 2:
                continue
            gidx = np.array(gidx)
            g, w = ld_class_block(geno[gidx], cls, sex)
            removed[gidx] = ld_prune_subcontig(g, majf[gidx], bps[gidx] if window_is_bp else None, window, step, r2_thresh, w=w, order=order)
    return removed


def read_bim(path: str):
    chrom, ids, bps = [], [], []
    with open(path) as f:
        for ln in f:
            t = ln.split()
            chrom.append(t[0])
            ids.append(t[1])
            bps.append(int(t[3]))
    return np.array(chrom), ids, np.array(bps, dtype=np.int64)


# --------------------------------------------------------------------------------------------- PCA
def pca_exact(grm_matrix: np.ndarray, pc_ct: int):
    """CalcPca exact branch (2.0/plink2_matrix_calc.cc:5942-6040): top-k eigenpairs, descending."""
    w, v = np.linalg.eigh(grm_matrix)
    return w[::-1][:pc_ct], v[:, ::-1][:, :pc_ct].T


def pca_approx(geno: np.ndarray, pc_ct: int, g1: np.ndarray, ref_freq: np.ndarray = None):
    """CalcPca approx branch (2.0/plink2_matrix_calc.cc:5697-5941) for a given Gaussian start matrix
    g1 [samples, 2k]: k+1 projections H_t = Y G_t with G_{t+1} = Y^T H_t / M kept side by side (qq,
    M x 2k(k+1)); Q = left singular vectors of qq; B = Y^T Q; eigvecs = first k left singular vectors
    of B, eigvals = sigma^2 / M.  Y = centered_varmaj (missing -> 0, always variance-standardised)."""
    if ref_freq is None:
        ref_freq = ref_allele_freqs(geno)
    y = centered_varmaj(geno, ref_freq, True)
    m = y.shape[0]
    c2 = 2 * pc_ct
    qq = np.empty((m, c2 * (pc_ct + 1)))
    g = np.array(g1, dtype=np.float64)
    for it in range(pc_ct + 1):
        h = y @ g
        qq[:, it * c2 : (it + 1) * c2] = h
        if it < pc_ct:
            g = (y.T @ h) * (1.0 / m)
    u, _, _ = np.linalg.svd(qq, full_matrices=False)
    b = y.T @ u
    ub, sb, _ = np.linalg.svd(b, full_matrices=False)
    return sb[:pc_ct] ** 2 * (1.0 / m), ub[:, :pc_ct].T
