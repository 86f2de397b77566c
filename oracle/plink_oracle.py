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
