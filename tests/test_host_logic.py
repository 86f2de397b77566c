"""CPU: host-side helpers that mirror reference conventions."""
import numpy as np

from plink_ng_b200.host import pack_genotypes, parallel_bounds, unpack_genotypes


def test_pack_roundtrip_and_layout():
    rng = np.random.default_rng(1)
    for n in (1, 15, 16, 31, 32, 33, 100, 129):
        g = rng.integers(0, 4, size=(7, n), dtype=np.uint8)
        gv = pack_genotypes(g)
        assert gv.shape == (7, (n + 31) // 32) and gv.dtype == np.uint64
        assert np.array_equal(unpack_genotypes(gv, n), g)
        # sample s at bits 2*(s%32) of word s/32 (pgenlib nypvec layout)
        s = n - 1
        assert int(gv[3, s // 32] >> np.uint64(2 * (s % 32))) & 3 == g[3, s]


def test_parallel_bounds_matches_reference_piece():
    # golden: `--parallel 2 3` on 100 samples wrote rows per34..per67? derive from the file instead
    pieces = [parallel_bounds(100, 1, k, 3) for k in range(3)]
    assert pieces[0][0] == 1 and pieces[-1][1] == 100
    assert all(pieces[k][1] == pieces[k + 1][0] for k in range(2))
    areas = [sum(range(a, b)) for a, b in pieces]
    assert max(areas) - min(areas) < 200
