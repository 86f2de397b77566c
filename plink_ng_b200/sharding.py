"""Multi-GPU decomposition of the N x N jobs (host logic shared by bench.py and the tests).

Output rows are split into equal-area blocks with the reference's own `--parallel` arithmetic
(ParallelBounds, 2.0/plink2_common.cc:4956-4961): rank r owns rows [r_start, r_end) of the lower
triangle and writes piece r+1 of n.  The only exchange is the genotype block: each rank decodes /
synthesises a contiguous 1/world slice of the step's variants and ONE all_gather assembles the block
on every rank (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from .host import parallel_bounds


def row_block(sample_ct: int, rank: int, world: int, include_diag: bool = False):
    """Rows [start, end) of rank `rank`: strict lower triangle (KING) or with diagonal (GRM)."""
    return parallel_bounds(sample_ct, 0 if include_diag else 1, rank, world)


def row_block_tiles(sample_ct: int, rank: int, world: int, tile_rows: int = 128, tile_cols: int = 80, include_diag: bool = False):
    """Tile-aligned variant of row_block for the multi-GPU product path: boundaries are multiples of the
    128-row pair tile and balance the number of 128 x 80 pair tiles per rank (what the tensor kernel's time is
    proportional to), so no rank computes a partial row tile twice.  Pieces still concatenate to the full
    triangle in row order, like the reference's --parallel pieces."""
    first = 0 if include_diag else 1
    row_tiles = (sample_ct + tile_rows - 1) // tile_rows
    cum = [0]
    for rt in range(row_tiles):
        row_end = min(sample_ct, (rt + 1) * tile_rows)
        cols = row_end if include_diag else row_end - 1
        cum.append(cum[-1] + (cols + tile_cols - 1) // tile_cols)
    total = cum[-1]

    def bound(k):
        if k == 0:
            return first
        if k == world:
            return sample_ct
        target = total * k / world
        rt = min(range(row_tiles + 1), key=lambda t: abs(cum[t] - target))
        return max(first, min(sample_ct, rt * tile_rows))

    return bound(rank), bound(rank + 1)


def variant_slice(variant_ct: int, rank: int, world: int):
    """(per_rank, v0, v1): every rank contributes `per_rank` rows to the gather (the last ones padded);
    rank holds variants [v0, v1) of the step."""
    per = (variant_ct + world - 1) // world
    return per, min(variant_ct, rank * per), min(variant_ct, (rank + 1) * per)


def assemble_block(dist, torch, local_rows, per_rank: int, world: int):
    """all_gather of the per-rank variant slices -> [per_rank * world, row_bytes] on every rank.
    `local_rows` is a uint8 tensor [per_rank, row_bytes] (rows beyond the rank's slice are padding)."""
    if world == 1:
        return local_rows
    full = torch.empty((per_rank * world, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(full.view(-1), local_rows.contiguous().view(-1))
    return full


def pairs_in_rows(r0: int, r1: int, include_diag: bool = False) -> int:
    if include_diag:
        return (r1 * (r1 + 1) - r0 * (r0 + 1)) // 2
    tri = lambda r: r * (r - 1) // 2 if r else 0  # noqa: E731
    return tri(r1) - tri(r0)
