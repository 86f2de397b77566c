"""CPU, world_size 2 over gloo: the N>1 host logic - equal-area row blocks (the reference's
--parallel pieces) cover the triangle exactly once, and the per-rank variant slices all_gather into
the same genotype block a single process would have produced."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, n, mb, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import synth_genovecs
    from plink_ng_b200.sharding import assemble_block, row_block, variant_slice

    per, v0, v1 = variant_slice(mb, rank, world)
    row_bytes = (n + 31) // 32 * 8
    local = torch.zeros((per, row_bytes), dtype=torch.uint8)
    if v1 > v0:
        local[: v1 - v0] = synth_genovecs(torch, n, v0, v1, "cpu")
    full = assemble_block(dist, torch, local, per, world)
    r0, r1 = row_block(n, rank, world)
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full[:mb].numpy())
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), np.array([r0, r1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_all_gather_and_row_blocks(tmp_path):
    from bench import synth_genovecs
    from plink_ng_b200.sharding import pairs_in_rows

    n, mb, world = 150, 37, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, mb, str(tmp_path)), nprocs=world, join=True)
    # each rank generated its slice in its own chunks; the assembled block must equal the slices of a
    # single-process run generated the same way
    per = (mb + world - 1) // world
    want = torch.cat([synth_genovecs(torch, n, r * per, min(mb, (r + 1) * per), "cpu") for r in range(world)]).numpy()
    blocks = [np.load(tmp_path / f"full_{r}.npy") for r in range(world)]
    assert all(np.array_equal(b, want) for b in blocks)
    rows = [np.load(tmp_path / f"rows_{r}.npy") for r in range(world)]
    assert rows[0][0] == 1 and rows[-1][1] == n and rows[0][1] == rows[1][0]
    assert sum(pairs_in_rows(int(a), int(b)) for a, b in rows) == n * (n - 1) // 2
    areas = [pairs_in_rows(int(a), int(b)) for a, b in rows]
    assert abs(areas[0] - areas[1]) <= 2 * n  # equal-area split up to one row each side


def test_row_blocks_many_world_sizes():
    from plink_ng_b200.sharding import pairs_in_rows, row_block

    for n in (2, 97, 1000, 100000):
        for world in (1, 2, 4, 8):
            blocks = [row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 1 and blocks[-1][1] == n
            assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
            assert sum(pairs_in_rows(a, b) for a, b in blocks) == n * (n - 1) // 2
            g = [row_block(n, r, world, include_diag=True) for r in range(world)]
            assert g[0][0] == 0 and g[-1][1] == n
            assert sum(pairs_in_rows(a, b, True) for a, b in g) == n * (n + 1) // 2


def test_tile_aligned_row_blocks_cover_triangle_and_balance():
    """row_block_tiles (the multi-GPU product split): contiguous, tile-aligned, complete, and balanced in pair tiles."""
    from plink_ng_b200.sharding import pairs_in_rows, row_block_tiles

    for n in (300, 4096, 100000):
        for world in (1, 2, 4, 8):
            blocks = [row_block_tiles(n, r, world) for r in range(world)]
            assert blocks[0][0] == 1 and blocks[-1][1] == n
            assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
            assert all(b[1] % 128 == 0 or b[1] in (1, n) for b in blocks[:-1])
            assert sum(pairs_in_rows(a, b) for a, b in blocks) == n * (n - 1) // 2
            if n == 100000:
                areas = [pairs_in_rows(a, b) for a, b in blocks]
                assert max(areas) / (sum(areas) / world) < 1.01
