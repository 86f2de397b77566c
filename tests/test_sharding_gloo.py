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


def _pca_worker(rank, world, port, out_dir):
    """Variant-sharded approx PCA (pl2gpu_pca_run_sharded's decomposition) restated in numpy over gloo: H_t stays on the
    rank that owns the variants, G' = Y^T H is an all-reduce of the N x 2k matrix, the Krylov basis is orthonormalised
    block by block with all-reduced Gram-Schmidt coefficients and an all-gathered block, B = Y^T Q is an all-reduce."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import plink_oracle as orc

    rng = np.random.default_rng(0)
    n, m, k = 120, 700, 3
    geno = rng.choice(4, size=(m, n), p=[0.45, 0.35, 0.17, 0.03]).astype(np.uint8)
    g1 = np.random.default_rng(1).standard_normal((n, 2 * k))
    y = orc.centered_varmaj(geno, orc.ref_allele_freqs(geno), True)
    shard = ((m + world - 1) // world + 127) // 128 * 128
    ys = y[rank * shard : min(m, (rank + 1) * shard)]
    c2 = 2 * k

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    g = g1.copy()
    blocks = []
    for it in range(k + 1):
        h = ys @ g
        blocks.append(h)
        if it < k:
            g = allreduce(ys.T @ h) * (1.0 / m)
    q_blocks = []
    for t, w in enumerate(blocks):
        w = w.copy()
        for _ in range(3):
            if q_blocks:
                qp = np.concatenate(q_blocks, axis=1)
                w = w - qp @ allreduce(qp.T @ w)
            # all-gather the block (padded to the largest shard), the same SVD on every rank, own rows back
            slot = torch.zeros((world, shard, c2), dtype=torch.float64)
            mine = torch.zeros((shard, c2), dtype=torch.float64)
            mine[: w.shape[0]] = torch.from_numpy(w)
            dist.all_gather(list(slot.unbind(0)), mine)
            full = slot.reshape(world * shard, c2).numpy()
            u, _, _ = np.linalg.svd(full, full_matrices=False)
            w = u[rank * shard : rank * shard + w.shape[0]]
        q_blocks.append(w)
    q = np.concatenate(q_blocks, axis=1)
    b = allreduce(ys.T @ q)
    ub, sb, _ = np.linalg.svd(b, full_matrices=False)
    np.save(os.path.join(out_dir, f"vals_{rank}.npy"), sb[:k] ** 2 / m)
    if rank == 0:
        want, _ = orc.pca_approx(geno, k, g1)
        np.save(os.path.join(out_dir, "want.npy"), want)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_variant_sharded_pca_decomposition(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_pca_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    vals = [np.load(tmp_path / f"vals_{r}.npy") for r in range(world)]
    want = np.load(tmp_path / "want.npy")
    assert np.array_equal(vals[0], vals[1])  # every rank ends with the same numbers
    assert np.allclose(vals[0], want, rtol=1e-6)
