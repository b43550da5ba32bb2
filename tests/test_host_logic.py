"""CPU tests of the host-side mirror: reshape/symmetrise helpers against the reference goldens, step adaptation,
and the chain-parallel sharding + final gather over a 2-rank gloo group."""
import os
import sys

import numpy as np
import pytest
import torch

from mspde import parallel, util


def test_util_helpers_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_util_helpers.npz"))
    assert np.array_equal(util.symmetrize(g["half"]), g["sym"])
    assert np.array_equal(util.symmetrize(torch.from_numpy(g["half"])).numpy(), g["sym"])
    assert np.array_equal(util.symmetrize_2D(g["uh"]), g["sym2D"])
    assert np.array_equal(util.symmetrize_2D(torch.from_numpy(g["uh"])).numpy(), g["sym2D"])
    assert np.array_equal(util.extend2D(g["sym2D"], 6), g["ext2D"])
    assert np.array_equal(util.extend2D(torch.from_numpy(g["sym2D"]), 6).numpy(), g["ext2D"])


def test_ravel_reshape_conventions():
    n = 3
    il = 2 * n - 1
    u = np.arange(il * il) + 0j
    ref = u.reshape(il, il, order="F")[:, n - 1:]
    assert np.array_equal(util.from_u_2D_ravel_to_uHalf_2D(u, n), ref)
    assert np.array_equal(util.from_u_2D_ravel_to_uHalf_2D(torch.from_numpy(u), n).numpy(), ref)
    a = np.arange(12).reshape(3, 4)
    assert np.array_equal(util.ravel_F(torch.from_numpy(a)).numpy(), a.ravel("F"))


def test_noise_half_vector():
    from oracle import mspde_oracle as o
    re, im = np.random.default_rng(0).standard_normal((2, 7))
    assert np.array_equal(util.w_half_from_normals(re, im), o.w_half_from_normals(re, im))


def test_sharding():
    for world in (1, 2, 4, 8):
        owned = [parallel.shard_chains(64, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(64))
        assert all(len(x) == 64 // world for x in owned)
    assert parallel.shard_chains(5, 2, 0) == [0, 2, 4] and parallel.shard_chains(5, 2, 1) == [1, 3]
    assert parallel.chain_seed(1, 3) == 4


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_chains = 5
    ids = parallel.shard_chains(n_chains, world, rank)
    local = torch.stack([torch.full((3, 4), complex(c, -c), dtype=torch.complex128) for c in ids])
    full = parallel.gather_samples(local, n_chains, world)
    ok = all(bool((full[c] == complex(c, -c)).all()) for c in range(n_chains))
    acc = parallel.pooled_acceptance(accepted_local=rank + 1, steps_local=10)
    q.put((rank, ok, acc))
    dist.destroy_process_group()


def test_gather_samples_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert all(abs(acc - 3 / 20) < 1e-12 for _, _, acc in res)


def test_block_rows_layout():
    """Ownership / offsets of the 1-D block-cyclic row layout used by the distributed Phi (index logic only, CPU)."""
    from mspde.dist import BlockRows
    n, nb = 1000, 128
    seen = []
    for rank in range(3):
        br = BlockRows(n, nb, 3, rank, torch.device("cpu"))
        assert br.owned == [i for i in range(8) if i % 3 == rank]
        rows = sum(br.rows(i) for i in br.owned)
        assert br.data.shape == (rows, 1000) and br.ld % 8 == 0
        for i in br.owned:
            assert br.block(i).shape == (min(nb, n - i * nb), n)
            br.block(i).fill_(i)
            seen.append(i)
        for i in br.owned:
            assert bool((br.block(i) == i).all())
    assert sorted(seen) == list(range(8))
    assert BlockRows(1000, 128, 3, 0, torch.device("cpu")).rows(7) == 1000 - 7 * 128


def test_block_cols_layout():
    """1-D block-cyclic column layout of the distributed LU / QR (index logic only, CPU)."""
    from mspde.dist import BlockCols
    rows, ntot = 10, 200            # 4 blocks: 64, 64, 64, 8
    full = torch.arange(rows * 199, dtype=torch.float64).reshape(rows, 199).to(torch.complex128)
    aug = torch.full((rows, 1), -1.0, dtype=torch.complex128)
    cover = torch.zeros(ntot)
    for rank in range(3):
        bc = BlockCols(rows, ntot, 3, rank, torch.device("cpu"))
        assert bc.owned == [b for b in range(4) if b % 3 == rank]
        bc.fill_from(full, aug)
        for b in bc.owned:
            g0, w, l0 = b * 64, bc.width(b), bc.col0[b]
            ref = torch.cat((full, aug), 1)[:, g0:g0 + w]
            assert torch.equal(bc.data[:, l0:l0 + w], ref)
            cover[g0:g0 + w] += 1
        start, n2 = bc.right_of(0, 64)
        assert n2 == sum(bc.width(b) for b in bc.owned if b > 0) and (n2 == 0 or start == bc.col0[[b for b in bc.owned if b > 0][0]])
        if 3 in bc.col0:                                   # last block: 7 real columns + the augmented one
            start, n2 = bc.right_of(3, 7)
            assert n2 == 1 and bool((bc.data[:, start] == -1).all())
    assert bool((cover == 1).all())


def test_merge_welford_matches_numpy():
    """Parallel-variance merge of per-chain Welford triples (pooled statistics of Legacy/runNumbaParallel.py:121-136)."""
    from mspde import parallel
    rng = np.random.default_rng(3)
    chunks = [rng.standard_normal((k, 5, 4)) * (1 + i) + i for i, k in enumerate([7, 1, 12, 3])]
    parts = [(len(c), torch.as_tensor(c.mean(0)), torch.as_tensor(((c - c.mean(0)) ** 2).sum(0))) for c in chunks]
    parts.insert(2, (0, torch.zeros(5, 4, dtype=torch.float64), torch.zeros(5, 4, dtype=torch.float64)))   # an empty chain
    n, mean, m2 = parallel.merge_welford(parts)
    allx = np.concatenate(chunks)
    assert n == len(allx)
    assert np.allclose(mean.numpy(), allx.mean(0), rtol=1e-13, atol=1e-13)
    assert np.allclose((m2 / n).numpy(), allx.var(0), rtol=1e-12, atol=1e-13)


def _welford_worker(rank, world, port, q):
    import torch.distributed as dist
    from mspde import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    data = [rng.standard_normal((9, 6)) + 2.0, rng.standard_normal((4, 6)) * 3.0]       # rank 0 owns 9 samples, rank 1 owns 4
    mine = data[rank]
    n, mean, m2 = parallel.pool_welford(len(mine), torch.as_tensor(mine.mean(0)), torch.as_tensor(((mine - mine.mean(0)) ** 2).sum(0)))
    allx = np.concatenate(data)
    ok = (n == 13 and np.allclose(mean.numpy(), allx.mean(0), rtol=1e-13) and np.allclose((m2 / n).numpy(), allx.var(0), rtol=1e-12))
    # a rank without samples still takes part in the collective
    n2, mean2, _ = parallel.pool_welford(len(mine) if rank == 0 else 0, torch.as_tensor(mine.mean(0)), torch.as_tensor(mine.var(0) * len(mine)))
    ok = ok and n2 == 9 and np.allclose(mean2.numpy(), data[0].mean(0), rtol=1e-13)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_pool_welford_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_welford_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


def test_dist_qr_places_every_right_hand_side():
    """Column placement of several right-hand sides in the block-cyclic QR layout (the Gibbs draw carries both accept outcomes)."""
    from mspde.dist import BlockCols
    n, nrhs, rows = 127, 2, 6                      # global columns 127 and 128 fall into different 64-blocks
    for world in (1, 2, 3):
        seen = {}
        for rank in range(world):
            mat = BlockCols(rows, n + nrhs, world, rank, torch.device("cpu"))
            for j in range(nrhs):
                b = (n + j) // BlockCols.NB
                if b in mat.col0:
                    seen[j] = (rank, mat.col0[b] + (n + j - b * BlockCols.NB))
                    assert seen[j][1] < mat.ncols
        assert sorted(seen) == [0, 1]


def test_fbp_reconstructs_a_disc_from_its_analytic_sinogram():
    """The iradon restatement (mspde/io.py:fbp) against a known answer: the Radon transform of a centred disc of radius R and
    unit density is 2 sqrt(R^2 - s^2) for every angle; FBP must give ~1 inside, ~0 outside."""
    from mspde import io
    n_r, n_theta, R = 127, 90, 30.0
    s = np.arange(n_r) - n_r // 2
    col = 2 * np.sqrt(np.clip(R * R - s * s, 0, None))
    sino = np.repeat(col[:, None], n_theta, axis=1)
    rec = io.fbp(sino, np.linspace(0, 180, n_theta, endpoint=False))
    yy, xx = np.mgrid[:n_r, :n_r] - n_r // 2
    rr = np.sqrt(xx ** 2 + yy ** 2)
    assert abs(rec[rr < R - 3].mean() - 1.0) < 0.02 and rec[rr < R - 3].std() < 0.03
    assert abs(rec[(rr > R + 3) & (rr < n_r // 2 - 2)].mean()) < 0.02
