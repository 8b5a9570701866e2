"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): event ranges, and the property
the sharding relies on -- per-rank results over disjoint event ranges, summed with one all-reduce,
equal the single-process result exactly (counter-based streams keyed by the global event index)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import iss_oracle as orc  # noqa: E402
from iss_b200 import sharding  # noqa: E402

LAM = np.array([0.0, 0.3, 4.5, 167.2, 2251.25, 12.5, 80.0])
NEV = 1001
SEED = 777


def qa_like_block(mult):
    """plain sums, like the QA block: events, per-species sum N and sum N^2, histogram of N"""
    blk = [float(len(mult))]
    blk += list(mult.sum(axis=0).astype(float))
    blk += list((mult.astype(float)**2).sum(axis=0))
    blk += list(np.bincount(np.minimum(mult[:, 2], 31), minlength=32).astype(float))
    return np.array(blk)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.split_events(NEV, world)[rank]
    sp = np.zeros(len(LAM), dtype=orc.OSpecies)
    mult, _ = orc.multiplicities(LAM, orc.poisson_pmode(LAM), sp, e - b, b, SEED)
    t = torch.from_numpy(qa_like_block(mult))
    sharding.allreduce_sum_(t)
    if rank == 0:
        np.save(out, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_split_events_cover_and_disjoint():
    for nev, world in [(10, 1), (10, 3), (1001, 2), (7, 8), (10000, 8)]:
        r = sharding.split_events(nev, world)
        assert r[0][0] == 0 and r[-1][1] == nev
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1
    seen = set()
    for step in range(3):
        for rank in range(4):
            b, e = sharding.weak_event_range(step, rank, 4, 100)
            assert not (seen & set(range(b, e)))
            seen |= set(range(b, e))
    assert seen == set(range(1200))


def test_two_ranks_equal_one_rank(tmp_path):
    out = str(tmp_path/"blk.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    sp = np.zeros(len(LAM), dtype=orc.OSpecies)
    mult, _ = orc.multiplicities(LAM, orc.poisson_pmode(LAM), sp, NEV, 0, SEED)
    assert np.array_equal(got, qa_like_block(mult))


# ---- species sharding of the smooth-spectra integrator (no reduction: a gather of tables)
def _spectra_worker(rank, world, port, out):
    import spectra_oracle as spo
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import spectra_cases as sc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = sc.load("sp3d_shear")
    opt = sc.options_of(g)
    species = sc.species_of(g)[:5]
    pT, phi, eta = sc.bin_tables()
    lab = g["lab"][:12]
    mine = sharding.split_species(len(species), world)[rank]
    local = np.array([spo.spectra(lab, species[k], opt, pT[:4], phi[:6], eta[::8])[0] for k in mine])
    local = local.reshape(len(mine), 4, 6)
    full = sharding.gather_species_tables(torch.from_numpy(local), len(species), world, rank)
    if rank == 0:
        np.save(out, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_split_species_round_robin():
    for ns, world in [(5, 2), (321, 8), (3, 4), (7, 1)]:
        parts = sharding.split_species(ns, world)
        assert sorted(i for p in parts for i in p) == list(range(ns))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_species_sharded_spectra_equal_single_process(tmp_path):
    import socket
    outs = []
    for world in (1, 2):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        out = str(tmp_path/("spectra_w%d.npy" % world))
        mp.spawn(_spectra_worker, args=(world, port, out), nprocs=world, join=True)
        outs.append(np.load(out))
    assert outs[0].shape == (5, 4, 6)
    assert np.array_equal(outs[0], outs[1])       # gathered tables are bit-identical


# ---- surface-chunk sharding: all-gather of tile sums, then the same fixed-order combination
def _chunk_case():
    rng = np.random.default_rng(11)
    ns, ncell = 4, 3*4096 + 1500
    y = rng.random((ns, ncell))*np.where(rng.random((ns, ncell)) < 0.2, 0.0, 1.0)
    y[1] *= 1e-7
    return y


def _chunk_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y = _chunk_case()
    ranges = sharding.split_cells(y.shape[1], world)
    b, e = ranges[rank]
    local, tilesum = orc.engine_tile_scan(y[:, b:e])            # what chunk_yields_local computes
    blocks = sharding.gather_tile_sums(torch.from_numpy(tilesum), ranges, y.shape[0])
    full = np.concatenate([t.numpy() for t in blocks], axis=1)  # what chunk_yields_finish assembles
    base, total = orc.engine_tile_bases(full)
    t0 = b//sharding.TILE
    P = (local + base[:, t0:t0 + local.shape[1], None]).reshape(y.shape[0], -1)[:, :e - b]
    np.savez(out % rank, P=P, total=total, b=b, e=e)
    dist.barrier()
    dist.destroy_process_group()


def test_split_cells_aligned_cover():
    for ncell, world in [(20000, 2), (966924, 8), (10**7, 8), (5000, 2), (4097, 3), (4096, 1)]:
        r = sharding.split_cells(ncell, world)
        assert r[0][0] == 0 and r[-1][1] == ncell
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert all(b % sharding.CHUNK_ALIGN == 0 for b, e in r if e > b)
        assert all(e % sharding.CHUNK_ALIGN == 0 or e == ncell for b, e in r)
        blocks = [-(-(e - b)//sharding.CHUNK_ALIGN) for b, e in r]
        assert max(blocks) - min(blocks) <= 1
        assert sum(sharding.ntiles_of(x) for x in r) == -(-ncell//sharding.TILE)


def test_chunked_prefix_equals_single_process(tmp_path):
    """Two ranks holding half the cells each obtain, after one all-gather of tile sums, the bits of
    the single-process prefix and totals; every draw v is then claimed by exactly one rank and
    resolves to the same cell."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path/"chunk_r%d.npz")
    mp.spawn(_chunk_worker, args=(2, port, out), nprocs=2, join=True)
    y = _chunk_case()
    P, total = orc.engine_prefix(y)
    # the engine's order differs from the reference's sequential sum only by rounding
    assert np.allclose(total, orc.species_totals(y), rtol=1e-13)
    parts = [np.load(out % r) for r in range(2)]
    for p in parts:
        assert np.array_equal(p["total"], total)
        assert np.array_equal(p["P"], P[:, int(p["b"]):int(p["e"])])
    rng = np.random.default_rng(3)
    for s_ in range(y.shape[0]):
        v = (total[s_] - 1e-15)*rng.random(2000)
        cell = np.minimum((P[s_][None, :] < v[:, None]).sum(axis=1), y.shape[1] - 1)
        claimed = np.zeros(len(v), dtype=int)
        for p in parts:
            b, e = int(p["b"]), int(p["e"])
            lo = P[s_, b - 1] if b > 0 else -np.inf
            mine = (v > lo) & ((v <= p["P"][s_, -1]) | (e == y.shape[1]))
            local = (p["P"][s_][None, :] < v[mine, None]).sum(axis=1)
            assert np.array_equal(np.minimum(local + b, y.shape[1] - 1), cell[mine])
            claimed += mine
        assert np.all(claimed == 1)


def test_weighted_cell_split_is_a_partition_on_block_boundaries():
    """sharding.split_cells_weighted: contiguous, aligned, every rank at least one block, and the
    heaviest rank close to the mean for a peaked cost profile"""
    import numpy as np
    from iss_b200 import sharding
    ncell = 966924
    nblk = (ncell + sharding.CHUNK_ALIGN - 1)//sharding.CHUNK_ALIGN
    x = np.linspace(-3, 3, nblk)
    cost = 0.02 + np.exp(-x**2)                 # hot mid-rapidity blocks
    for world in (1, 2, 3, 8):
        r = sharding.split_cells_weighted(cost, ncell, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == ncell
        for (b0, e0), (b1, e1) in zip(r[:-1], r[1:]):
            assert e0 == b1 and e0 % sharding.CHUNK_ALIGN == 0
        assert all(e > b for b, e in r)
        per = [cost[b//sharding.CHUNK_ALIGN:(e + sharding.CHUNK_ALIGN - 1)//sharding.CHUNK_ALIGN].sum()
               for b, e in r]
        assert max(per) < 1.15*cost.sum()/world
    even = sharding.split_cells(ncell, 8)
    per_even = [cost[b//sharding.CHUNK_ALIGN:(e + sharding.CHUNK_ALIGN - 1)//sharding.CHUNK_ALIGN].sum()
                for b, e in even]
    assert max(per_even) > 1.5*cost.sum()/8     # what the weights are for
    # fewer blocks than ranks: the even cut (trailing ranks empty)
    assert sharding.split_cells_weighted(np.ones(2), 8000, 4) == sharding.split_cells(8000, 4)
