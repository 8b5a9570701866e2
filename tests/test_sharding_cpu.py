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
