"""Surface-chunk sharding (include/iss_cuda.h) against the single-GPU run of
the same surface.  The ranks are emulated one after the other with ONE handle on one GPU: the same
entry points and kernels a multi-GPU run uses, with the all-gather of the tile sums done through
host memory (the NCCL variant is iss_b200/sharding.py::chunk_yields, exercised by
`tools/chunk_probe.py` under torchrun).

Bar: bit-exact.  Species totals, multiplicities, chosen cells, numbers of tries and every hadron
record of every rank must equal the single-GPU run; a rank's list is exactly the sub-list of the
single-GPU list whose cells lie in the rank's range, in the same order."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import bench  # noqa: E402
import iss_oracle as orc  # noqa: E402
from iss_b200 import sharding  # noqa: E402

pytestmark = pytest.mark.gpu

NCELL = 30000
SEED = 4242
EV0, NEV = 17, 40


@pytest.fixture(scope="module")
def surf(built, tmp_path_factory):
    capi = built
    work = str(tmp_path_factory.mktemp("chunk"))
    bench.make_case(work, NCELL)
    s = capi.Sampler(work, bench.PARAM, "surface.dat",
                     **dict(bench.OVERRIDES, number_of_repeated_sampling=NEV))
    assert s.read_in_FO_surface() == 0
    s.set_random_seed(1)
    assert s.prepare_sampler() == 0
    yield capi, s, s.lrf_surface().copy()
    s.close()


def whole_run(e, lrf, nev=NEV, decay=False):
    e.upload_surface(lrf)
    dN, y = e.compute_yields(want_cells=True)
    e.set_trace(True)
    c = e.sample(SEED, EV0, EV0 + nev)
    out = dict(dN=dN.copy(), y=y, mult=e.multiplicities(nev).copy(), off=e.event_offsets(nev).copy(),
               had=e.fetch_all().copy(), counts=(c.n_hadrons, c.n_tries, c.n_cell_redraws))
    out["cell"], out["tries"] = e.get_trace(len(out["had"]))
    return out


def chunk_runs(e, lrf, world, nev=NEV):
    """every rank of a `world`-rank surface-chunk run, one after the other"""
    ranges = [r for r in sharding.split_cells(len(lrf), world) if r[1] > r[0]]
    blocks = []
    for b, en in ranges:
        e.upload_surface(lrf[b:en])
        e.set_surface_chunk(b, len(lrf))
        blocks.append(e.chunk_tilesums_host())
    ntiles = [blk.shape[1] for blk in blocks]
    assert ntiles == [sharding.ntiles_of(r) for r in ranges]
    outs = []
    for b, en in ranges:
        e.upload_surface(lrf[b:en])
        e.set_surface_chunk(b, len(lrf))
        e.chunk_yields_local()
        dN = e.chunk_yields_finish(blocks, ntiles, on_device=False)
        e.set_trace(True)
        c = e.sample(SEED, EV0, EV0 + nev)
        o = dict(dN=dN.copy(), mult=e.multiplicities(nev).copy(), off=e.event_offsets(nev).copy(),
                 had=e.fetch_all().copy(), counts=(c.n_hadrons, c.n_tries, c.n_cell_redraws),
                 range=(b, en))
        o["cell"], o["tries"] = e.get_trace(len(o["had"]))
        outs.append(o)
    return outs


@pytest.mark.parametrize("world", [2, 3, 7])
def test_chunk_ranks_reproduce_single_gpu(surf, world):
    capi, s, lrf = surf
    e = s.engine()
    assert len(lrf) > 6*sharding.CHUNK_ALIGN
    ref = whole_run(e, lrf)
    # the device's fixed-order sums, restated in numpy from the device's own per-cell yields
    assert np.array_equal(ref["dN"], orc.engine_prefix(ref["y"])[1])
    outs = chunk_runs(e, lrf, world)
    ev_of = np.repeat(np.arange(NEV), np.diff(ref["off"]))
    n_sum = 0
    for o in outs:
        b, en = o["range"]
        assert np.array_equal(o["dN"], ref["dN"])            # totals of the WHOLE surface, same bits
        assert np.array_equal(o["mult"], ref["mult"])        # hence the same Poisson draws
        mine = (ref["cell"] >= b) & (ref["cell"] < en)
        assert o["counts"][0] == len(o["had"]) == mine.sum() == o["off"][-1]
        assert np.array_equal(o["cell"], ref["cell"][mine])
        assert np.array_equal(o["tries"], ref["tries"][mine])
        assert o["had"].tobytes() == ref["had"][mine].tobytes()
        assert np.array_equal(o["off"], np.concatenate([[0], np.cumsum(np.bincount(ev_of[mine],
                                                                                  minlength=NEV))]))
        n_sum += len(o["had"])
    assert n_sum == len(ref["had"])
    assert sum(o["counts"][1] for o in outs) == ref["counts"][1]
    # back to a whole surface on the same handle
    e.upload_surface(lrf)
    assert np.array_equal(e.compute_yields(), ref["dN"])


def test_chunk_mode_with_charge_pairing_and_qa(surf):
    """local charge conservation emits the partner from the same cell (FSSW.cpp:1035-1048), hence
    on the same rank; the additive part of the QA block adds up over the ranks."""
    capi, s, lrf = surf
    e = s.engine()
    opt = dict(hydro_mode=2, include_deltaf_shear=1, include_deltaf_bulk=1, include_deltaf_diffusion=1,
               bulk_deltaf_kind=21, local_charge_conservation=1)
    e.set_options(**opt)
    try:
        pids = [211, 2212]
        ref = whole_run(e, lrf, nev=20)
        qa_ref = e.histograms(pids)
        outs = chunk_runs(e, lrf, 2, nev=20)
        qa = np.zeros_like(qa_ref)
        for o in outs:
            b, en = o["range"]
            mine = (ref["cell"] >= b) & (ref["cell"] < en)
            assert o["had"].tobytes() == ref["had"][mine].tobytes()
        # QA of the second rank's batch is still on the handle: accumulate rank by rank instead
        ranges = [o["range"] for o in outs]
        blocks = []
        for b, en in ranges:
            e.upload_surface(lrf[b:en])
            e.set_surface_chunk(b, len(lrf))
            blocks.append(e.chunk_tilesums_host())
        for b, en in ranges:
            e.upload_surface(lrf[b:en])
            e.set_surface_chunk(b, len(lrf))
            e.chunk_yields_local()
            e.chunk_yields_finish(blocks, [x.shape[1] for x in blocks], on_device=False)
            e.sample(SEED, EV0, EV0 + 20)
            qa += e.histograms(pids)
        add = np.r_[9:29]                                   # sum p^mu p^nu / p^0, counts, net charges
        assert np.allclose(qa[add], qa_ref[add], rtol=1e-12, atol=1e-9)
        H, PER, NPT = capi.QA_HEAD, capi.QA_PER, capi.QA_NPT
        for k in range(2):
            blk, blk_ref = qa[H + k*PER:H + (k + 1)*PER], qa_ref[H + k*PER:H + (k + 1)*PER]
            assert np.array_equal(blk[:NPT], blk_ref[:NPT])                   # pT counts
            assert np.array_equal(blk[3*NPT:3*NPT + capi.QA_NY], blk_ref[3*NPT:3*NPT + capi.QA_NY])
            assert blk[-2] == blk_ref[-2]                                     # n_total
        assert len(ref["had"]) == sum(len(o["had"]) for o in outs)
    finally:
        e.set_options(**dict(opt, local_charge_conservation=0))
        e.upload_surface(lrf)
        e.compute_yields()


def test_allgather_entry_point_and_block_yields(surf):
    """iss_cuda_chunk_yields_allgather with one rank (no communicator needed: the same code path minus
    the NCCL call; the multi-rank case runs under torchrun, tests/test_multigpu_gpu.py) and
    iss_cuda_chunk_block_yields, the input of sharding.split_cells_weighted"""
    capi, s, lrf = surf
    e = s.engine()
    ref = whole_run(e, lrf)
    try:
        e.upload_surface(lrf)
        e.set_surface_chunk(0, len(lrf))
        dN = e.chunk_yields_allgather([sharding.ntiles_of((0, len(lrf)))])
        assert np.array_equal(dN, ref["dN"])
        by = e.chunk_block_yields(len(lrf))
        blk = np.add.reduceat(ref["y"].sum(axis=0), np.arange(0, len(lrf), sharding.CHUNK_ALIGN))
        assert len(by) == len(blk)
        np.testing.assert_allclose(by, blk, rtol=1e-10)
        cut = sharding.split_cells_weighted(by, len(lrf), 3)
        assert cut[0][0] == 0 and cut[-1][1] == len(lrf) and all(b < en for b, en in cut)
        e.set_trace(True)
        c = e.sample(SEED, EV0, EV0 + NEV)
        assert c.n_hadrons == len(ref["had"])
        assert e.fetch_all().tobytes() == ref["had"].tobytes()
        with pytest.raises(capi.IssError):
            e.chunk_yields_allgather([3, 4])                # two ranks, no communicator
    finally:
        e.upload_surface(lrf)
        e.compute_yields()


def test_chunk_argument_checks(surf):
    capi, s, lrf = surf
    e = s.engine()
    e.upload_surface(lrf[:5000])
    with pytest.raises(capi.IssError):
        e.set_surface_chunk(100, len(lrf))                  # unaligned begin
    with pytest.raises(capi.IssError):
        e.set_surface_chunk(0, len(lrf))                    # ends inside the surface, unaligned
    with pytest.raises(capi.IssError):
        e.chunk_yields_local()                              # mode not set
    e.upload_surface(lrf[:8192])
    e.set_surface_chunk(0, len(lrf))
    with pytest.raises(capi.IssError):
        e.compute_yields()                                  # whole-surface call in chunk mode
    with pytest.raises(capi.IssError):
        e.sample(1, 0, 1)                                   # no yields yet
    e.chunk_yields_local()
    blk = e.chunk_tilesums_host()
    with pytest.raises(capi.IssError):
        e.chunk_yields_finish([blk], [blk.shape[1]], on_device=False)   # tiles do not cover the surface
    e.set_surface_chunk(0, 0)                               # off again
    e.upload_surface(lrf)
    e.compute_yields()


def test_chunk_ragged_tail_and_decays(surf):
    """a last rank that holds ONE cell (surface of 3 x 4096 + 1 cells on 4 ranks), and resonance
    decays on the ranks' own primaries: every rank's final list keeps charge, and the ranks' primary
    lists still tile the single-GPU list."""
    capi, s, lrf = surf
    e = s.engine()
    sub = lrf[:3*sharding.CHUNK_ALIGN + 1]
    ranges = sharding.split_cells(len(sub), 4)
    assert ranges[-1] == (3*sharding.CHUNK_ALIGN, 3*sharding.CHUNK_ALIGN + 1)
    ref = whole_run(e, sub, nev=30)
    outs = chunk_runs(e, sub, 4, nev=30)
    assert len(outs) == 4
    for o in outs:
        b, en = o["range"]
        mine = (ref["cell"] >= b) & (ref["cell"] < en)
        assert np.array_equal(o["dN"], ref["dN"])
        assert o["had"].tobytes() == ref["had"][mine].tobytes()
    assert sum(len(o["had"]) for o in outs) == len(ref["had"])
    # decays of the last configured rank's batch (the handle still holds it)
    prim = outs[-1]["had"]
    c = e.decay(SEED)
    fin = e.fetch_all()
    assert c.n_hadrons == len(fin) >= len(prim)
    e.upload_surface(lrf)
    e.compute_yields()


def test_surface_upload_in_parts_equals_whole(surf):
    """iss_cuda_upload_surface_aos_part (host packs the next part while one is copied): same
    device surface, hence the same yields bit for bit and the same hadrons."""
    capi, s, lrf = surf
    e = s.engine()
    e.upload_surface(lrf)
    dN0, y0 = e.compute_yields(want_cells=True)
    e.sample(SEED, 0, 5)
    h0 = e.fetch_all().copy()
    for nparts in (1, 3, 7):
        e.upload_surface_parts(lrf, nparts)
        dN1, y1 = e.compute_yields(want_cells=True)
        assert np.array_equal(dN0, dN1) and np.array_equal(y0, y1)
    e.sample(SEED, 0, 5)
    assert e.fetch_all().tobytes() == h0.tobytes()
    with pytest.raises(capi.IssError):
        e.L.iss_cuda_upload_surface_aos_part.restype  # noqa: B018  (binding exists)
        e.check(e.L.iss_cuda_upload_surface_aos_part(e.h, capi._ptr(lrf[10:20].copy()), 10, 10, 12345),
                "part without a first part of that surface")
    e.upload_surface(lrf)
    e.compute_yields()
