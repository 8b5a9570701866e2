"""BASELINE.json configs[4] size (C5: 10^7 cells, 321 species) through size-independent properties:
integer bookkeeping, mass shell, bit-reproducibility under event splits, and the surface-chunk
identity (two ranks emulated on one GPU: species totals with the same bits as the whole-surface
run, the union of the ranks' lists = the whole-surface list).  The per-cell arithmetic is pinned
against the reference at small sizes (tests/test_yields_gpu.py); this file covers what changes with
size: 64-bit offsets, the five-level cell search over 2 x 26 GB of yields / prefix, the tile sums
of 9.8e3 tiles per species."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from iss_b200 import sharding  # noqa: E402

pytestmark = pytest.mark.gpu
NCELL = 10000000
SEED, EV0, NEV = 31337, 3, 10


@pytest.fixture(scope="module")
def c5(built, tmp_path_factory):
    capi = built
    work = str(tmp_path_factory.mktemp("c5"))
    bench.make_case(work, NCELL, "c5")
    s = capi.Sampler(work, bench.PARAM, "surface.dat",
                     **dict(bench.overrides_of("c5"), number_of_repeated_sampling=NEV))
    assert s.read_in_FO_surface() == 0
    os.remove(os.path.join(work, "surface.dat"))     # 1.4 GB
    s.set_random_seed(1)
    assert s.prepare_sampler() == 0
    yield capi, s
    s.close()


def test_c5_bookkeeping_determinism_and_chunks(c5):
    capi, s = c5
    e = s.engine()
    lrf = s.lrf_surface()
    assert len(lrf) > 9000000
    dN = e.compute_yields().copy()
    assert np.all(dN >= 0) and dN.sum() > 4e5       # ~5.5e5 hadrons per event
    e.set_trace(True)
    c = e.sample(SEED, EV0, EV0 + NEV)
    mult = e.multiplicities(NEV).copy()
    off = e.event_offsets(NEV).copy()
    had = e.fetch_all().copy()
    cell, tries = e.get_trace(len(had))
    assert c.n_hadrons == len(had) == mult.sum() == off[-1] > 4000000
    assert np.array_equal(np.diff(off), mult.sum(axis=1))
    assert abs(len(had)/NEV - dN.sum()) < 6*np.sqrt(dN.sum()/NEV)
    assert cell.min() >= 0 and cell.max() < len(lrf) and cell.max() > 0.99*len(lrf)
    sp = s.species()
    assert np.array_equal(had[off[0]:off[1]]["pid"], np.repeat(sp["pid"], mult[0]))
    E, px, py, pz, m = (had[k].astype(np.float64) for k in ("E", "px", "py", "pz", "mass"))
    assert np.abs(E*E - px*px - py*py - pz*pz - m*m).max() < 2e-5*np.max(E*E)
    # positions are those of the chosen cells
    sub = np.arange(0, len(had), 4001)
    assert np.array_equal(had["x"][sub], lrf[cell[sub], 1]) and np.array_equal(had["y"][sub], lrf[cell[sub], 2])
    # any split of the event range gives the same bytes
    parts = []
    for a, b in ((EV0, EV0 + 3), (EV0 + 3, EV0 + NEV)):
        e.sample(SEED, a, b)
        parts.append(e.fetch_all().copy())
    assert np.concatenate(parts).tobytes() == had.tobytes()

    # two ranks, 5e6 cells each, emulated one after the other on this handle
    lrf = lrf.copy()
    ranges = sharding.split_cells(len(lrf), 2)
    blocks = []
    for b, en in ranges:
        e.upload_surface(lrf[b:en])
        e.set_surface_chunk(b, len(lrf))
        blocks.append(e.chunk_tilesums_host())
    ntiles = [blk.shape[1] for blk in blocks]
    total = 0
    for b, en in ranges:
        e.upload_surface(lrf[b:en])
        e.set_surface_chunk(b, len(lrf))
        e.chunk_yields_local()
        dN_r = e.chunk_yields_finish(blocks, ntiles, on_device=False)
        assert np.array_equal(dN_r, dN)             # totals of the WHOLE surface, same bits
        e.set_trace(True)
        c_r = e.sample(SEED, EV0, EV0 + NEV)
        had_r = e.fetch_all().copy()
        cell_r, tries_r = e.get_trace(len(had_r))
        mine = (cell >= b) & (cell < en)
        assert c_r.n_hadrons == mine.sum()
        assert np.array_equal(cell_r, cell[mine]) and np.array_equal(tries_r, tries[mine])
        assert had_r.tobytes() == had[mine].tobytes()
        total += len(had_r)
    assert total == len(had)
