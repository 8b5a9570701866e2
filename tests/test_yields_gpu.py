"""GPU parity of the yield path against the reference itself.

Golden values = FSSW::calculate_dN_dxtdy_for_one_particle_species (FSSW.cpp:565-715) of the
unmodified reference, dumped by oracle/ref_driver.cpp (tests/golden/make_golden.py).  The engine
is driven through the reference-facing facade (class iSS via include/iss_host.h) for ingest and
through the C ABI (include/iss_cuda.h) for the yields.  Tolerance: 1e-6 relative (north_star),
in FP64; observed agreement is ~1e-12."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.mark.parametrize("name", cases.ONE_CELL + cases.SYNTH)
def test_yields_match_reference(name, built, tmp_path):
    capi = built
    g = cases.load(name)
    param, surf, over = cases.materialise(g, str(tmp_path))
    s = capi.Sampler(str(tmp_path), param, surf, table_path=cases.tables_for(g), **over)
    try:
        assert s.read_in_FO_surface() == 0
        s.set_random_seed(1)
        lrf = s.lrf_surface()
        # ingest is bit-exact (float32 records)
        assert lrf.shape == g["lrf"].shape
        assert np.array_equal(lrf.view(np.uint32), g["lrf"].view(np.uint32))
        assert s.prepare_sampler() == 0
        sp = s.species()
        assert np.array_equal(sp["pid"], g["species"][:, 0].astype(np.int64))
        e = s.engine()
        dN, y = e.compute_yields(want_cells=True)
        ref = g["yields"]
        assert y.shape == ref.shape
        scale = np.maximum(np.abs(ref), 1e-300)
        err = np.abs(y - ref)/scale
        # cells with zero yield (clamped) must be zero on both sides
        assert np.array_equal(ref == 0.0, y == 0.0)
        assert err[ref != 0].max() < RTOL, "max rel err %g" % err[ref != 0].max()
        tot = ref.sum(axis=1)
        assert np.all(np.abs(dN - tot) <= RTOL*np.abs(tot))
    finally:
        s.close()
