"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0 (yields ->
multiplicities -> sampler -> decays -> QA through the drop-in facade and the C ABI), checked
against the oracle."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))


def run():
    import cases
    import iss_oracle as orc
    from iss_b200 import capi
    from test_oracle_cpu import mode_of

    g = cases.load("s3d_ce_diff")
    d = tempfile.mkdtemp(prefix="iss_smoke_")
    param, surf, over = cases.materialise(g, d)
    s = capi.Sampler(d, param, surf, **over)
    assert s.read_in_FO_surface() == 0
    s.set_random_seed(1)
    assert s.prepare_sampler() == 0
    e = s.engine()
    dN, y = e.compute_yields(want_cells=True)
    ref = g["yields"]
    nz = ref != 0
    err = np.abs(y[nz] - ref[nz])/np.abs(ref[nz])
    assert err.max() < 1e-6, err.max()
    nev = 500
    e.set_trace(True)
    e.sample(3, 0, nev)
    had = e.fetch_all()
    cell, tries = e.get_trace(len(had))
    lam, pm = e.poisson_params()
    sp = s.species()
    m = mode_of(g)
    omult, ocount = orc.multiplicities(lam, pm, sp, nev, 0, 3)
    assert np.array_equal(e.multiplicities(nev), omult)
    lrf = s.lrf_surface()
    tabs = orc.Tables(afterburner=m["afterburner"], kind=m["kind"], include_bulk=m["include_bulk"],
                      include_diff=m["include_diff"])
    coef = orc.cell_coefficients(lrf, tabs, m["kind"], m["include_bulk"], m["include_diff"])
    opt = orc.make_options(hydro_mode=2, include_shear=m["include_shear"],
                           include_bulk=m["include_bulk"], include_diff=m["include_diff"],
                           bulk_kind=m["kind"])
    ohad, ocell, otries = orc.sample(lrf, coef, y, sp, opt, 3, 0, omult, ocount.sum())
    assert len(ohad) == len(had)
    same = (cell == ocell) & (tries == otries)
    assert same.mean() > 0.999
    for f in ("E", "px", "py", "pz", "t", "z"):
        assert np.allclose(had[f][same], ohad[f][same], rtol=2e-6, atol=1e-6)
    qa = e.histograms([211, 2212])
    assert qa[0] == nev and qa[25] == len(had)
    s.close()
    print("smoke ok: %d cells x %d species yields max rel err %.2e; %d hadrons match the oracle"
          % (y.shape[1], y.shape[0], err.max(), len(had)))
