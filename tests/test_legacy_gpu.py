"""GPU parity of the legacy "conventional" sampler (MC_sampling = 2,
EmissionFunctionArray::sample_using_dN_dxtdy_4all_particles_conventional,
src/emissionfunction.cpp:3273-3623) through the C ABI (iss_cuda_legacy_*):

  - per-cell x per-species yields and estimate_maximum values against dumps of the compiled
    reference (tests/golden/legacy_*.npz): 1e-10 of the species' largest yield (north_star asks
    1e-6), maxima at 1e-9 relative;
  - hadron by hadron against the CPU restatement (oracle/iss_oracle.c: oracle_legacy_sample) driven
    by the same Philox streams: chosen cells and numbers of tries bit-exact, records within 4
    float32 ulp (two maths libraries);
  - spectra against the reference's own MC_sampling = 2 samples (chi2, >= 10^4 reference events)."""
import numpy as np
import pytest
from scipy import stats

import cases
import legacy_cases as lc
import obs
import spectra_cases as sc
from legacy_cases import lgo, orc
from test_sampler_gpu import compare_hadrons

pytestmark = pytest.mark.gpu


def make_engine(capi, g, par):
    e = capi.Engine(0)
    sp = lc.species_array(g["species"])
    e.upload_species(sp)
    if int(par["include_deltaf_diffusion"]) == 1:
        e.upload_table(capi.TABLE_KAPPA_B, sc.kappa_table(), 150, 100, [0.05, 0.001, 0.0, 0.007892])
    if int(par["include_deltaf_bulk"]) == 1 and int(par["bulk_deltaf_kind"]) == 0:
        tb = np.ascontiguousarray(lgo.load_bulk14())
        e.upload_table(capi.TABLE_BULK14, tb, len(tb), 4)
    e.set_options(hydro_mode=int(par["hydro_mode"]), y_LB=par["y_lb"], y_RB=par["y_rb"],
                  dN_dy_sampling_model=30,
                  local_charge_conservation=int(par.get("local_charge_conservation", 0)))
    e.legacy_setup(g["lab"], g["pos"], lgo.load_z_table(), **lc.engine_options(par))
    return e, sp


@pytest.mark.parametrize("name", lc.YIELD_CASES)
def test_legacy_yields_and_maxima_match_reference(name, built):
    capi = built
    g = cases.load(name, "legacy")
    par = lc.parameters(g)
    e, sp = make_engine(capi, g, par)
    try:
        dN, y, mx = e.legacy_compute_yields(want_cells=True, want_maximum=True)
        ref = g["yields"]
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(y - ref)/scale).max() < 1e-10
        assert np.allclose(dN, np.maximum(ref, 0).sum(axis=1), rtol=1e-10, atol=1e-300)
        assert np.allclose(mx, g["maximum"], rtol=1e-9, atol=0)
    finally:
        e.close()


@pytest.mark.parametrize("name,nev,extra", [
    ("l3d_shear", 300, {}), ("l3d_bulk1_diff", 300, {}), ("l2d_ideal_smash", 40, {}),
    ("l3d_bulk2", 300, {}), ("l3d_bulk3_norestrict", 300, {}), ("l3d_bulk4_boltzmann", 300, {}),
    ("l3d_bulk0", 300, {}),
    ("l3d_shear", 300, {"local_charge_conservation": 1}),
    ("l2d_ideal_smash", 40, {"local_charge_conservation": 1})])
def test_legacy_hadrons_match_oracle(name, nev, extra, built):
    capi = built
    g = cases.load(name, "legacy")
    par = lc.parameters(g)
    par.update(extra)
    lcc = int(par.get("local_charge_conservation", 0))
    e, sp = make_engine(capi, g, par)
    try:
        dN, y, _ = e.legacy_compute_yields(want_cells=True)
        e.set_trace(True)
        seed, ev0 = 987654321, 3
        cnt = e.sample(seed, ev0, ev0 + nev)
        mult = e.multiplicities(nev)
        off = e.event_offsets(nev)
        had = e.fetch_all()
        cell, tries = e.get_trace(len(had))
        lam, pm = e.poisson_params()
        boost_inv = int(par["hydro_mode"]) != 2
        assert np.array_equal(lam, dN*(par["y_rb"] - par["y_lb"]) if boost_inv else dN)
        omult, ocount = orc.multiplicities(lam, pm, sp, nev, ev0, seed, lcc=lcc)
        assert np.array_equal(mult, omult)
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(ocount.sum(axis=1))]))
        assert cnt.n_hadrons == ocount.sum() == len(had) > 300

        opt = lc.oracle_options(par)
        coef = lgo.cell_coefficients(g["lab"], opt, lgo.load_kappa())
        ohad, ocell, otries = lgo.sample(g["lab"], g["pos"], coef, y, sp, opt, lgo.load_z_table(),
                                         seed, ev0, omult, ocount.sum() + 8)
        assert len(ohad) == len(had)
        if lcc:
            # every positive hadron is followed by its conjugate from the same cell
            pos_charge = np.isin(had["pid"], sp["pid"][sp["charge"] > 0])
            idx = np.nonzero(pos_charge)[0]
            assert len(idx) > 100
            assert np.array_equal(had["pid"][idx + 1], -had["pid"][idx])
            assert np.array_equal(cell[idx + 1], cell[idx])
            assert not np.isin(had["pid"][np.setdiff1d(np.arange(len(had)), idx + 1)],
                               sp["pid"][sp["charge"] < 0]).any()
        same_path = (cell == ocell) & (tries == otries)
        assert same_path.mean() >= 1 - 2e-3, "paths differ for %d of %d" % ((~same_path).sum(), len(had))
        ident, close = compare_hadrons(had[same_path], ohad[same_path])
        assert close.all(), "%d hadrons differ beyond 4 ulp" % (~close).sum()
        assert ident.mean() >= 0.80, ident.mean()
        assert tries.mean() > 50
        assert abs(cnt.n_tries - otries.sum()) <= 10 + 20000*(~same_path).sum()
    finally:
        e.close()


@pytest.mark.parametrize("name", lc.STATS_CASES)
def test_legacy_spectra_match_reference_sampler(name, built):
    capi = built
    g = cases.load(name, "legacy_stats")
    par = lc.parameters(g)
    e, sp = make_engine(capi, g, par)
    try:
        e.legacy_compute_yields()
        nev_ref = int(g["nev"])
        nev = nev_ref
        e.sample(4242, 0, nev)
        off = e.event_offsets(nev)
        had = e.fetch_all()
        mine = obs.summarize(had, off)
    finally:
        e.close()
    tot_chi2, tot_ndf, worst = 0.0, 0, (1.0, "")
    for pid in obs.PIDS:
        tag = "p%d" % pid if pid > 0 else "m%d" % (-pid)
        for kind in ("pt", "y", "phi"):
            chi2, ndf = obs.chi2_two_hist(mine[tag + "_" + kind], g[tag + "_" + kind], nev, nev_ref)
            if ndf == 0:
                continue
            p = stats.chi2.sf(chi2, ndf)
            tot_chi2 += chi2
            tot_ndf += ndf
            if p < worst[0]:
                worst = (p, tag + "_" + kind)
    assert worst[0] > 1e-4, worst
    assert tot_ndf > 100
    assert stats.chi2.sf(tot_chi2, tot_ndf) > 0.01, (tot_chi2, tot_ndf)


def test_legacy_per_species_sample_files(built, tmp_path):
    """output_samples_into_files = 1 (emissionfunction.cpp:3353-3375, 3441-3560, 3578-3617): one
    samples_<monval>.dat / samples_control_<monval>.dat pair per species and samples_format.dat; the
    lines are the primaries of the hadron lists, event after event, in the 18-column layout."""
    capi = built
    g = cases.load("surf3d_bulk1", "legacy_stats")
    nev = 25
    s = _facade(capi, g, tmp_path, number_of_repeated_sampling=nev, output_samples_into_files=1)
    try:
        assert s.read_in_FO_surface() == 0
        s.set_random_seed(3)
        assert s.generate_samples() == 0
        h, off = s.hadrons()
    finally:
        s.close()
    fmt = open(tmp_path/"samples_format.dat").read()
    assert "Total_number_of_columns = 18" in fmt and "p_z = 16" in fmt
    n_lines = 0
    for pid in np.unique(h["pid"]):
        ctl = np.loadtxt(tmp_path/("samples_control_%d.dat" % pid), ndmin=1).astype(int)
        rows = np.loadtxt(tmp_path/("samples_%d.dat" % pid), ndmin=2)
        assert len(ctl) == nev and ctl.sum() == len(rows) == (h["pid"] == pid).sum()
        assert rows.shape[1] == 18
        mine = h[h["pid"] == pid]
        per_ev = [(h["pid"][off[e]:off[e + 1]] == pid).sum() for e in range(nev)]
        assert np.array_equal(ctl, per_ev)
        # columns 15..18: E, p_z, t, z; 3, 4: x, y of the cell; 6, 7: pT, phi; 13: y
        for col, key in ((14, "E"), (15, "pz"), (16, "t"), (17, "z"), (2, "x"), (3, "y")):
            assert np.allclose(rows[:, col], mine[key], rtol=2e-6, atol=1e-6), key
        assert np.allclose(rows[:, 5]*np.cos(rows[:, 6]), mine["px"], rtol=1e-5, atol=1e-6)
        assert np.all((rows[:, 6] >= 0) & (rows[:, 6] < 2*np.pi + 1e-6))
        assert np.allclose(rows[:, 12], rows[:, 4] + rows[:, 13], atol=1e-5)     # y = (y - eta_s) + eta_s
        cells = rows[:, 0].astype(int)
        assert np.allclose(rows[:, 1], g["lab"][cells, 0], rtol=2e-6)           # tau of the cell
        n_lines += len(rows)
    assert n_lines == len(h) > 100


# ---- through the drop-in facade (class iSS with MC_sampling = 2) ---------------------------------
def _facade(capi, g, tmp_path, **extra):
    param, surf, over = cases.materialise(g, str(tmp_path))
    over.update(use_OSCAR_format=0, use_gzip_format=0, use_binary_format=0, perform_checks=0)
    over.update(extra)
    return capi.Sampler(str(tmp_path), param, surf, **over)


@pytest.mark.parametrize("name", lc.STATS_CASES)
def test_legacy_facade_matches_reference_sampler(name, built, tmp_path):
    """iSS::read_in_FO_surface / generate_samples / hadron lists with MC_sampling = 2: species
    totals against the reference's yields, spectra against the reference's own samples."""
    capi = built
    g = cases.load(name, "legacy_stats")
    nev_ref = int(g["nev"])
    nev = nev_ref
    s = _facade(capi, g, tmp_path, number_of_repeated_sampling=nev)
    try:
        assert s.read_in_FO_surface() == 0
        s.set_random_seed(777)
        assert s.generate_samples() == 0
        assert s.get_number_of_sampled_events() == nev
        h, off = s.hadrons()
        mine = obs.summarize(h, off)
        # the species the facade sampled, in the reference's order, and their totals
        sp = s.species()
        assert np.array_equal(sp["pid"], g["species"][:, 0].astype(np.int64))
        par = lc.parameters(g)
        opt = lc.oracle_options(par)
        coef = lgo.cell_coefficients(g["lab"], opt, lgo.load_kappa())
        y = lgo.yields(g["lab"], lc.species_array(g["species"]), opt, coef)
        assert np.allclose(s.species_dN(), np.maximum(y, 0).sum(axis=1), rtol=1e-9, atol=1e-300)
    finally:
        s.close()
    tot_chi2, tot_ndf, worst = 0.0, 0, (1.0, "")
    for pid in obs.PIDS:
        tag = "p%d" % pid if pid > 0 else "m%d" % (-pid)
        for kind in ("pt", "y", "phi"):
            chi2, ndf = obs.chi2_two_hist(mine[tag + "_" + kind], g[tag + "_" + kind], nev, nev_ref)
            if ndf == 0:
                continue
            p = stats.chi2.sf(chi2, ndf)
            tot_chi2 += chi2
            tot_ndf += ndf
            if p < worst[0]:
                worst = (p, tag + "_" + kind)
    assert worst[0] > 1e-4, worst
    assert stats.chi2.sf(tot_chi2, tot_ndf) > 0.01, (tot_chi2, tot_ndf)


def test_legacy_facade_decays(built, tmp_path):
    """EmissionFunctionArray::shell runs perform_resonance_feed_down on the sampled events
    (emissionfunction.cpp:2570-2572, 3970-4004): same decay kernel as the FSSW path.  (With the
    SMASH list the reference's pole-mass decayer meets 365 channels below threshold and produces
    NaN momenta -- the reason FSSW::shell skips SMASH, FSSW.cpp:346; the engine reports those as
    ISS_ERR_RANGE instead, so the UrQMD list is used here.)"""
    capi = built
    g = cases.load("l3d_shear", "legacy")
    stable = {211, -211, 111, 321, -321, 2212, -2212, 2112, -2112, 22}
    out = {}
    for decays in (0, 1):
        s = _facade(capi, g, tmp_path/("d%d" % decays), number_of_repeated_sampling=1500,
                    perform_decays=decays)
        try:
            assert s.read_in_FO_surface() == 0
            s.set_random_seed(5)
            assert s.generate_samples() == 0
            h, off = s.hadrons()
            out[decays] = h.copy()
        finally:
            s.close()
    assert len(out[1]) > len(out[0]) > 1000
    frac0 = np.isin(out[0]["pid"], list(stable)).mean()
    frac1 = np.isin(out[1]["pid"], list(stable)).mean()
    assert frac1 > frac0 + 0.15
    rho_omega = [113, 213, -213, 223]
    assert np.isin(out[0]["pid"], rho_omega).any() and not np.isin(out[1]["pid"], rho_omega).any()
    # energy is conserved by the feed-down up to the 4-body channels, which emit nothing
    assert out[1]["E"].sum() <= out[0]["E"].sum()*(1 + 1e-5)
    assert out[1]["E"].sum() > 0.8*out[0]["E"].sum()
