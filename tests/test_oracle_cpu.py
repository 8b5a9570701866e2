"""Pins the oracle (oracle/iss_oracle.py + oracle/iss_oracle.c) against the unmodified reference:
golden vectors dumped by oracle/ref_driver.cpp / iSS.e in this container (tests/golden/).
CPU only.  The oracle is then the checker of the CUDA engine in the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
from scipy import stats

import cases
import obs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import iss_oracle as orc  # noqa: E402


def species_array(g):
    sp = g["species"]
    a = np.zeros(len(sp), dtype=[("pid", "<i4"), ("gspin", "<i4"), ("baryon", "<i4"),
                                 ("strange", "<i4"), ("charge", "<i4"), ("sign", "<i4"),
                                 ("decay_idx", "<i4"), ("reserved", "<i4"), ("mass", "<f8")])
    a["pid"], a["mass"], a["gspin"] = sp[:, 0], sp[:, 1], sp[:, 2]
    a["baryon"], a["strange"], a["charge"], a["sign"] = sp[:, 3], sp[:, 4], sp[:, 5], sp[:, 6]
    return a


def mode_of(g):
    """(kind, include_bulk, include_diff, include_shear, afterburner) from the parameter file + overrides."""
    par = {}
    for line in open(os.path.join(cases.FIX, str(g["param"]))):
        line = line.split("#")[0]
        if "=" in line:
            k, v = line.split("=")[:2]
            try:
                par[k.strip().lower()] = float(v)
            except ValueError:
                pass
    par.update({k.lower(): v for k, v in cases.overrides_of(g).items()})
    ns = len(g["species"])
    return dict(kind=int(par["bulk_deltaf_kind"]), include_bulk=int(par["include_deltaf_bulk"]),
                include_diff=int(par["include_deltaf_diffusion"]),
                include_shear=int(par["include_deltaf_shear"]),
                afterburner="smash" if ns == 400 else "urqmd", hydro_mode=int(par["hydro_mode"]),
                y_LB=par["y_lb"], y_RB=par["y_rb"])


@pytest.mark.parametrize("name", cases.ONE_CELL + cases.SYNTH)
def test_oracle_yields_match_reference(name):
    g = cases.load(name)
    m = mode_of(g)
    tabs = orc.Tables(table_path=cases.tables_for(g), afterburner=m["afterburner"], kind=m["kind"],
                      include_bulk=m["include_bulk"],
                      include_diff=m["include_diff"])
    y = orc.yields(g["lrf"], species_array(g), tabs, m["kind"], m["include_bulk"], m["include_diff"])
    ref = g["yields"]
    assert np.array_equal(ref == 0.0, y == 0.0)
    nz = ref != 0
    err = np.abs(y[nz] - ref[nz])/np.abs(ref[nz])
    assert err.max() < 1e-9, err.max()
    tot = orc.species_totals(y)
    assert np.allclose(tot, ref.sum(axis=1), rtol=1e-9, atol=0)


def test_known_answers_survey_appendix_d():
    """per-species totals printed by the reference for the CI fixtures."""
    g = cases.load("ideal1")
    tot = dict(zip(g["species"][:, 0].astype(int), g["yields"].sum(axis=1)))
    assert abs(tot[-211] - 2251.2525604315738) < 1e-9
    assert abs(tot[2212] - 137.62717093646677) < 1e-9
    assert abs(g["yields"].sum() - 16569.643299947027) < 1e-7
    g = cases.load("viscous2")
    tot = dict(zip(g["species"][:, 0].astype(int), g["yields"].sum(axis=1)))
    assert abs(tot[-211] - 2293.6021367503986) < 1e-9


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (Salmon et al., kat_vectors)."""
    assert list(orc.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(orc.philox([0xffffffff]*4, [0xffffffff]*2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6,
                                                                  0x6d5451fd]
    assert list(orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                           [0xa4093822, 0x299f31d0])) == [0xd16cfe09, 0x94fdcceb, 0x5001e420,
                                                          0x24126ea1]
    u = orc.stream_uniforms(12345, 2, 7, 3, 11, 1000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.05


def test_poisson_inversion_is_poisson():
    lam = np.array([0.0, 1e-16, 0.3, 4.5, 167.2, 2251.25, 23000.7])
    pm = orc.poisson_pmode(lam)
    sp = np.zeros(len(lam), dtype=orc.OSpecies)
    mult, outc = orc.multiplicities(lam, pm, sp, 20000, 0, 99)
    assert np.array_equal(mult, outc)
    assert np.all(mult[:, :2] == 0)
    for j in range(2, len(lam)):
        n = mult[:, j].astype(float)
        # mean and variance of a Poisson within 5 sigma
        assert abs(n.mean() - lam[j]) < 5*np.sqrt(lam[j]/len(n))
        assert abs(n.var() - lam[j]) < 5*lam[j]*np.sqrt(2.0/len(n)) + 5*np.sqrt(lam[j]/len(n))
    # small-mean pmf by chi2
    k = np.arange(0, 20)
    cnt = np.bincount(mult[:, 3], minlength=20)[:20]
    exp = stats.poisson.pmf(k, 4.5)*len(mult)
    msk = exp > 10
    chi2 = ((cnt[msk] - exp[msk])**2/exp[msk]).sum()
    assert stats.chi2.sf(chi2, msk.sum() - 1) > 1e-3
    # model 1: floor + Bernoulli
    m1, _ = orc.multiplicities(lam, pm, sp, 20000, 0, 5, model=1)
    assert set(np.unique(m1[:, 3])) <= {4, 5}
    assert abs(m1[:, 3].mean() - 4.5) < 0.02


def test_negative_binomial_models():
    """models 10 and 20 (FSSW.cpp:275-292): X ~ NB(n = k, p = 1/(1+para1)); mean k para1,
    variance k para1 (1 + para1); model 10 adds the draw to the integer part."""
    para1 = 0.16
    lam = np.array([0.37, 4.5, 167.2, 2251.25])
    sp = np.zeros(len(lam), dtype=orc.OSpecies)
    pm = orc.poisson_pmode(lam)
    nev = 40000
    m20, _ = orc.multiplicities(lam, pm, sp, nev, 0, 7, model=20, para1=para1)
    for j, dN in enumerate(lam):
        k = para1*dN
        mean, var = k*para1, k*para1*(1 + para1)
        x = m20[:, j].astype(float)
        assert abs(x.mean() - mean) < 5*np.sqrt(var/nev) + 1e-12
        assert abs(x.var() - var) < 0.06*var + 5*var*np.sqrt(2.0/nev)
    # pmf of one case against scipy
    k = para1*lam[2]
    cnt = np.bincount(m20[:, 2], minlength=30)[:30]
    exp = stats.nbinom.pmf(np.arange(30), k, 1.0/(1 + para1))*nev
    msk = exp > 10
    chi2 = ((cnt[msk] - exp[msk])**2/exp[msk]).sum()
    assert stats.chi2.sf(chi2, msk.sum() - 1) > 1e-3
    m10, _ = orc.multiplicities(lam, pm, sp, nev, 0, 8, model=10, para1=para1)
    for j, dN in enumerate(lam):
        frac = dN - np.floor(dN)
        x = m10[:, j].astype(float) - np.floor(dN)
        assert x.min() >= 0
        assert abs(x.mean() - para1*frac*para1) < 5*np.sqrt(para1*frac*para1*(1 + para1)/nev) + 1e-12


def test_momentum_sampler_matches_reference():
    """|p| spectra of MomentumSamplerShell::Sample_a_momentum (2e6 samples per case from the
    compiled reference) against the C restatement; two-sample chi2, p > 1e-3 per case (ten cases)
    and p > 0.01 combined."""
    g = np.load(os.path.join(cases.GOLDEN, "momentum_sampler.npz"))
    n = 1000000
    tot_chi2, tot_ndf = 0.0, 0
    for i, (m, T, mu, sign) in enumerate(g["cases"]):
        p = orc.sample_momentum(m, T, mu, int(sign), n, 1000 + i)
        h = np.histogram(p, g["edges"])[0]
        chi2, ndf = obs.chi2_two_hist(h, g["hist"][i], n, int(g["n"]))
        assert stats.chi2.sf(chi2, ndf) > 1e-3, (i, chi2, ndf)
        tot_chi2 += chi2
        tot_ndf += ndf
    assert stats.chi2.sf(tot_chi2, tot_ndf) > 0.01


def test_decay_matches_reference():
    """particle_decay::perform_decays on fixed mothers (reference dump) vs the C restatement:
    daughter-number distribution, channel frequencies, mean four-momentum, energy spectrum."""
    g = np.load(os.path.join(cases.GOLDEN, "decay.npz"))
    ds, dc = orc.read_pdg_table(os.path.join(orc.TABLES, "pdg-urqmd_v3.3+.dat"))
    n = 200000
    nref = int(g["n"])
    for pid in g["pids"]:
        mass = np.float32(ds["mass"][ds["pid"] == pid][0])
        mom = (int(pid), mass, np.sqrt(np.float32(mass*mass + 0.3**2 + 0.2**2 + 0.5**2)), 0.3, -0.2, 0.5,
               1.0, 0.5, -0.5, 0.25)
        out, nd = orc.decay_once_many(int(pid), n, 77, ds, dc, mom)
        ndh = np.bincount(nd, minlength=8)[:8]
        chi2, ndf = obs.chi2_two_hist(ndh, g["nd_%d" % pid], n, nref)
        if ndf > 1:
            assert stats.chi2.sf(chi2, ndf - 1) > 1e-4, (pid, ndh, g["nd_%d" % pid])
        else:
            assert np.array_equal(ndh > 0, g["nd_%d" % pid] > 0)
        # four-momentum conservation per decay and on average
        p4 = np.array([out["E"].sum(), out["px"].sum(), out["py"].sum(), out["pz"].sum()])/n
        ref = g["p4sum_%d" % pid]/nref
        assert np.allclose(p4, ref, rtol=2e-2, atol=2e-2), (pid, p4, ref)
        eh = np.bincount(np.minimum(49, (out["E"]/0.05).astype(int)), minlength=50)[:50]
        chi2, ndf = obs.chi2_two_hist(eh, g["ehist_%d" % pid], n, nref)
        assert stats.chi2.sf(chi2, ndf) > 1e-4, (pid, chi2, ndf)
