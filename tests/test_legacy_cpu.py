"""Pins the CPU restatement of the reference's legacy sampler (MC_sampling = 2,
EmissionFunctionArray::sample_using_dN_dxtdy_4all_particles_conventional,
src/emissionfunction.cpp:3273-3623; oracle/legacy_oracle.py + oracle/iss_oracle.c) against the
compiled reference: per-cell yields and estimate_maximum values dumped by oracle/ref_driver.cpp
(tests/golden/legacy_*.npz) and the spectra of the reference's own samples
(tests/golden/legacy_stats_*.npz)."""
import numpy as np
import pytest
from scipy import special, stats

import cases
import legacy_cases as lc
import obs
from legacy_cases import lgo, orc


def test_lambert_w_known_answers():
    lib = orc.clib()
    import ctypes as C
    lib.oracle_lambert_w0.restype = C.c_double
    for x in (0.0, 1e-6, 0.1, 1.0, np.e, 10.0, 150.0):
        w = lib.oracle_lambert_w0(C.c_double(x))
        assert abs(w - special.lambertw(x).real) <= 1e-14*(1 + abs(w))
    assert abs(lib.oracle_lambert_w0(C.c_double(np.e)) - 1.0) < 1e-15


@pytest.mark.parametrize("name", lc.YIELD_CASES)
def test_oracle_yields_and_maxima_match_reference(name):
    g = cases.load(name, "legacy")
    par = lc.parameters(g)
    opt = lc.oracle_options(par)
    sp = lc.species_array(g["species"])
    coef = lgo.cell_coefficients(g["lab"], opt, lgo.load_kappa())
    y = lgo.yields(g["lab"], sp, opt, coef)
    ref = g["yields"]
    assert (ref < 0).any() or name in lc.MORE_CASES   # not clamped (cells with u.dsigma < 0)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(y - ref)/scale).max() < 1e-12
    mx = lgo.estimate_maximum(g["lab"], coef, sp, opt, lgo.load_z_table())
    assert not np.isnan(mx).any()
    assert np.allclose(mx, g["maximum"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", ["cell_shear", "cell_lcc", "cell_bulk3", "surf3d_bulk0"])
def test_oracle_sampler_matches_reference_sampler(name, tmp_path):
    """chi2 of pT / y / phi spectra of the restated sampler against the reference's own samples
    (10^4 events); acceptance as in tests/test_stats_gpu.py."""
    g = cases.load(name, "legacy_stats")
    par = lc.parameters(g)
    opt = lc.oracle_options(par)
    lab, pos, sp = g["lab"], g["pos"], lc.species_array(g["species"])
    coef = lgo.cell_coefficients(lab, opt, lgo.load_kappa())
    y = lgo.yields(lab, sp, opt, coef)
    dN = np.maximum(y, 0).sum(axis=1)
    if opt.boost_invariant:
        dN = dN*(opt.y_RB - opt.y_LB)
    nev, nev_ref = 4000, int(g["nev"])
    mult, outc = orc.multiplicities(dN, orc.poisson_pmode(dN), sp, nev, 0, 99, lcc=opt.lcc)
    h, cell, tries = lgo.sample(lab, pos, coef, y, sp, opt, lgo.load_z_table(), 99, 0, mult,
                                outc.sum() + 8)
    assert len(h) == outc.sum()
    assert tries.mean() > 50            # the legacy sampler needs hundreds of tries per hadron
    off = np.concatenate([[0], np.cumsum(outc.sum(axis=1))])
    mine = obs.summarize(h, off)
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


def test_facade_refuses_what_the_legacy_class_offers_beyond_the_conventional_sampler(built, tmp_path):
    """class iSS: the grid samplers of the legacy class (MC_sampling = 1 / 3) are out of scope and end
    with a message and a non-zero exit (before any device is touched); nothing is approximated or
    silently ignored."""
    import os
    import subprocess
    capi = built
    g = cases.load("cell_shear", "legacy_stats")
    folder = str(tmp_path/"case")
    param, surf, over = cases.materialise(g, folder)
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iSS.e")
    base = [exe, param, "case", surf] + ["%s=%g" % kv for kv in over.items()]
    os.symlink(orc.TABLES, str(tmp_path/"iSS_tables"))
    for extra, text in ((["MC_sampling=1"], "MC_sampling = 1/3"), (["MC_sampling=3"], "MC_sampling = 1/3")):
        r = subprocess.run(base + extra, cwd=str(tmp_path), capture_output=True, text=True,
                           env=dict(os.environ, ISS_INGEST="host"))
        assert r.returncode != 0
        assert text in r.stdout + r.stderr, (extra, r.stdout[-400:])


@pytest.mark.parametrize("kind,name", [("legacy", n) for n in lc.YIELD_CASES]
                         + [("legacy_stats", n) for n in lc.STATS_CASES])
def test_host_keeps_the_reference_lab_frame_cells(kind, name, built, tmp_path):
    """With MC_sampling = 2 iSS::read_in_FO_surface keeps the lab-frame (Milne) cells
    (src/iSS.cpp:105-109): records, positions and the legacy species order of the host code are
    bit-identical to the dump of the compiled reference (binary and text surfaces, 3+1D and
    boost-invariant, with and without grouping_particles)."""
    import os
    import subprocess
    capi = built
    g = cases.load(name, kind)
    param, surf, over = cases.materialise(g, str(tmp_path/"case"))
    os.symlink(orc.TABLES, str(tmp_path/"iSS_tables"))
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iss_host_dump")
    r = subprocess.run([exe, param, "case", surf, "out"] + ["%s=%g" % kv for kv in over.items()],
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=dict(os.environ, ISS_INGEST="host"))
    assert r.returncode == 0, r.stdout.decode()[-1500:]
    with open(tmp_path/"out.lab.bin", "rb") as f:
        n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        lab = np.fromfile(f, dtype=np.float32).reshape(n, 32)
    pos = np.fromfile(tmp_path/"out.pos.bin", dtype=np.float32).reshape(n, 4)
    assert np.array_equal(lab.view(np.uint32), g["lab"].view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), g["pos"].view(np.uint32))
    sp = np.loadtxt(tmp_path/"out.species.txt", ndmin=2)
    assert np.array_equal(sp[:, :7], g["species"][:, :7])
