"""Full-size checks on the benchmark workload (BASELINE.json configs[3]: 10^6-cell 3+1D surface,
EOS 14 + diffusion, 321 species) through size-independent properties, since no oracle finishes
in seconds at this size: internal consistency of the integer bookkeeping, mass-shell and
space-time constraints of every hadron, run-to-run determinism, yields against the numpy oracle on
a random subset of cells, and the command line end to end."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import bench  # noqa: E402
import iss_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(built, tmp_path_factory):
    capi = built
    work = str(tmp_path_factory.mktemp("c4"))
    bench.make_case(work, 1000000)
    s = capi.Sampler(work, bench.PARAM, "surface.dat",
                     **dict(bench.OVERRIDES, number_of_repeated_sampling=100))
    assert s.read_in_FO_surface() == 0
    s.set_random_seed(7)
    assert s.prepare_sampler() == 0
    yield capi, s
    s.close()


def test_yields_subset_against_oracle(big):
    capi, s = big
    e = s.engine()
    dN, y = e.compute_yields(want_cells=True)
    lrf = s.lrf_surface()
    sp = s.species()
    assert y.shape == (321, len(lrf)) and len(lrf) > 900000
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(len(lrf), 300, replace=False))
    tabs = orc.Tables(afterburner="urqmd", kind=21, include_bulk=1, include_diff=1)
    ref = orc.yields(lrf[pick], sp, tabs, 21, 1, 1)
    got = y[:, pick]
    nz = ref != 0
    assert np.array_equal(nz, got != 0)
    assert (np.abs(got[nz] - ref[nz])/np.abs(ref[nz])).max() < 1e-6
    # species totals: sequential FP64 sum (reference order) vs the device's tiled scan
    assert np.allclose(dN, orc.species_totals(y), rtol=1e-11)
    assert np.all(y >= 0)


def test_sampling_properties_and_determinism(big):
    capi, s = big
    e = s.engine()
    e.compute_yields()
    nev = 100
    c = e.sample(99, 1000, 1000 + nev)
    mult = e.multiplicities(nev)
    off = e.event_offsets(nev)
    had = e.fetch_all().copy()
    assert c.n_hadrons == len(had) == mult.sum() == off[-1]
    assert np.array_equal(np.diff(off), mult.sum(axis=1))
    sp = s.species()
    # species-major inside every event, in sampling (mass) order
    pid_of = sp["pid"]
    ev0 = had[off[0]:off[1]]
    assert np.array_equal(ev0["pid"], np.repeat(pid_of, mult[0]))
    # mass shell and kinematics in float32
    E, px, py, pz, m = (had[k].astype(np.float64) for k in ("E", "px", "py", "pz", "mass"))
    assert np.all(E > 0)
    assert np.abs(E*E - px*px - py*py - pz*pz - m*m).max() < 2e-5*np.max(E*E)
    # position: (x, y) of a cell, tau^2 = t^2 - z^2 of the same cell
    lrf = s.lrf_surface()
    key = {(float(a), float(b)): float(t) for a, b, t in zip(lrf[:, 1], lrf[:, 2], lrf[:, 0])}
    sub = had[::997]
    for h in sub:
        tau = key[(float(h["x"]), float(h["y"]))]
        assert abs(np.sqrt(float(h["t"])**2 - float(h["z"])**2) - tau) < 2e-4*max(1.0, float(h["t"]))
    # Poisson means: total multiplicity per event vs sum of yields
    dN = s.species_dN() if len(s.species_dN()) else None
    lam, _ = e.poisson_params()
    tot = mult.sum(axis=1).astype(np.float64)
    assert abs(tot.mean() - lam.sum()) < 5*np.sqrt(lam.sum()/nev)
    # same seed, same events -> same bytes; other events -> different
    e.sample(99, 1000, 1000 + nev)
    assert e.fetch_all().tobytes() == had.tobytes()
    e.sample(99, 5000, 5000 + nev)
    assert e.fetch_all().tobytes() != had.tobytes()
    assert c.n_tries >= c.n_hadrons


def test_command_line_end_to_end(built, tmp_path):
    """iSS.e [param] [path] [surface] key=value ... (reference src/main.cpp:22-74) on a one-cell CI
    fixture: runs, writes OSCAR.DAT and the perform_checks files, closure within the reference's
    CI bound."""
    import cases
    capi = built
    g = cases.load("viscous2")
    case = tmp_path/"case"
    param, surf, over = cases.materialise(g, str(case))
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iSS.e")
    args = [exe, param, "case", surf, "number_of_repeated_sampling=200", "perform_checks=1",
            "use_OSCAR_format=1", "randomSeed=3"] + ["%s=%r" % kv for kv in over.items()]
    r = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-2000:]
    assert "Program executed normally." in out
    data = np.loadtxt(tmp_path/"checkReconstructedTmunu.dat")
    assert np.mean(np.abs(data[:, 2])) < 1e-3
    lines = open(tmp_path/"OSCAR.DAT").read().splitlines()
    assert lines[0].startswith("OSC1997A")
    n0 = int(lines[3].split()[1])
    assert int(lines[3].split()[0]) == 0 and n0 > 10000
    assert len(lines[4].split()) == 11
    assert os.path.exists(tmp_path/"check_211_spectra.dat")


def test_decays_at_scale_conserve_four_momentum(built, tmp_path):
    """BASELINE.json configs[2]-like: boost-invariant 10^5-cell surface, CE delta-f, resonance
    decays (UrQMD table, since the reference skips decays for SMASH, FSSW.cpp:346).  Pole-mass
    decays conserve four-momentum: per event, sum p^mu of the final hadrons equals that of the
    primaries whose decay chain was kept (chains through 4-body / Npart<0 channels drop the
    mother, particle_decay.cpp:291-327), so the totals can only decrease, and by little."""
    from iss_b200 import synthetic
    capi = built
    work = str(tmp_path)
    synthetic.make_case(work, ncell=100000, seed=12345, eos=9, boost_invariant=True)
    over = dict(bench.OVERRIDES, hydro_mode=1, include_deltaf_diffusion=0, perform_decays=1,
                number_of_repeated_sampling=20, y_LB=-2.0, y_RB=2.0)
    s = capi.Sampler(work, bench.PARAM, "surface.dat", **over)
    try:
        s.read_in_FO_surface()
        s.set_random_seed(3)
        s.prepare_sampler()
        e = s.engine()
        e.compute_yields()
        nev = 20
        e.timing(enable=True, reset=True)
        e.sample(3, 0, nev)
        prim = e.fetch_all().copy()
        off = e.event_offsets(nev).copy()
        e.decay(3)
        fin = e.fetch_all()
        foff = e.event_offsets(nev)
        ms, _ = e.timing(enable=False)
        print("primaries %d finals %d  decay kernels %.3f ms  proposal %.3f ms" %
              (len(prim), len(fin), ms["decay"], ms["sample"]))
        assert len(fin) > 1.2*len(prim)
        dsp = {int(p): (int(st)) for p, st in []}
        for ev in range(nev):
            a, b = prim[off[ev]:off[ev + 1]], fin[foff[ev]:foff[ev + 1]]
            Ea, Eb = a["E"].astype(np.float64).sum(), b["E"].astype(np.float64).sum()
            assert Eb <= Ea*(1 + 1e-6)
            assert Eb >= 0.93*Ea
            for k in ("px", "py", "pz"):
                pa, pb = a[k].astype(np.float64).sum(), b[k].astype(np.float64).sum()
                assert abs(pa - pb) < 0.05*Ea
        # rapidity window of the boost-invariant mode for primaries
        mT = np.sqrt(prim["mass"].astype(np.float64)**2 + prim["px"].astype(np.float64)**2
                     + prim["py"].astype(np.float64)**2)
        y = np.arcsinh(prim["pz"]/mT)
        assert y.min() > -2.0 - 1e-4 and y.max() < 2.0 + 1e-4
    finally:
        s.close()
