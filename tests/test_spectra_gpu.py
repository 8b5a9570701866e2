"""Smooth Cooper-Frye spectra on the GPU (iss_cuda_spectra, through the C ABI) against the golden
vectors of the unmodified reference and against the numpy restatement.  SURVEY.md section 8 (f)-3."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, REPO)
import spectra_oracle as spo  # noqa: E402
import spectra_cases as sc  # noqa: E402
from iss_b200 import capi  # noqa: E402

pytestmark = pytest.mark.gpu

# FP64 tolerance of this path: the kernel regroups the reference's sums (p.u = mT A - B ...),
# multiplies by 1/T instead of dividing and sums cells in chunks: 1e-10 relative to the largest
# entry of a species' table, 1e-9 element-wise.
RTOL_TABLE, RTOL_ELEM = 1e-10, 1e-9


def species_array(sp_list):
    a = np.zeros(len(sp_list), dtype=capi.SPECIES_DTYPE)
    for k, sp in enumerate(sp_list):
        a[k]["pid"], a[k]["mass"], a[k]["gspin"] = sp["pid"], sp["mass"], int(sp["gspin"])
        a[k]["baryon"], a[k]["strange"], a[k]["charge"] = sp["baryon"], sp["strange"], sp["charge"]
        a[k]["sign"] = int(sp["sign"])
    return a


def engine_for(opt):
    e = capi.Engine(0)
    if opt["include_diff"] == 1:
        e.upload_table(capi.TABLE_KAPPA_B, sc.kappa_table(), 150, 100, [0.05, 0.001, 0.0, 0.007892])
    return e


def run(e, g, opt, species):
    pT, phi, eta = sc.bin_tables()
    e.upload_surface_lab(g["lab"])
    return e.spectra(species_array(species), pT[:, 0], phi[:, 0], eta[:, 0], eta[:, 1],
                     include_deltaf_shear=opt["include_shear"], include_deltaf_bulk=opt["include_bulk"],
                     bulk_deltaf_kind=opt["bulk_kind"], include_deltaf_diffusion=opt["include_diff"],
                     restrict_deltaf=opt["restrict_deltaf"], use_pos_dN_only=opt["use_pos_dN_only"],
                     deltaf_max_ratio=opt["deltaf_max_ratio"])


@pytest.mark.parametrize("name", sc.SPECTRA)
def test_spectra_match_reference(name):
    g = sc.load(name)
    opt = sc.options_of(g)
    species = sc.species_of(g)
    e = engine_for(opt)
    try:
        dN, dN_max = run(e, g, opt, species)
    finally:
        e.close()
    for k, sp in enumerate(species):
        ref, ref_max = g["dN"][k], g["dN_max"][k]
        scale = np.abs(ref).max()
        assert np.abs(dN[k] - ref).max() <= RTOL_TABLE*scale, (name, sp["pid"])
        assert np.allclose(dN[k], ref, rtol=RTOL_ELEM, atol=RTOL_TABLE*scale), (name, sp["pid"])
        assert np.allclose(dN_max[k], ref_max, rtol=RTOL_ELEM, atol=RTOL_TABLE*ref_max.max())


def test_spectra_chunked_surface_matches_oracle_and_is_deterministic():
    """more cells than one chunk (2048) and a ragged last cell tile; repeated calls and a
    different species batching give bit-identical tables"""
    g = sc.load("sp3d_bulk1_diff")
    opt = sc.options_of(g)
    species = sc.species_of(g)[:3]
    rng = np.random.default_rng(5)
    lab = g["lab"][rng.integers(0, len(g["lab"]), size=2048*2 + 37)]
    pT, phi, eta = sc.bin_tables()
    e = engine_for(opt)
    try:
        g2 = {"lab": lab}
        a, amax = run(e, g2, opt, species)
        b, bmax = run(e, g2, opt, species)
        c, _ = run(e, g2, opt, species[1:2])
    finally:
        e.close()
    assert np.array_equal(a, b) and np.array_equal(amax, bmax)
    assert np.array_equal(a[1], c[0])
    want, want_max = spo.spectra(lab, species[1], opt, pT, phi, eta, sc.kappa_table())
    assert np.abs(a[1] - want).max() <= RTOL_TABLE*np.abs(want).max()
    assert np.allclose(amax[1], want_max, rtol=RTOL_ELEM)


def test_spectra_errors():
    e = capi.Engine(0)
    try:
        pT, phi, eta = sc.bin_tables()
        sp = species_array(sc.species_of(sc.load("sp3d_shear"))[:1])
        with pytest.raises(capi.IssError):      # no lab surface uploaded
            e.spectra(sp, pT[:, 0], phi[:, 0], eta[:, 0], eta[:, 1])
        e.upload_surface_lab(sc.load("sp3d_shear")["lab"])
        with pytest.raises(capi.IssError):      # diffusion without the kappa table
            e.spectra(sp, pT[:, 0], phi[:, 0], eta[:, 0], eta[:, 1], include_deltaf_diffusion=1)
        with pytest.raises(capi.IssError):      # y - eta_s table too long
            e.spectra(sp, pT[:, 0], phi[:, 0], np.zeros(200), np.zeros(200))
    finally:
        e.close()


# ---------------------------------------------------------------------------------------------
# whole program: class iSS with MC_sampling = 0, calculate_vn = 1 against the files written by the
# reference's iSS.e on the same surface (tests/golden/flows_*.npz)
def _tables_with_chosen(tmp, chosen):
    d = os.path.join(tmp, "tables")
    os.makedirs(d)
    for f in os.listdir(sc.TABLES):
        if f != "chosen_particles_SMASH.dat":
            os.symlink(os.path.join(sc.TABLES, f), os.path.join(d, f))
    with open(os.path.join(d, "chosen_particles_SMASH.dat"), "w") as f:
        f.write("".join("%d\n" % m for m in chosen))
    return d


def _compare_text(name, got, want):
    gl, wl = got.split("\n"), want.split("\n")
    assert len(gl) == len(wl), name
    assert [len(x) for x in gl] == [len(x) for x in wl], name       # same layout, column for column
    gt, wt = got.split(), want.split()
    same = 0
    for a, b in zip(gt, wt):
        if a == b:
            same += 1
            continue
        if a.startswith("#") or b.startswith("#"):
            assert a == b, name
        try:
            fa, fb = float(a), float(b)
        except ValueError:
            assert a == b, (name, a, b)       # names in the comment lines of the historic format
            continue
        # 9 significant digits in the files; flows are ratios of sums that cancel
        assert abs(fa - fb) <= 3e-8*max(abs(fa), abs(fb)) + 1e-13, (name, a, b)
    # v_n are ratios of cancelling sums: a 1e-13 difference in dN can flip the ninth printed digit
    assert same >= 0.90*len(wt), (name, same, len(wt))


@pytest.mark.parametrize("which", ["new", "old"])
def test_facade_spectra_and_flows_match_reference_files(which, tmp_path):
    import cases
    g = np.load(os.path.join(sc.GOLDEN, "flows_%s.npz" % which), allow_pickle=False)
    folder = str(tmp_path/"case")
    param, surf, over = cases.materialise(g, folder)
    tables = _tables_with_chosen(str(tmp_path), [int(m) for m in g["chosen"]])
    s = capi.Sampler(folder, param, surf, table_path=tables, **over)
    try:
        assert s.read_in_FO_surface() == 0
        assert s.generate_samples() == 0
        tab, ms, ev = s.spectra_table(211)
        assert tab.shape == (15, 48) and ev > 0 and ms > 0
    finally:
        s.close()
    names = [str(n) for n in g["names"]]
    for i, n in enumerate(names):
        want = bytes(g["file_%d" % i]).decode()
        got = open(os.path.join(folder, n)).read()
        _compare_text(n, got, want)
    produced = sorted(f for f in os.listdir(folder) if f.startswith(("thermal_", "dN_", "v2data")))
    assert produced == sorted(names)


def test_facade_rejects_legacy_samplers(tmp_path):
    """MC_sampling = 1/3 (legacy EmissionFunctionArray grid samplers) exit with an error, they do
    not fall back to anything (MC_sampling = 2 is implemented: tests/test_legacy_gpu.py)"""
    import subprocess
    import cases
    g = np.load(os.path.join(sc.GOLDEN, "flows_new.npz"), allow_pickle=False)
    folder = str(tmp_path/"case")
    param, surf, over = cases.materialise(g, folder)
    code = ("import sys; sys.path.insert(0, %r); from iss_b200 import capi; "
            "s = capi.Sampler(%r, %r, %r, MC_sampling=3); s.read_in_FO_surface()" %
            (REPO, folder, param, surf))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "out of scope" in r.stdout
