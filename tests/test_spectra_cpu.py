"""Smooth Cooper-Frye spectra: the numpy restatement
(oracle/spectra_oracle.py) against the golden vectors produced by the unmodified reference
(EmissionFunctionArray::calculate_dN_pTdpTdphidy / calculate_flows, tests/golden/make_golden.py
`spectra`).  CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, REPO)
import iss_oracle as orc  # noqa: E402
import spectra_oracle as spo  # noqa: E402
import spectra_cases as sc  # noqa: E402


@pytest.mark.parametrize("name", sc.SPECTRA)
def test_oracle_spectra_match_reference(name):
    g = sc.load(name)
    opt = sc.options_of(g)
    pT, phi, eta = sc.bin_tables()
    kappa = sc.kappa_table() if opt["include_diff"] == 1 else None
    worst = 0.0
    for k, sp in enumerate(sc.species_of(g)):
        dN, dN_max = spo.spectra(g["lab"], sp, opt, pT, phi, eta, kappa)
        ref, ref_max = g["dN"][k], g["dN_max"][k]
        scale = np.abs(ref).max()
        # FP64 parity: same expression order, sequential sum; libm exp vs numpy exp ~1 ulp
        assert np.allclose(dN, ref, rtol=1e-11, atol=1e-13*scale), (name, sp["pid"])
        assert np.allclose(dN_max, ref_max, rtol=1e-11, atol=1e-13*ref_max.max()), (name, sp["pid"])
        worst = max(worst, np.abs(dN - ref).max()/scale)
    assert worst < 1e-11


@pytest.mark.parametrize("name", ["sp3d_shear", "sp3d_bulk1_diff"])
def test_oracle_flows_match_reference(name):
    g = sc.load(name)
    pT, phi, _ = sc.bin_tables()
    for k, sp in enumerate(sc.species_of(g)):
        vd, vi = spo.flows(g["dN"][k], pT, phi, sp["mass"], 9)
        # the reference's files carry 9 significant digits
        assert np.allclose(vd, g["vndiff"][k], rtol=2e-8, atol=1e-12)
        assert np.allclose(vi, g["vninte"][k], rtol=2e-8, atol=1e-12)
    # text format of Table::printTable, byte for byte, from the reference's own numbers
    sp0 = sc.species_of(g)[0]
    vd, _ = spo.flows(g["dN"][0], pT, phi, sp0["mass"], 9)
    want = bytes(g["vndiff_text0"]).decode()
    got = spo.format_table(vd)
    assert len(got) == len(want)
    assert [len(x) for x in got.split("\n")] == [len(x) for x in want.split("\n")]
    same = sum(a == b for a, b in zip(got.split(), want.split()))
    assert same >= 0.99*len(want.split())


def test_bulk_polynomials_kind1_matches_fssw_port():
    T = np.linspace(0.1, 0.18, 17)
    # same coefficients; the FSSW port builds the powers with pow(), the sums cancel to ~1e-7
    assert np.allclose(spo.bulk_coefficients(1, T)[:, :2], orc.coef_poly1(T)[:, :2], rtol=1e-7)
