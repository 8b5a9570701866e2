"""oracle/spectra_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy restatement of the smooth Cooper-Frye spectra of the reference's legacy
EmissionFunctionArray.  Pinned against the unmodified reference
by tests/golden/spectra_*.npz (generated with oracle/ref_driver `spectra`).

  bulk_coefficients()   EmissionFunctionArray::getbulkvisCoefficients     src/emissionfunction.cpp:3625-3762
  deltaf_bulk()         EmissionFunctionArray::get_deltaf_bulk            src/emissionfunction.cpp:4156-4186
  spectra()             EmissionFunctionArray::calculate_dN_pTdpTdphidy   src/emissionfunction.cpp:624-829
  flows()               EmissionFunctionArray::calculate_flows            src/emissionfunction.cpp:875-1016
  dN_dphi()             ...::calculate_dN_dphi_using_dN_pTdpTdphidy       src/emissionfunction.cpp:2428-2455
  format_table()        Table::printTable                                 src/Table.cpp:166-176
  particles_are_the_same()                                                src/emissionfunction.cpp:2511-2537

Lab-frame cell record (32 float32, the ISS_L_* order of include/iss_cuda.h):
  tau u0 u1 u2 u3 da0 da1 da2 da3 T P e muB muS muQ pi00 pi01 pi02 pi03 pi11 pi12 pi13 pi22 pi23
  pi33 bulkPi Bn q0 q1 q2 q3 spare
"""
import os

import numpy as np

import iss_oracle as O

HBARC = O.HBARC

L = {n: i for i, n in enumerate(
    "tau u0 u1 u2 u3 da0 da1 da2 da3 T P e muB muS muQ pi00 pi01 pi02 pi03 pi11 pi12 pi13 pi22 "
    "pi23 pi33 bulkPi Bn q0 q1 q2 q3 spare".split())}

# degree-10 polynomials in T [1/fm] of bulk_deltaf_kind 1..4: [kind][coefficient][power]
BULK_POLY = {
    1: [[642096.624265727, -8163329.49562861, 47162768.4292073, -162590040.002683, 369637951.096896,
         -578181331.809836, 629434830.225675, -470493661.096657, 230936465.421, -67175218.4629078,
         8789472.32652964],
        [1.18171174036192, -17.6740645873717, 136.298469057177, -635.999435106846, 1918.77100633321,
         -3836.32258307711, 5136.35746882372, -4566.22991441914, 2593.45375240886, -853.908199724349,
         124.260460450113]],
    2: [[21091365.1182649, -290482229.281782, 1800423055.01882, -6608608560.99887, 15900800422.7138,
         -26194517161.8205, 29912485360.2916, -23375101221.2855, 11960898238.0134, -3618358144.18576,
         491369134.205902],
        [4007863.29316896, -55199395.3534188, 342115196.396492, -1255681487.77798, 3021026280.08401,
         -4976331606.85766, 5682163732.74188, -4439937810.57449, 2271692965.05568, -687164038.128814,
         93308348.3137008]],
    3: [[160421664.93603, -2212807124.97991, 13707913981.1425, -50204536518.1767, 120354649094.362,
         -197298426823.223, 223953760788.288, -173790947240.829, 88231322888.0423, -26461154892.6963,
         3559805050.19592],
        [33369186.2536556, -460293490.420478, 2851449676.09981, -10443297927.601, 25035517099.7809,
         -41040777943.4963, 46585225878.8723, -36150531001.3718, 18353035766.9323, -5504165325.05431,
         740468257.784873]],
    4: [[1167272041.90731, -16378866444.6842, 103037615761.617, -382670727905.111, 929111866739.436,
         -1540948583116.54, 1767975890298.1, -1385606389545.0, 709922576963.213, -214726945096.326,
         29116298091.9219],
        [5103633637.7213, -71612903872.8163, 450509014334.964, -1673143669281.46, 4062340452589.89,
         -6737468792456.4, 7730102407679.65, -6058276038129.83, 3103990764357.81, -938850005883.612,
         127305171097.249]],
}


def bulk_coefficients(kind, T):
    """[ncell][3].  Kinds 1-4: polynomials, powers by repeated multiplication and the terms added
    left to right as the reference does (the sums cancel to ~1e-7 of the largest term, so the
    order matters).  Kind 0 is never evaluated by calculate_dN_pTdpTdphidy
    (emissionfunction.cpp:728-735 skips the call), every other kind leaves the zeros."""
    T = np.asarray(T, dtype=np.float64)
    out = np.zeros(T.shape + (3,))
    if kind in BULK_POLY:
        x = T/HBARC
        p = [None, x]
        for k in range(2, 11):
            p.append(p[k - 1]*x)
        for c in range(2):
            a = BULK_POLY[kind][c]
            acc = np.full(T.shape, a[0])
            for k in range(1, 11):
                acc = acc + a[k]*p[k]
            out[..., c] = acc
    return out


def deltaf_bulk(kind, include_bulk, mass, pdotu, bulkPi, T, sign, f0, c):
    if include_bulk == 0:
        return np.zeros_like(pdotu)
    if kind == 0:
        return -(1. - sign*f0)*bulkPi*(c[0]*mass*mass + c[1]*pdotu + c[2]*pdotu*pdotu)
    if kind == 1:
        EoT = pdotu/T
        moT = mass/T
        return -1.0*(1. - sign*f0)*c[0]*(moT*moT/(3.*EoT) - c[1]*EoT)*bulkPi
    if kind == 2:
        EoT = pdotu/T
        return -1.*(1. - sign*f0)*bulkPi*(-c[0] + c[1]*EoT)
    if kind == 3:
        EoT = pdotu/T
        return -1.*(1. - sign*f0)*bulkPi/np.sqrt(EoT)*(-c[0] + c[1]*EoT)
    if kind == 4:
        EoT = pdotu/T
        return -1.*(1. - sign*f0)*bulkPi*(c[0] - c[1]/EoT)
    return np.zeros_like(pdotu)


def load_bin_tables(table_path):
    d = os.path.join(table_path, "bin_tables")
    return (np.loadtxt(os.path.join(d, "pT_gauss_table.dat")),
            np.loadtxt(os.path.join(d, "phi_gauss_table.dat")),
            np.loadtxt(os.path.join(d, "eta_uni_table.dat")))


def spectra(lab, species, opt, pT_tab, phi_tab, eta_tab, kappa_tb=None):
    """dN/(pT dpT dphi dy) and its per-(cell, eta) maximum for one species.

    lab: [ncell][32] float32; species: dict(mass, sign, gspin, baryon, strange, charge);
    opt: dict(include_shear, include_bulk, bulk_kind, include_diff, restrict_deltaf,
    deltaf_max_ratio, use_pos_dN_only).  Returns (dN[npT][nphi], dN_max[npT][nphi]).
    The sum runs over cells (outer) and the y - eta_s table (inner) in the reference's order."""
    c = {n: lab[:, i].astype(np.float64)[:, None] for n, i in L.items()}
    mass, sign, degen = species["mass"], species["sign"], species["gspin"]
    B, S, Q = species["baryon"], species["strange"], species["charge"]
    prefactor = 1.0/(8.0*(np.pi*np.pi*np.pi))/HBARC/HBARC/HBARC
    T = c["T"]
    # int * float products and their sum are FLOAT arithmetic in the reference
    # (`double mu = baryon*surf->muB + strange*surf->muS + charge*surf->muQ`, emissionfunction.cpp:700)
    f32 = np.float32
    mu = ((f32(B)*lab[:, L["muB"]] + f32(S)*lab[:, L["muS"]]) + f32(Q)*lab[:, L["muQ"]])
    mu = mu.astype(np.float64)[:, None]
    shear_pref = 1.0/(2.0*T*T*(c["e"] + c["P"])) if opt["include_shear"] else 0.0
    bulkPi = np.zeros_like(T)
    coef = np.zeros((3,) + T.shape)
    if opt["include_bulk"] == 1:
        if opt["bulk_kind"] == 0:
            bulkPi = c["bulkPi"]
        else:
            bulkPi = c["bulkPi"]/HBARC
            bc = bulk_coefficients(opt["bulk_kind"], T[:, 0])
            coef = np.stack([bc[:, k][:, None] for k in range(3)])
    if opt["include_diff"] == 1:
        kappa = O.coef_kappa(kappa_tb, T[:, 0], c["muB"][:, 0])[:, None]
        pref_q = c["Bn"]/(c["e"] + c["P"])
    y_me = eta_tab[:, 0]
    d_eta = eta_tab[:, 1][None, :]
    ch = np.cosh(-y_me)[None, :]
    sh = np.sinh(-y_me)[None, :]
    npT, nphi = len(pT_tab), len(phi_tab)
    dN = np.zeros((npT, nphi))
    dN_max = np.zeros((npT, nphi))
    for i in range(npT):
        pT = pT_tab[i, 0]
        mT = np.sqrt(mass*mass + pT*pT)
        pt = mT*ch
        pz = mT*sh
        for j in range(nphi):
            px = pT*np.cos(phi_tab[j, 0])
            py = pT*np.sin(phi_tab[j, 0])
            pdotu = pt*c["u0"] - px*c["u1"] - py*c["u2"] - pz*c["u3"]
            expon = (pdotu - mu)/T
            with np.errstate(over="ignore"):
                f0 = 1./(np.exp(expon) + sign)
            pdsigma = pt*c["da0"] + px*c["da1"] + py*c["da2"] + pz*c["da3"]/c["tau"]
            df_shear = 0.0
            if opt["include_shear"]:
                W = (pt*pt*c["pi00"] - 2.0*pt*px*c["pi01"] - 2.0*pt*py*c["pi02"]
                     - 2.0*pt*pz*c["pi03"]
                     + px*px*c["pi11"] + 2.0*px*py*c["pi12"] + 2.0*px*pz*c["pi13"]
                     + py*py*c["pi22"] + 2.0*py*pz*c["pi23"]
                     + pz*pz*c["pi33"])
                df_shear = (1 - sign*f0)*W*shear_pref
            df_bulk = deltaf_bulk(opt["bulk_kind"], opt["include_bulk"], mass, pdotu, bulkPi, T, sign,
                                  f0, coef)
            df_q = 0.0
            if opt["include_diff"] == 1:
                qf = pt*c["q0"] - px*c["q1"] - py*c["q2"] - pz*c["q3"]
                df_q = (1. - sign*f0)*(pref_q - B/pdotu)*qf/kappa
            resize = 1.0
            if opt["restrict_deltaf"] == 1:
                size = np.abs(df_shear + df_bulk + df_q)
                resize = np.minimum(1., opt["deltaf_max_ratio"]/(size + 1e-10))
            result = (prefactor*degen*f0*pdsigma*c["tau"]
                      *(1. + (df_shear + df_bulk + df_q)*resize))
            if opt["use_pos_dN_only"]:
                keep = ~(result < 0.)
                result = np.where(keep, result, 0.)
            terms = (result*d_eta).ravel()
            dN[i, j] = np.cumsum(terms)[-1]             # sequential, cells outer / eta inner
            dN_max[i, j] = max(0.0, result.max())
    return dN, dN_max


def flows(dN, pT_tab, phi_tab, mass, to_order):
    """(vn_diff [npT][3 + 3 n], vn_inte [to_order + 1][6]) as written by calculate_flows."""
    npT, nphi = dN.shape
    n = to_order
    vn_diff = np.zeros((npT, 3 + 3*n))
    norm = np.zeros(npT)
    vn = np.zeros((npT, n, 2))
    for i in range(npT):
        pT = pT_tab[i, 0]
        mT = np.sqrt(mass*mass + pT*pT)
        for j in range(nphi):
            phi, w = phi_tab[j, 0], phi_tab[j, 1]
            norm[i] += dN[i, j]*w
            for order in range(1, n + 1):
                vn[i, order - 1, 0] += dN[i, j]*w*np.cos(order*phi)
                vn[i, order - 1, 1] += dN[i, j]*w*np.sin(order*phi)
        norm[i] = norm[i] + 1e-30
        vn_diff[i, 0] = pT
        vn_diff[i, 1] = mT - mass
        vn_diff[i, 2] = norm[i]/(2.0*np.pi)
        for t in range(n):
            vn_diff[i, 3 + 3*t] = vn[i, t, 0]/norm[i]
            vn_diff[i, 4 + 3*t] = vn[i, t, 1]/norm[i]
            vn_diff[i, 5 + 3*t] = np.sqrt(vn[i, t, 0]**2 + vn[i, t, 1]**2)/norm[i]
    normi = 0.0
    vni = np.zeros((n, 2))
    for i in range(npT):
        pT, w = pT_tab[i, 0], pT_tab[i, 1]
        normi += norm[i]*pT*w
        for t in range(n):
            vni[t, 0] += vn[i, t, 0]*pT*w
            vni[t, 1] += vn[i, t, 1]*pT*w
    vn_inte = np.zeros((n + 1, 6))
    vn_inte[0] = [0, normi, 0, 1, 0, 1]
    for t in range(n):
        vn_inte[t + 1] = [1 + t, vni[t, 0], vni[t, 1], vni[t, 0]/normi, vni[t, 1]/normi,
                          np.sqrt(vni[t, 0]**2 + vni[t, 1]**2)/normi]
    return vn_diff, vn_inte


def dN_dphi(dN, pT_tab):
    return (dN*(pT_tab[:, 0]*pT_tab[:, 1])[:, None]).sum(axis=0)


def format_table(tab):
    """Table::printTable: every value as `scientific << setw(15) << setprecision(8)` + two blanks."""
    return "".join("".join("%15.8e  " % v for v in row) + "\n" for row in np.atleast_2d(tab))


def particles_are_the_same(p1, p2, tolerance):
    for k in ("sign", "gspin", "baryon", "strange", "charge"):
        if p1[k] != p2[k]:
            return False
    return not abs((p1["mass"] - p2["mass"])/(p2["mass"] + 1e-30)) > tolerance
