"""oracle/legacy_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of the reference's legacy "conventional" sampler, MC_sampling = 2
(EmissionFunctionArray, src/emissionfunction.cpp): the per-cell x per-species yields
(calculate_dN_dxtdy_for_one_particle_species :2977-3097, calculate_dN_analytic :3101-3208) in
numpy, and ctypes access to the C restatement (oracle/iss_oracle.c) of estimate_maximum
(:4006-4153, 4309-4421) and of the sampling loops (:3330-3560, 4188-4306, 4423-4475).

Parity pinning: tests/test_legacy_cpu.py checks the yields and the maxima against dumps of the
compiled reference (oracle/ref_driver.cpp `legacy`, fixtures tests/golden/legacy_*.npz) and the
sampled spectra against the reference's own particle_samples.bin (tests/golden/legacy_stats_*.npz).
"""
import ctypes as C
import os

import numpy as np
from scipy import special

import iss_oracle as O
import spectra_oracle as SO

HBARC = O.HBARC
L = {n: i for i, n in enumerate(
    "tau u0 u1 u2 u3 da0 da1 da2 da3 T P e muB muS muQ pi00 pi01 pi02 pi03 pi11 pi12 pi13 pi22 "
    "pi23 pi33 bulkPi nB q0 q1 q2 q3 spare".split())}


class LegacyOpt(C.Structure):
    _fields_ = [("include_shear", C.c_int32), ("include_bulk", C.c_int32),
                ("bulk_kind", C.c_int32), ("include_diff", C.c_int32),
                ("restrict_deltaf", C.c_int32), ("boost_invariant", C.c_int32),
                ("lcc", C.c_int32), ("reserved1", C.c_int32),
                ("deltaf_max_ratio", C.c_double), ("pT_to", C.c_double),
                ("y_minus_eta_s_range", C.c_double), ("y_LB", C.c_double), ("y_RB", C.c_double)]


def make_opt(include_shear=0, include_bulk=0, bulk_kind=1, include_diff=0, restrict_deltaf=0,
             deltaf_max_ratio=1.0, boost_invariant=0, pT_to=4.0, y_range=4.0, y_LB=-5.0, y_RB=5.0,
             lcc=0):
    o = LegacyOpt()
    o.include_shear, o.include_bulk, o.bulk_kind, o.include_diff = (include_shear, include_bulk,
                                                                    bulk_kind, include_diff)
    o.restrict_deltaf, o.boost_invariant = restrict_deltaf, boost_invariant
    o.deltaf_max_ratio, o.pT_to, o.y_minus_eta_s_range = deltaf_max_ratio, pT_to, y_range
    o.y_LB, o.y_RB = y_LB, y_RB
    o.lcc = lcc
    return o


def load_z_table(table_path=O.TABLES):
    v = np.loadtxt(os.path.join(table_path, "z_exp_m_z.dat"))
    return np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1])


def load_kappa(table_path=O.TABLES):
    v = np.loadtxt(os.path.join(table_path, "deltaf_tables", "Coefficients_RTA_diffusion.dat"))
    return v[:150*100, 2].reshape(100, 150).T.copy()   # [T][mu]


def cubic_direct(x, y, xx):
    """interpCubicDirect without extrapolation (arsenal.cpp:58-110) on an equally spaced table,
    vectorised over xx: the edge intervals use the parabola through three points, the others the
    cubic through four."""
    x, y, xx = np.asarray(x, float), np.asarray(y, float), np.asarray(xx, float)
    n = len(x)
    x0, dx = x[0], x[1] - x[0]
    idx = np.floor((xx - x0)/dx).astype(np.int64)
    if np.any((idx < 0) | (idx >= n - 1)):
        raise ValueError("interpCubicDirect: x out of bounds (the reference exits)")
    out = np.empty_like(xx)
    lo, hi = idx == 0, idx == n - 2
    for sel, base in ((lo, 0), (hi, n - 3)):
        if sel.any():
            A0, A1, A2 = y[base], y[base + 1], y[base + 2]
            d = xx[sel] - (x0 + base*dx)
            out[sel] = ((A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d - (3.0*A0 - 4.0*A1 + A2)/(2.0*dx)*d + A0)
    mid = ~(lo | hi)
    if mid.any():
        i = idx[mid]
        A0, A1, A2, A3 = y[i - 1], y[i], y[i + 1], y[i + 2]
        d = xx[mid] - (x0 + i*dx)
        out[mid] = ((-A0 + 3.0*A1 - 3.0*A2 + A3)/(6.0*dx*dx*dx)*d*d*d + (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d
                    - (2.0*A0 + 3.0*A1 - 6.0*A2 + A3)/(6.0*dx)*d + A1)
    return out


def load_bulk14(table_path=O.TABLES):
    """rows T[1/fm], B0, D0, E0 (emissionfunction.cpp:298-301)"""
    return np.loadtxt(os.path.join(table_path, "deltaf_tables",
                                   "BulkDf_Coefficients_Hadrons_s95p-v0-PCE.dat"))


def cell_coefficients(lab, opt, kappa_tb=None, bulk14=None):
    """[ncell][4]: bulkvisCoefficients c0..c2 (getbulkvisCoefficients(T), :3625-3762: polynomials
    for kinds 1-4, cubic interpolation of the s95p-PCE table for kind 0) and kappa_hat (1 when
    diffusion is off)."""
    lab = np.asarray(lab, dtype=np.float32)
    T = lab[:, L["T"]].astype(np.float64)
    out = np.zeros((len(lab), 4))
    out[:, 3] = 1.0
    if opt.include_bulk == 1 and opt.bulk_kind == 0:
        tb = load_bulk14() if bulk14 is None else bulk14
        T_fm = T/HBARC
        out[:, 0] = cubic_direct(tb[:, 0], tb[:, 1], T_fm)/HBARC**3
        out[:, 1] = cubic_direct(tb[:, 0], tb[:, 2], T_fm)/HBARC**2
        out[:, 2] = cubic_direct(tb[:, 0], tb[:, 3], T_fm)/HBARC**3
    elif opt.include_bulk == 1:
        out[:, 0:3] = SO.bulk_coefficients(opt.bulk_kind, T)
    if opt.include_diff == 1:
        out[:, 3] = O.coef_kappa(kappa_tb, T, lab[:, L["muB"]].astype(np.float64))
    return out


def yields(lab, species, opt, coef):
    """dN_dxtdy_for_one_particle_species for every species: [ns][ncell] FP64, NOT clamped."""
    cf = np.asarray(lab, dtype=np.float32)
    c = cf.astype(np.float64)
    T = c[:, L["T"]]
    beta = 1./T
    tau = c[:, L["tau"]]
    sdu = tau*(c[:, L["da0"]]*c[:, L["u0"]] + c[:, L["u1"]]*c[:, L["da1"]] + c[:, L["u2"]]*c[:, L["da2"]]
               + c[:, L["u3"]]*c[:, L["da3"]]/tau)
    bulkPi = np.zeros(len(c))
    if opt.include_bulk == 1:
        bulkPi = c[:, L["bulkPi"]] if opt.bulk_kind == 0 else c[:, L["bulkPi"]]/HBARC
    sdq = 0.
    pref_q = 0.
    if opt.include_diff == 1:
        sdq = tau*(c[:, L["da0"]]*c[:, L["q0"]] + c[:, L["da1"]]*c[:, L["q1"]]
                   + c[:, L["da2"]]*c[:, L["q2"]] + c[:, L["da3"]]*c[:, L["q3"]]/tau)
        pref_q = c[:, L["nB"]]/(c[:, L["e"]] + c[:, L["P"]])
    n_sf = int((O.SF_X_MAX - O.SF_X_MIN)/O.SF_DX) + 1
    x = O.SF_X_MIN + np.arange(n_sf)*O.SF_DX
    K1t, K2t = special.kn(1, x), special.kn(2, x)
    Et = [special.expn(2*k + 2, x) for k in range(9)] if opt.include_diff == 1 else None
    unit = 1.0/HBARC**3
    out = np.zeros((len(species), len(c)))
    for s, p in enumerate(species):
        m = float(p["mass"])
        sign = int(p["sign"])
        B, S, Q = np.float32(p["baryon"]), np.float32(p["strange"]), np.float32(p["charge"])
        mu = (B*cf[:, L["muB"]] + S*cf[:, L["muS"]] + Q*cf[:, L["muQ"]]).astype(np.float64)
        lam = np.exp(beta*mu)
        R = np.zeros((5, len(c)))
        for n in range(1, (10 if m < 0.7 else 1) + 1):
            arg = n*m*beta
            theta = float(-sign)**(n - 1)
            fug = lam**n
            K2 = O.sf_lerp(K2t, lambda z: special.kn(2, z), arg)
            R[0] += theta/n*fug*K2
            if opt.include_bulk == 1 and opt.bulk_kind == 1:
                K1 = O.sf_lerp(K1t, lambda z: special.kn(1, z), arg)
                R[1] += theta*fug*(m*beta*K1 + 3./n*K2)
                R[2] += theta*fug*K1
            if opt.include_diff == 1:
                R[3] += theta/n*fug*K2
                En = [O.sf_lerp(Et[k], (lambda kk: (lambda z: special.expn(2*kk + 2, z)))(k), arg)
                      for k in range(9)]
                I = np.exp(-arg)/arg*(2./(arg*arg) + 2./arg - 1./2.) + 3./8.*En[0]
                dfac, fac, two_k = 1., 2., 4.
                for k in range(3, 11):
                    dfac *= (2*k - 5)
                    fac *= k
                    two_k *= 2
                    I = I + 3.*dfac/two_k/fac*En[k - 2]
                mb = m*beta
                I = -(mb*mb*mb)*I
                R[4] += n*theta*fug*I
        R[0] *= m*m*T
        R[1] *= m*m/beta
        R[2] *= m*m*m/3.
        R[3] *= m*m/(beta*beta)
        R[4] *= 1./(3.*beta*beta*beta)
        pref = int(p["gspin"])/(2.*np.pi*np.pi)
        tot = unit*pref*sdu*R[0]
        if opt.include_bulk == 1:
            tot = tot + unit*pref*sdu*(-bulkPi*coef[:, 0])*(-coef[:, 1]*R[1] + R[2])
        if opt.include_diff == 1:
            tot = tot + unit*pref*sdq/coef[:, 3]*(-pref_q*R[3] - int(p["baryon"])*R[4])
        out[s] = tot
    return out


def estimate_maximum(lab, coef, species, opt, ztab):
    """[ns][ncell] maximum_guess of EmissionFunctionArray::estimate_maximum."""
    lib = O.clib()
    lib.oracle_legacy_estimate_maximum.restype = C.c_double
    lab = np.ascontiguousarray(lab, dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    zx, zy = ztab
    out = np.zeros((len(species), len(lab)))
    for s, p in enumerate(species):
        for l in range(len(lab)):
            out[s, l] = lib.oracle_legacy_estimate_maximum(
                lab[l].ctypes.data_as(C.c_void_p), coef[l].ctypes.data_as(C.c_void_p), C.byref(opt),
                C.c_double(float(p["mass"])), C.c_int(int(p["sign"])), C.c_int(int(p["gspin"])),
                C.c_int(int(p["baryon"])), C.c_int(int(p["strange"])), C.c_int(int(p["charge"])),
                O._p(zx), O._p(zy), C.c_int(len(zx)))
    return out


def sample(lab, pos, coef, y, species, opt, ztab, seed, ev_begin, mult, cap, cdf=None):
    lib = O.clib()
    lib.oracle_legacy_sample.restype = C.c_int64
    lab = np.ascontiguousarray(lab, dtype=np.float32)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    sp = np.ascontiguousarray(species)
    mult = np.ascontiguousarray(mult, dtype=np.int64)
    nev, ns = mult.shape
    zx, zy = ztab
    out = np.zeros(int(cap), dtype=O.HADRON_DTYPE)
    cell = np.zeros(int(cap), dtype=np.int32)
    tries = np.zeros(int(cap), dtype=np.int32)
    cdf_p = None
    if cdf is not None:
        cdf = np.ascontiguousarray(cdf, dtype=np.float64)
        assert cdf.shape == (ns, len(lab) + 1)
        cdf_p = O._p(cdf)
    n = lib.oracle_legacy_sample(O._p(lab), O._p(pos), C.c_int64(len(lab)), O._p(coef), O._p(y), cdf_p,
                                 O._p(sp), C.c_int(ns), C.byref(opt), O._p(zx), O._p(zy),
                                 C.c_int(len(zx)), C.c_uint64(seed), C.c_int64(ev_begin),
                                 C.c_int64(nev), O._p(mult), O._p(out), C.c_int64(int(cap)),
                                 O._p(cell), O._p(tries))
    if n < 0:
        raise RuntimeError("oracle_legacy_sample failed (%d)" % n)
    return out[:n], cell[:n], tries[:n]
