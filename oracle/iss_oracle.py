"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the yield half of the reference's hot path
and ctypes access to the plain-C restatement of the sampling half (oracle/iss_oracle.c).

Only tests/, __graft_entry__.smoke() and the cpu_baseline leg of bench.py import this module, and
only as the checker of the CUDA engine; nothing in iss_b200/ imports it.

Restated functions (paths relative to the reference tree, chunshen1987/iSS):
  yields()            FSSW::calculate_dN_dxtdy_for_one_particle_species  src/FSSW.cpp:565-715
                      FSSW::calculate_dN_analytic                        src/FSSW.cpp:719-848
  sf tables / lerp    FSSW::initialize_special_function_arrays, get_special_function_K1/K2/K3/En
                                                                         src/FSSW.cpp:1609-1727
  coef_ce()           FSSW::getCENEOSBQSCoefficients                     src/FSSW.cpp:1433-1487
  coef_22mom()        FSSW::get22momNEOSBQSCoefficients                  src/FSSW.cpp:1490-1543
  coef_14mom()        FSSW::getbulkvisCoefficients(T, muB)               src/FSSW.cpp:1379-1430
  coef_poly1()        FSSW::getbulkvisCoefficients(T), kind 1            src/FSSW.cpp:1109-1136
  coef_kappa()        FSSW::get_deltaf_qmu_coeff                         src/FSSW.cpp:1571-1606
  cdf / totals        RandomVariable1DArray ctor                         src/RandomVariable1DArray.cpp:25-52
K_n and E_n come from scipy.special (the reference calls GSL, a third-party dependency that is not
part of the reference tree and whose version CMakeLists.txt:16 does not pin; both are standard
functions and agree to ~1e-14).

Pinning: tests/test_oracle_cpu.py checks yields() against the per-cell x per-species values dumped
from the compiled, unmodified reference (tests/golden/yields_*.npz) for the six runnable CI
fixtures and seven synthetic surfaces covering every delta-f mode, at 1e-9 relative.
"""
import ctypes as C
import os

import numpy as np
from scipy import special

HBARC = 0.197327053
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TABLES = os.path.join(REPO, "iSS_tables")

F = {n: i for i, n in enumerate(
    "tau x y eta da0 da1 da2 da3 ut ux uy uz e T P nB muB muS muQ bulkPi "
    "pixx pixy pixz piyy piyz qx qy qz".split())}

SF_X_MIN, SF_X_MAX, SF_DX = 0.5, 400.0, 0.05


# ----------------------------------------------------------------------------- tables
class Tables:
    """delta-f coefficient tables in the reference's file formats."""

    def __init__(self, table_path=TABLES, afterburner="urqmd", kind=21, include_bulk=1,
                 include_diff=0):
        smash = afterburner.lower() == "smash"
        d = os.path.join(table_path, "deltaf_tables")
        self.ce = None
        self.mom22 = None
        self.mom14 = None
        self.kappa = None
        if kind == 20:
            # not part of the reference tree: the tests generate it (iss_b200/synthetic.py) into
            # the table folder both the compiled reference and the engine read
            f = os.path.join(d, "smash" if smash else "urqmd", "NEoSBQS_22mom_deltafCoeff.dat")
            self.mom22 = np.loadtxt(f, skiprows=1)[:200*200].reshape(200*200, 8)
        if kind == 21:
            f = os.path.join(d, "smash" if smash else "urqmd", "NEoSBQS_CE_deltafCoeff.dat")
            self.ce = np.loadtxt(f, skiprows=1)[:200*200].reshape(200*200, 5)
        if include_bulk == 1 and kind == 11:
            folder = os.path.join(d, "smash_box" if smash else "urqmd")
            tabs = []
            for c in range(3):
                v = np.loadtxt(os.path.join(folder, "c%d.dat" % c), skiprows=3)
                nT, nmu = 101, 81
                tabs.append(v[:, 2].reshape(nmu, nT).T.copy())      # [T][mu]
                if c == 1:
                    # the grid comes from the last table read with valid numbers (FSSW.cpp:1278-1294)
                    self.g14 = (v[0, 0], v[1, 0] - v[0, 0], v[0, 1], v[nT, 1] - v[0, 1])
            self.mom14 = np.array(tabs)
        if include_diff == 1:
            v = np.loadtxt(os.path.join(d, "Coefficients_RTA_diffusion.dat"))
            self.kappa = v[:150*100, 2].reshape(100, 150).T.copy()   # [T][mu]
        # special-function grids
        n = int((SF_X_MAX - SF_X_MIN)/SF_DX) + 1
        x = SF_X_MIN + np.arange(n)*SF_DX
        self.K = np.stack([special.kn(1, x), special.kn(2, x), special.kn(3, x)], axis=1)
        self.E = None
        if include_diff == 1:
            self.E = np.stack([special.expn(2*k + 2, x) for k in range(9)], axis=1)


def sf_lerp(tab, exact, arg):
    """table lerp inside [x_min, x_max - dx], exact function outside (FSSW.cpp:1646-1727)."""
    arg = np.asarray(arg, dtype=np.float64)
    inside = ~((arg < SF_X_MIN) | (arg > SF_X_MAX - SF_DX))
    a = np.where(inside, arg, 1.0)
    idx = ((a - SF_X_MIN)/SF_DX).astype(np.int64)
    frac = (a - SF_X_MIN - idx*SF_DX)/SF_DX
    val = (1. - frac)*tab[idx] + frac*tab[idx + 1]
    if not inside.all():
        val = np.where(inside, val, exact(arg))
    return val


def _neos_bqs_interp(tb, Edec, nB, ncol):
    """bilinear look-up shared by the CE and 22-moment tables (FSSW.cpp:1433-1543): columns
    2..ncol-1; the index in n_B is clamped from above only (from below as well here: a negative
    index is undefined behaviour in the reference)."""
    n = 200
    de = tb[n, 0] - tb[0, 0]
    idx_e = ((Edec - tb[0, 0])/de).astype(np.int64)
    idx_e = np.clip(idx_e, 0, n - 2)
    Ne1, Ne2 = idx_e*n, (idx_e + 1)*n
    e_frac = (Edec - tb[Ne1, 0])/de
    dnB1, dnB2 = tb[Ne1 + 1, 1], tb[Ne2 + 1, 1]
    i1 = np.clip((nB/dnB1).astype(np.int64), 0, n - 2)
    i2 = np.clip((nB/dnB2).astype(np.int64), 0, n - 2)
    f1 = np.minimum(1., (nB - tb[Ne1 + i1, 1])/dnB1)
    f2 = np.minimum(1., (nB - tb[Ne2 + i2, 1])/dnB2)
    ip = []
    for i in range(2, ncol):
        t1 = tb[Ne1 + i1, i]*(1 - f1) + tb[Ne1 + i1 + 1, i]*f1
        t2 = tb[Ne2 + i2, i]*(1 - f2) + tb[Ne2 + i2 + 1, i]*f2
        ip.append(t1*(1. - e_frac) + t2*e_frac)
    return ip


def coef_22mom(tb, Edec, nB):
    """FSSW::get22momNEOSBQSCoefficients (FSSW.cpp:1490-1543): the six interpolated columns as
    they are (c[0] multiplies the shear W factor, c[1..5] enter the bulk term)."""
    return np.stack(_neos_bqs_interp(tb, Edec, nB, 8), axis=1)


def coef_ce(tb, Edec, nB):
    n = 200
    de = tb[n, 0] - tb[0, 0]
    idx_e = ((Edec - tb[0, 0])/de).astype(np.int64)
    idx_e = np.clip(idx_e, 0, n - 2)
    Ne1, Ne2 = idx_e*n, (idx_e + 1)*n
    e_frac = (Edec - tb[Ne1, 0])/de
    dnB1, dnB2 = tb[Ne1 + 1, 1], tb[Ne2 + 1, 1]
    i1 = np.minimum(n - 2, (nB/dnB1).astype(np.int64))
    i2 = np.minimum(n - 2, (nB/dnB2).astype(np.int64))
    f1 = np.minimum(1., (nB - tb[Ne1 + i1, 1])/dnB1)
    f2 = np.minimum(1., (nB - tb[Ne2 + i2, 1])/dnB2)
    ip = []
    for i in range(2, 5):
        t1 = tb[Ne1 + i1, i]*(1 - f1) + tb[Ne1 + i1 + 1, i]*f1
        t2 = tb[Ne2 + i2, i]*(1 - f2) + tb[Ne2 + i2 + 1, i]*f2
        ip.append(t1*(1. - e_frac) + t2*e_frac)
    return np.stack([1./ip[1], 1./3. - ip[0], ip[2]], axis=1)


def _bilinear(tb, ix, iy, fx, fy, clamp):
    nx, ny = tb.shape
    if clamp:
        ix1, ix2 = np.clip(ix, 0, nx - 1), np.clip(ix + 1, 0, nx - 1)
        iy1, iy2 = np.clip(iy, 0, ny - 1), np.clip(iy + 1, 0, ny - 1)
    else:
        ix1, ix2, iy1, iy2 = ix, ix + 1, iy, iy + 1
    f1, f2, f3, f4 = tb[ix1, iy1], tb[ix1, iy2], tb[ix2, iy2], tb[ix2, iy1]
    return f1*(1. - fx)*(1. - fy) + f2*(1. - fx)*fy + f3*fx*fy + f4*fx*(1. - fy)


def coef_14mom(tabs, grid, T, muB):
    T0, dT, mu0, dmu = grid
    ix = ((T - T0)/dT).astype(np.int64)
    iy = ((muB - mu0)/dmu).astype(np.int64)
    fx = (T - T0)/dT - ix
    fy = (muB - mu0)/dmu - iy
    c0, c1, c2 = (_bilinear(tabs[k], ix, iy, fx, fy, True) for k in range(3))
    T3 = T*T*T
    T4 = T3*T
    return np.stack([(c0 - c2)/T4, c1/T3, (4.*c2 - c0)/T4], axis=1)


def coef_poly1(T):
    x = T/HBARC
    p = [x**i for i in range(11)]
    c0 = (642096.624265727 - 8163329.49562861*p[1] + 47162768.4292073*p[2]
          - 162590040.002683*p[3] + 369637951.096896*p[4] - 578181331.809836*p[5]
          + 629434830.225675*p[6] - 470493661.096657*p[7] + 230936465.421*p[8]
          - 67175218.4629078*p[9] + 8789472.32652964*p[10])
    c1 = (1.18171174036192 - 17.6740645873717*p[1] + 136.298469057177*p[2]
          - 635.999435106846*p[3] + 1918.77100633321*p[4] - 3836.32258307711*p[5]
          + 5136.35746882372*p[6] - 4566.22991441914*p[7] + 2593.45375240886*p[8]
          - 853.908199724349*p[9] + 124.260460450113*p[10])
    return np.stack([c0, c1, np.zeros_like(c0)], axis=1)


def coef_kappa(tb, T, muB):
    T0, dT, mu0, dmu = 0.05, 0.001, 0.0, 0.007892
    ix = ((T - T0)/dT).astype(np.int64)
    iy = ((muB - mu0)/dmu).astype(np.int64)
    fx = (T - T0)/dT - ix
    fy = (muB - mu0)/dmu - iy
    nx, ny = tb.shape
    bad = (iy > ny - 2) | (ix > nx - 2) | (iy < 0) | (ix < 0)
    v = _bilinear(tb, np.where(bad, 0, ix), np.where(bad, 0, iy), fx, fy, False)
    return np.where(bad, 1e30, v)


def cell_coefficients(cells, tables, kind, include_bulk, include_diff):
    """[ncell][7]: bulkvisCoefficients c0..c5 and kappa (1 when diffusion is off), the values the
    reference recomputes per species (FSSW.cpp:596-637) and per sample (:982-1008)."""
    c = cells.astype(np.float64)
    n = len(c)
    out = np.zeros((n, 7))
    out[:, 6] = 1.0
    if kind == 21:
        out[:, 0:3] = coef_ce(tables.ce, c[:, F["e"]], c[:, F["nB"]])
    elif kind == 20:
        out[:, 0:6] = coef_22mom(tables.mom22, c[:, F["e"]], c[:, F["nB"]])
    if include_bulk == 1 and kind not in (20, 21):
        if kind == 11:
            out[:, 0:3] = coef_14mom(tables.mom14, tables.g14, c[:, F["T"]], c[:, F["muB"]])
        elif kind == 1:
            out[:, 0:3] = coef_poly1(c[:, F["T"]])
    if include_diff == 1:
        out[:, 6] = coef_kappa(tables.kappa, c[:, F["T"]], c[:, F["muB"]])
    return out


def yields(cells, species, tables, kind=21, include_bulk=1, include_diff=0, coef=None):
    """Per-cell x per-species yields [ns][ncell] in FP64.

    cells: float32 [ncell][28] local-rest-frame records (FO_surf_LRF, ISS_F_* order);
    species: structured array with mass, gspin, baryon, strange, charge, sign."""
    cf = cells.astype(np.float32)
    c = cf.astype(np.float64)
    if coef is None:
        coef = cell_coefficients(cf, tables, kind, include_bulk, include_diff)
    T = c[:, F["T"]]
    beta = 1./T
    sigma = c[:, F["da0"]]
    unit = 1.0/HBARC**3
    bulkPi = np.zeros(len(c))
    if include_bulk == 1:
        bulkPi = c[:, F["bulkPi"]] if kind in (21, 20, 11, 0) else c[:, F["bulkPi"]]/HBARC
    if include_diff == 1:
        dsq = (cf[:, F["qx"]]*cf[:, F["da1"]] + cf[:, F["qy"]]*cf[:, F["da2"]]
               + cf[:, F["qz"]]*cf[:, F["da3"]]).astype(np.float64)       # float arithmetic
        pref_q = c[:, F["nB"]]/(c[:, F["e"]] + c[:, F["P"]])
    K, E = tables.K, tables.E
    out = np.zeros((len(species), len(c)))
    for s, p in enumerate(species):
        m = float(p["mass"])
        sign = int(p["sign"])
        B, S, Q = np.float32(p["baryon"]), np.float32(p["strange"]), np.float32(p["charge"])
        mu = (B*cf[:, F["muB"]] + S*cf[:, F["muS"]] + Q*cf[:, F["muQ"]]).astype(np.float64)
        lam = np.exp(beta*mu)
        trunc = np.where((m < 0.7) & (T > 0.05), 10, 1)
        R = np.zeros((6, len(c)))
        for n in range(1, 11):
            act = trunc >= n
            if not act.any():
                break
            arg = n*m*beta
            theta = float(-sign)**(n - 1)
            fug = lam**n
            K2 = sf_lerp(K[:, 1], lambda x: special.kn(2, x), arg)
            t = np.zeros_like(R)
            t[0] = theta/n*fug*K2
            if include_bulk == 1:
                K1 = sf_lerp(K[:, 0], lambda x: special.kn(1, x), arg)
                if kind in (1, 21):
                    t[1] = theta*fug*(m*beta*K1 + 3*K2/n)
                    t[2] = theta*fug*K1
                elif kind in (11, 20):
                    K3 = sf_lerp(K[:, 2], lambda x: special.kn(3, x), arg)
                    t[1] = theta*fug*K2
                    t[2] = theta*fug*(m*beta*K1 + 3*K2/n)
                    t[3] = theta*fug*(m*beta*K2 + 3*K3/n)
            if include_diff == 1:
                t[4] = theta/n*fug*K2
                En = [sf_lerp(E[:, k], (lambda kk: (lambda x: special.expn(2*kk + 2, x)))(k), arg)
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
                t[5] = n*theta*fug*I
            R += np.where(act, t, 0.0)
        R[0] *= m*m*T
        if include_bulk == 1 and kind in (1, 21):
            R[1] *= m*m/beta
            R[2] *= m*m*m/3.
            R[3] = 0.
        elif include_bulk == 1 and kind in (11, 20):
            R[1] *= m*m/beta
            R[2] *= m*m/(beta*beta)
            R[3] *= m*m*m/(beta*beta)
        else:
            R[1:4] = 0.
        if include_diff == 1:
            R[4] *= m*m/(beta*beta)
            R[5] *= 1./(3.*beta*beta*beta)
        pref = int(p["gspin"])/(2.*np.pi*np.pi)
        Neq = unit*pref*sigma*R[0]
        dNb = 0.
        if include_bulk == 1:
            if kind in (1, 21):
                dNb = unit*pref*sigma*(-bulkPi*coef[:, 0])*(-coef[:, 1]*R[1] + R[2])
            elif kind == 11:
                dNb = unit*pref*sigma*bulkPi*(R[1]*m*m*coef[:, 0] + R[2]*int(p["baryon"])*coef[:, 1]
                                              + R[3]*coef[:, 2])
            elif kind == 20:
                dNb = unit*pref*sigma*bulkPi*(
                    R[1]*m*m*coef[:, 2]
                    + R[2]*(int(p["baryon"])*coef[:, 3] + int(p["strange"])*coef[:, 4]
                            + int(p["charge"])*coef[:, 5])
                    + R[3]*(coef[:, 1] - coef[:, 2]))
        dNq = 0.
        if include_diff == 1:
            dNq = unit*pref*dsq/coef[:, 6]*(-pref_q*R[4] - int(p["baryon"])*R[5])
        out[s] = np.maximum(0., Neq + dNb + dNq)
    return out


def species_totals(y):
    """sum over cells in the reference's order: sequential left-to-right FP64 accumulation of
    max(val, 0) (RandomVariable1DArray.cpp:38-50); np.cumsum accumulates sequentially."""
    return np.cumsum(np.maximum(y, 0.), axis=1)[:, -1]


# ---- the engine's fixed-order evaluation of the same sums (engine design: the reference's
# sequential sum cannot be parallelised; any fixed order is reproducible, and the tile structure is
# what lets GPUs that hold different cell ranges obtain identical bits, iss_b200/csrc/yields.cu)
ENGINE_TILE = 1024


def _hillis_steele32(a):
    """inclusive scan over the last axis (32 lanes) in the order of a warp shuffle-up scan"""
    a = a.copy()
    for d in (1, 2, 4, 8, 16):
        a[..., d:] = a[..., d:] + a[..., :-d]
    return a


def engine_tile_scan(y):
    """y: [ns, ncell] -> (local [ns, ntile, 1024] inclusive prefix inside each tile, tilesum
    [ns, ntile]) in the association of tile_scan_kernel: 4 cells per thread sequentially, warp
    scan of the thread sums, warp bases added sequentially."""
    y = np.maximum(np.asarray(y, dtype=np.float64), 0.)     # the yield kernel already clamps
    ns, ncell = y.shape
    ntile = (ncell + ENGINE_TILE - 1)//ENGINE_TILE
    v = np.zeros((ns, ntile*ENGINE_TILE))
    v[:, :ncell] = y
    v = v.reshape(ns, ntile, 8, 32, 4)                        # tile, warp, lane, item
    v = v.copy()
    for i in (1, 2, 3):
        v[..., i] = v[..., i] + v[..., i - 1]
    incl = _hillis_steele32(v[..., 3])                        # [ns, ntile, 8, 32]
    warp_tot = incl[..., 31]                                  # [ns, ntile, 8]
    wbase = np.zeros_like(warp_tot)
    for w in range(1, 8):
        wbase[..., w] = wbase[..., w - 1] + warp_tot[..., w - 1]
    offset = (incl - v[..., 3]) + wbase[..., None]
    local = v + offset[..., None]
    tilesum = local[:, :, 7, 31, 3]
    return local.reshape(ns, ntile, ENGINE_TILE), tilesum.copy()


def engine_tile_bases(tilesum):
    """[ns, ntile] -> (tilebase [ns, ntile], total [ns]) in the order of tile_base_kernel: groups of
    32 tiles, shuffle scan inside a group, sequential carry between groups."""
    ns, ntile = tilesum.shape
    base = np.zeros((ns, ntile))
    carry = np.zeros(ns)
    for t0 in range(0, ntile, 32):
        v = np.zeros((ns, 32))
        n = min(32, ntile - t0)
        v[:, :n] = tilesum[:, t0:t0 + n]
        incl = _hillis_steele32(v)
        base[:, t0:t0 + n] = (carry[:, None] + (incl - v))[:, :n]
        carry = carry + incl[:, 31]
    return base, carry


def engine_prefix(y):
    """global inclusive prefix [ns, ncell] and totals [ns] exactly as the device computes them"""
    local, tilesum = engine_tile_scan(y)
    base, total = engine_tile_bases(tilesum)
    P = (local + base[:, :, None]).reshape(y.shape[0], -1)[:, :y.shape[1]]
    return P, total


def poisson_pmode(lam):
    """pmf of Poisson(lam) at its mode floor(lam) (engine design, see iss_oracle.c)."""
    lam = np.asarray(lam, dtype=np.float64)
    m = np.floor(lam)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.exp(m*np.log(lam) - lam - special.gammaln(m + 1.0))
    return np.where(lam >= 1e-15, v, 1.0)


# ----------------------------------------------------------------------------- C restatement
class OSpecies(C.Structure):
    _fields_ = [("pid", C.c_int32), ("gspin", C.c_int32), ("baryon", C.c_int32),
                ("strange", C.c_int32), ("charge", C.c_int32), ("sign", C.c_int32),
                ("decay_idx", C.c_int32), ("reserved", C.c_int32), ("mass", C.c_double)]


class OOptions(C.Structure):
    _fields_ = [("hydro_mode", C.c_int32), ("include_shear", C.c_int32),
                ("include_bulk", C.c_int32), ("include_diff", C.c_int32),
                ("bulk_kind", C.c_int32), ("model", C.c_int32), ("lcc", C.c_int32),
                ("reserved", C.c_int32), ("para1", C.c_double), ("y_LB", C.c_double),
                ("y_RB", C.c_double)]


HADRON_DTYPE = np.dtype([("pid", "<i4"), ("mass", "<f4"), ("E", "<f4"), ("px", "<f4"),
                         ("py", "<f4"), ("pz", "<f4"), ("t", "<f4"), ("x", "<f4"), ("y", "<f4"),
                         ("z", "<f4")])
DSPECIES_DTYPE = np.dtype([("pid", "<i4"), ("stable", "<i4"), ("n_channels", "<i4"),
                           ("first_channel", "<i4"), ("baryon", "<i4"), ("strange", "<i4"),
                           ("charge", "<i4"), ("reserved", "<i4"), ("mass", "<f8"),
                           ("width", "<f8")])
DCHANNEL_DTYPE = np.dtype([("n_part", "<i4"), ("daughter", "<i4", 5), ("br", "<f8")])
assert DCHANNEL_DTYPE.itemsize == 32

_lib = None


def clib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "libiss_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(path)
        _lib.oracle_sample.restype = C.c_int64
        _lib.oracle_decay.restype = C.c_int64
        _lib.oracle_decay_once_many.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def stream_uniforms(seed, stream, species, event, draw, n):
    out = np.zeros(n)
    clib().oracle_stream_uniforms(C.c_uint64(seed), C.c_uint32(stream), C.c_uint32(species),
                                  C.c_uint32(event), C.c_uint32(draw), C.c_int(n), _p(out))
    return out


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    clib().oracle_philox(_p(c), _p(k), _p(out))
    return out


def multiplicities(lam, pmode, species, nev, ev_begin, seed, model=30, lcc=0, para1=0.16):
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    pmode = np.ascontiguousarray(pmode, dtype=np.float64)
    sp = np.ascontiguousarray(species)
    ns = len(sp)
    mult = np.zeros((nev, ns), dtype=np.int64)
    outc = np.zeros((nev, ns), dtype=np.int64)
    clib().oracle_multiplicities(_p(lam), _p(pmode), _p(sp), C.c_int(ns), C.c_int64(nev),
                                 C.c_int64(ev_begin), C.c_uint64(seed), C.c_int(model),
                                 C.c_double(para1), C.c_int(lcc), _p(mult), _p(outc))
    return mult, outc


def make_options(hydro_mode=2, include_shear=0, include_bulk=0, include_diff=0, bulk_kind=21,
                 model=30, lcc=0, y_LB=-5.0, y_RB=5.0):
    o = OOptions()
    o.hydro_mode, o.include_shear, o.include_bulk, o.include_diff = (hydro_mode, include_shear,
                                                                     include_bulk, include_diff)
    o.bulk_kind, o.model, o.lcc, o.y_LB, o.y_RB = bulk_kind, model, lcc, y_LB, y_RB
    return o


def sample(cells, coef, y, species, opt, seed, ev_begin, mult, out_count_total):
    cells = np.ascontiguousarray(cells, dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    sp = np.ascontiguousarray(species)
    mult = np.ascontiguousarray(mult, dtype=np.int64)
    nev, ns = mult.shape
    cap = int(out_count_total)
    out = np.zeros(cap, dtype=HADRON_DTYPE)
    cell = np.zeros(cap, dtype=np.int32)
    tries = np.zeros(cap, dtype=np.int32)
    n = clib().oracle_sample(_p(cells), C.c_int64(len(cells)), _p(coef), _p(y), _p(sp), C.c_int(ns),
                             C.byref(opt), C.c_uint64(seed), C.c_int64(ev_begin), C.c_int64(nev),
                             _p(mult), _p(out), C.c_int64(cap), _p(cell), _p(tries))
    if n < 0:
        raise RuntimeError("oracle_sample failed (%d)" % n)
    return out[:n], cell[:n], tries[:n]


def sample_momentum(m, T, mu, sign, n, seed):
    out = np.zeros(n)
    rc = clib().oracle_sample_momentum(C.c_double(m), C.c_double(T), C.c_double(mu), C.c_int(sign),
                                       C.c_int64(n), C.c_uint64(seed), _p(out))
    if rc != 0:
        raise RuntimeError("momentum table range")
    return out


def momentum_table(r):
    n = clib().oracle_momentum_table(C.c_int(r), None)
    t = np.zeros((4, n))
    clib().oracle_momentum_table(C.c_int(r), _p(t))
    return t


def read_pdg_table(path, quantum_statistics=True):
    """pdg-*.dat with generated anti-baryons in the order of read_FOdata::read_resonances_list
    (src/readindata.cpp:971-1116) / particle_decay::read_resonances_list (particle_decay.cpp:33-172).
    Returns (dspecies, dchannels) arrays for the C decay oracle."""
    tok = open(path).read().split()
    pos = 0
    parts = []
    while pos + 12 <= len(tok):
        monval = int(tok[pos]); name = tok[pos + 1]
        mass, width = float(tok[pos + 2]), float(tok[pos + 3])
        gspin, baryon, strange, charm, bottom, gisospin, charge, decays = (int(x) for x in tok[pos + 4:pos + 12])
        pos += 12
        chans = []
        for _ in range(decays):
            npart = int(tok[pos + 1]); br = float(tok[pos + 2])
            d = [int(x) for x in tok[pos + 3:pos + 8]]
            pos += 8
            chans.append((npart, br, d))
        p = dict(pid=monval, mass=mass, width=width, gspin=gspin, baryon=baryon, strange=strange,
                 charge=charge, chans=chans, name=name)
        p["stable"] = 1 if (chans and chans[0][0] == 1) else 0
        parts.append(p)
        if baryon > 0:
            a = dict(p)
            a.update(pid=-monval, baryon=-baryon, strange=-strange, charge=-charge)
            ac = []
            by_pid = {q["pid"]: q for q in parts}
            for (npart, br, d) in chans:
                dd = []
                for x in d:
                    if x == 0:
                        dd.append(0)
                        continue
                    q = by_pid.get(x)
                    if q is None:
                        dd.append(-x)
                    else:
                        dd.append(x if (q["baryon"] == 0 and q["charge"] == 0 and q["strange"] == 0) else -x)
                ac.append((npart, br, dd))
            a["chans"] = ac
            parts.append(a)
    idx = {p["pid"]: i for i, p in enumerate(parts)}
    ds = np.zeros(len(parts), dtype=DSPECIES_DTYPE)
    ch = []
    for i, p in enumerate(parts):
        ds[i] = (p["pid"], p["stable"], len(p["chans"]), len(ch), p["baryon"], p["strange"],
                 p["charge"], 0, p["mass"], p["width"])
        for (npart, br, d) in p["chans"]:
            ch.append((npart, [idx.get(x, -1) if x != 0 else -1 for x in d], br))
    dc = np.zeros(len(ch), dtype=DCHANNEL_DTYPE)
    for i, (npart, d, br) in enumerate(ch):
        dc[i] = (npart, d, br)
    return ds, dc


def decay(hadrons, event_off, ev_begin, ds, dc, seed, cap_factor=6):
    h = np.ascontiguousarray(hadrons, dtype=HADRON_DTYPE)
    off = np.ascontiguousarray(event_off, dtype=np.int64)
    nev = len(off) - 1
    cap = int(len(h)*cap_factor + 1024)
    out = np.zeros(cap, dtype=HADRON_DTYPE)
    off_out = np.zeros(nev + 1, dtype=np.int64)
    n = clib().oracle_decay(_p(h), _p(off), C.c_int64(nev), C.c_int64(ev_begin), _p(ds),
                            C.c_int(len(ds)), _p(dc), C.c_uint64(seed), _p(out), C.c_int64(cap),
                            _p(off_out))
    if n < 0:
        raise RuntimeError("oracle_decay failed (%d)" % n)
    return out[:n], off_out


def decay_once_many(pid, n, seed, ds, dc, mother):
    m = np.zeros(1, dtype=HADRON_DTYPE)
    m[0] = mother
    out = np.zeros(3*n, dtype=HADRON_DTYPE)
    nd = np.zeros(n, dtype=np.int32)
    w = clib().oracle_decay_once_many(C.c_int(pid), C.c_int64(n), C.c_uint64(seed), _p(ds),
                                      C.c_int(len(ds)), _p(dc), _p(m), _p(out), _p(nd))
    if w < 0:
        raise RuntimeError("oracle_decay_once_many failed (%d)" % w)
    return out[:w], nd
