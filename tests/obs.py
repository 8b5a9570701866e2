"""Observables shared by the golden-fixture generator and the statistical parity tests
(SURVEY.md appendix E).  Pure numpy; works on structured hadron arrays of capi.HADRON_DTYPE."""
import numpy as np

PIDS = [211, -211, 111, 321, -321, 2212, -2212, 3122, -3122]
PT_EDGES = np.linspace(0.0, 4.0, 41)
Y_EDGES = np.linspace(-6.0, 6.0, 49)
PHI_EDGES = np.linspace(-np.pi, np.pi, 33)
V2_PT_EDGES = np.linspace(0.0, 3.0, 11)


def read_reference_bin(path):
    """particle_samples.bin of the reference (FSSW.cpp:529-561): int N, then N x
    {int pid, float mass,t,x,y,z,E,px,py,pz}.  Returns (records, event_offsets)."""
    raw = np.fromfile(path, dtype=np.uint8)
    rec_dt = np.dtype([("pid", "<i4"), ("mass", "<f4"), ("t", "<f4"), ("x", "<f4"), ("y", "<f4"),
                       ("z", "<f4"), ("E", "<f4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4")])
    pos = 0
    chunks = []
    off = [0]
    while pos < len(raw):
        n = int(raw[pos:pos + 4].view("<i4")[0])
        pos += 4
        chunks.append(raw[pos:pos + 40*n].view(rec_dt))
        pos += 40*n
        off.append(off[-1] + n)
    rec = np.concatenate(chunks) if chunks else np.zeros(0, dtype=rec_dt)
    return rec, np.asarray(off, dtype=np.int64)


def summarize(h, off):
    """Histograms and moments of a hadron sample; h has fields pid, mass, E, px, py, pz, t, x, y, z."""
    nev = len(off) - 1
    out = {"nev": np.int64(nev)}
    px, py, pz, E = (h[k].astype(np.float64) for k in ("px", "py", "pz", "E"))
    pT = np.hypot(px, py)
    mT = np.sqrt(h["mass"].astype(np.float64)**2 + pT**2)
    y = np.arcsinh(pz/mT)
    phi = np.arctan2(py, px)
    ev = np.repeat(np.arange(nev), np.diff(off))
    for pid in PIDS:
        m = h["pid"] == pid
        tag = "p%d" % pid if pid > 0 else "m%d" % (-pid)
        out[tag + "_pt"] = np.histogram(pT[m], PT_EDGES)[0]
        out[tag + "_y"] = np.histogram(y[m], Y_EDGES)[0]
        out[tag + "_phi"] = np.histogram(phi[m], PHI_EDGES)[0]
        c2 = np.cos(2*phi[m])
        out[tag + "_v2num"] = np.histogram(pT[m], V2_PT_EDGES, weights=c2)[0]
        out[tag + "_v2sq"] = np.histogram(pT[m], V2_PT_EDGES, weights=c2*c2)[0]
        out[tag + "_v2den"] = np.histogram(pT[m], V2_PT_EDGES)[0]
        nper = np.bincount(ev[m], minlength=nev).astype(np.float64)
        out[tag + "_n"] = np.array([nper.sum(), (nper**2).sum()])
    # total four-momentum per event: sums and sums of squares
    P = np.stack([np.bincount(ev, weights=a, minlength=nev) for a in (E, px, py, pz)])
    out["P_sum"] = P.sum(axis=1)
    out["P_sq"] = (P**2).sum(axis=1)
    # space-time: t, x, y, z means (all hadrons)
    for k in ("t", "x", "y", "z"):
        a = h[k].astype(np.float64)
        out["pos_" + k] = np.array([a.sum(), (a*a).sum(), len(a)])
    # T^{mu nu} numerator sum p^mu p^nu / p^0
    p4 = np.stack([E, px, py, pz])
    out["Tmunu"] = np.einsum("in,jn->ij", p4, p4/np.where(E > 0, E, 1.0))
    # per-pid totals for every species present
    pids, cnt = np.unique(h["pid"], return_counts=True)
    out["all_pids"] = pids.astype(np.int64)
    out["all_counts"] = cnt.astype(np.int64)
    return out


def chi2_two_hist(a, b, na, nb, min_count=20):
    """Two-sample chi2 of count histograms a, b taken from na and nb events (normalised per event).
    Returns (chi2, ndf)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = (a + b) >= min_count
    if m.sum() == 0:
        return 0.0, 0
    ra, rb = a[m]/na, b[m]/nb
    var = a[m]/na**2 + b[m]/nb**2
    return float(((ra - rb)**2/var).sum()), int(m.sum())
