"""Helpers for the spectra golden fixtures (tests/golden/spectra_*.npz)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
FIX = os.path.join(HERE, "fixtures")
TABLES = os.path.join(REPO, "iSS_tables")

SPECTRA = ["sp3d_shear", "sp3d_bulk1_diff", "sp3d_bulk2_norestrict", "sp3d_bulk3_pos", "sp3d_bulk4",
           "sp3d_bulk0_quirk", "sp2d_ideal_boltzmann"]


def load(name):
    return np.load(os.path.join(GOLDEN, "spectra_%s.npz" % name), allow_pickle=False)


def read_params(path):
    out = {}
    for line in open(path):
        line = line.split("#")[0]
        if "=" in line:
            k, v = line.split("=")[:2]
            out[k.strip()] = float(v)
    return out


def options_of(g):
    p = read_params(os.path.join(FIX, str(g["param"])))
    for kv in g["overrides"]:
        k, v = str(kv).split("=")
        p[k] = float(v)
    return dict(include_shear=int(p["include_deltaf_shear"]), include_bulk=int(p["include_deltaf_bulk"]),
                bulk_kind=int(p["bulk_deltaf_kind"]), include_diff=int(p["include_deltaf_diffusion"]),
                restrict_deltaf=int(p["restrict_deltaf"]), deltaf_max_ratio=float(p["deltaf_max_ratio"]),
                use_pos_dN_only=int(p["use_pos_dN_only"]))


def species_of(g):
    return [dict(pid=int(r[0]), mass=float(r[1]), gspin=float(r[2]), baryon=int(r[3]), strange=int(r[4]),
                 charge=int(r[5]), sign=float(r[6])) for r in g["species"]]


def bin_tables():
    d = os.path.join(TABLES, "bin_tables")
    return (np.loadtxt(os.path.join(d, "pT_gauss_table.dat")),
            np.loadtxt(os.path.join(d, "phi_gauss_table.dat")),
            np.loadtxt(os.path.join(d, "eta_uni_table.dat")))


def kappa_table():
    v = np.loadtxt(os.path.join(TABLES, "deltaf_tables", "Coefficients_RTA_diffusion.dat"))
    return v[:150*100, 2].reshape(100, 150).T.copy()   # [T][mu]
