"""Helpers shared by the legacy-sampler tests (MC_sampling = 2, EmissionFunctionArray)."""
import os
import sys

import numpy as np

import cases

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import iss_oracle as orc  # noqa: E402
import legacy_oracle as lgo  # noqa: E402

# bulk kinds 2-4, Boltzmann statistics and unrestricted delta f (MORE_CASES) run on the GPU like
# the first three; on the CPU the oracle is pinned against the reference for all of them
BASE_CASES = ["l3d_shear", "l3d_bulk1_diff", "l2d_ideal_smash"]
MORE_CASES = ["l3d_bulk2", "l3d_bulk3_norestrict", "l3d_bulk4_boltzmann", "l3d_bulk0"]
YIELD_CASES = BASE_CASES + MORE_CASES
STATS_CASES = ["cell_shear", "surf3d_bulk1", "cell_lcc", "cell_bulk3", "surf3d_bulk0"]


def species_array(sp):
    a = np.zeros(len(sp), dtype=[("pid", "<i4"), ("gspin", "<i4"), ("baryon", "<i4"),
                                 ("strange", "<i4"), ("charge", "<i4"), ("sign", "<i4"),
                                 ("decay_idx", "<i4"), ("reserved", "<i4"), ("mass", "<f8")])
    a["pid"], a["mass"], a["gspin"] = sp[:, 0], sp[:, 1], sp[:, 2]
    a["baryon"], a["strange"], a["charge"], a["sign"] = sp[:, 3], sp[:, 4], sp[:, 5], sp[:, 6]
    a["decay_idx"] = -1
    return a


def parameters(g):
    """parameter file + overrides of a legacy fixture, keys lower-cased like ParameterReader."""
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
    if par["sample_pt_up_to"] < 0:      # emissionfunction.cpp:3302-3305: last row of the pT table
        pT = np.loadtxt(os.path.join(orc.TABLES, "bin_tables", "pT_gauss_table.dat"))
        par["sample_pt_up_to"] = float(pT[-1, 0])
    return par


def oracle_options(par):
    return lgo.make_opt(include_shear=int(par["include_deltaf_shear"]),
                        include_bulk=int(par["include_deltaf_bulk"]),
                        bulk_kind=int(par["bulk_deltaf_kind"]),
                        include_diff=int(par["include_deltaf_diffusion"]),
                        restrict_deltaf=int(par["restrict_deltaf"]),
                        deltaf_max_ratio=par["deltaf_max_ratio"],
                        boost_invariant=int(int(par["hydro_mode"]) != 2),
                        pT_to=par["sample_pt_up_to"], y_range=par["sample_y_minus_eta_s_range"],
                        y_LB=par["y_lb"], y_RB=par["y_rb"],
                        lcc=int(par.get("local_charge_conservation", 0)))


def engine_options(par):
    """keyword arguments of capi.Engine.legacy_setup"""
    return dict(include_deltaf_shear=int(par["include_deltaf_shear"]),
                include_deltaf_bulk=int(par["include_deltaf_bulk"]),
                bulk_deltaf_kind=int(par["bulk_deltaf_kind"]),
                include_deltaf_diffusion=int(par["include_deltaf_diffusion"]),
                restrict_deltaf=int(par["restrict_deltaf"]), deltaf_max_ratio=par["deltaf_max_ratio"],
                sample_pT_up_to=par["sample_pt_up_to"],
                sample_y_minus_eta_s_range=par["sample_y_minus_eta_s_range"])
