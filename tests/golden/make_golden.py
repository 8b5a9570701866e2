#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(compiled by oracle/Makefile into oracle/_ref/) in THIS container.  The fixtures travel to the
GPU box; /root/reference does not.

    python tests/golden/make_golden.py [yields] [stats] [momentum] [decay] [writers] [spectra] [legacy]

yields   : per-cell x per-species yields (FSSW::calculate_dN_dxtdy_for_one_particle_species)
           + the reference's local-rest-frame surface + species order, for the six runnable
           one-cell CI fixtures and for small synthetic surfaces in every delta-f mode.
stats    : histograms (tests/obs.py) of particle_samples.bin written by the reference's own
           sampler (iSS.e) with fixed seeds, >= 10^4 events each.
momentum : |p| samples of MomentumSamplerShell::Sample_a_momentum reduced to histograms.
decay    : daughters of particle_decay::perform_decays for a few resonances.
legacy   : the MC_sampling = 2 path (EmissionFunctionArray "conventional" sampler): lab-frame cells,
           per-cell yields and estimate_maximum values (ref_driver legacy) and histograms of the
           reference's own samples.
spectra  : dN/(pT dpT dphi dy) tables of EmissionFunctionArray::calculate_dN_pTdpTdphidy and the
           flow tables of calculate_flows for a few species on small synthetic surfaces, with the
           lab-frame cells the reference used.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(HERE))
from iss_b200 import synthetic  # noqa: E402
import obs  # noqa: E402

REF = os.path.join(REPO, "oracle", "_ref")
FIX = os.path.join(REPO, "tests", "fixtures")
REF_TABLES = "/root/reference/iSS_tables"

# name -> (music_input suffix, parameter file, surface fixture, overrides)
ONE_CELL = {
    "ideal1": ("91", "iSS_parameters_ideal.dat", "testIdealOneFluidCell1.dat", ["bulk_deltaf_kind=21"]),
    "ideal2": ("9", "iSS_parameters_ideal.dat", "testIdealOneFluidCell2.dat", ["bulk_deltaf_kind=21"]),
    "ideal3": ("12", "iSS_parameters_ideal.dat", "testIdealOneFluidCell3.dat", ["bulk_deltaf_kind=21"]),
    "ideal4": ("14", "iSS_parameters_ideal.dat", "testIdealOneFluidCell4.dat", ["bulk_deltaf_kind=21"]),
    "viscous1": ("91", "iSS_parameters_CEdeltaf.dat", "testViscousOneFluidCell1.dat", []),
    "viscous2": ("9", "iSS_parameters_CEdeltaf.dat", "testViscousOneFluidCell2.dat", []),
    # the reference's own CI cases 3 and 4: bulk_deltaf_kind = 20 (22-moment), whose table is not in
    # the reference tree -> run with the synthetic table of iss_b200/synthetic.py (needs_22mom)
    "viscous3": ("14", "iSS_parameters.dat", "testViscousOneFluidCell3.dat", []),
    "viscous4": ("14", "iSS_parameters.dat", "testViscousOneFluidCell4.dat", []),
}

# name -> dict(generator kwargs, parameter file, overrides)
SYNTH = {
    "s3d_ce": dict(gen=dict(ncell=240, seed=12345, eos=9), param="iSS_parameters_CEdeltaf.dat", over=[]),
    "s3d_ce_diff": dict(gen=dict(ncell=240, seed=2024, eos=14, rhob=1, diffusion=1, binary=1),
                        param="iSS_parameters_CEdeltaf.dat", over=["include_deltaf_diffusion=1"]),
    "s3d_14mom": dict(gen=dict(ncell=240, seed=12345, eos=14, rhob=1),
                      param="iSS_parameters.dat",
                      over=["bulk_deltaf_kind=11", "include_deltaf_bulk=1", "oracle_fix_14mom_c0=1"]),
    "s2d_smash_ce": dict(gen=dict(ncell=240, seed=12345, eos=91, boost_invariant=True),
                         param="iSS_parameters_CEdeltaf.dat", over=["hydro_mode=1"]),
    "s3d_ideal_b": dict(gen=dict(ncell=240, seed=777, eos=12, rhob=1),
                        param="iSS_parameters_ideal.dat", over=["bulk_deltaf_kind=21"]),
    "s3d_bulk1": dict(gen=dict(ncell=240, seed=99, eos=9), param="iSS_parameters_CEdeltaf.dat",
                      over=["bulk_deltaf_kind=1"]),
    "s2d_urqmd_bin": dict(gen=dict(ncell=240, seed=4242, eos=9, boost_invariant=True, binary=1),
                          param="iSS_parameters_CEdeltaf.dat", over=["hydro_mode=1"]),
    "s3d_boltzmann": dict(gen=dict(ncell=120, seed=5, eos=9), param="iSS_parameters_CEdeltaf.dat",
                          over=["quantum_statistics=0"]),
    # 22-moment delta f (kind 20) on the synthetic table: shear + bulk, and with diffusion on top
    "s3d_22mom": dict(gen=dict(ncell=240, seed=2020, eos=14, rhob=1), param="iSS_parameters.dat",
                      over=[]),
    "s3d_22mom_diff": dict(gen=dict(ncell=240, seed=2021, eos=14, rhob=1, diffusion=1, binary=1),
                           param="iSS_parameters.dat", over=["include_deltaf_diffusion=1"]),
}


def needs_22mom(param, over):
    """bulk_deltaf_kind = 20 after the overrides?"""
    kind = None
    for line in open(os.path.join(FIX, param)):
        line = line.split("#")[0]
        if "=" in line and line.split("=")[0].strip() == "bulk_deltaf_kind":
            kind = int(float(line.split("=")[1]))
    for kv in over:
        if kv.startswith("bulk_deltaf_kind="):
            kind = int(float(kv.split("=")[1]))
    return kind == 20


def workdir(with_22mom=False):
    d = tempfile.mkdtemp(prefix="iss_golden_")
    if with_22mom:
        synthetic.tables_with_22mom(os.path.join(d, "iSS_tables"), REF_TABLES)
    else:
        os.symlink(REF_TABLES, os.path.join(d, "iSS_tables"))
    return d


def run(cmd, cwd, log):
    with open(log, "w") as f:
        subprocess.run(cmd, cwd=cwd, stdout=f, stderr=subprocess.STDOUT, check=True)


def read_dump(prefix):
    with open(prefix + ".lrf.bin", "rb") as f:
        n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        lrf = np.fromfile(f, dtype=np.float32).reshape(n, 28)
    sp = np.loadtxt(prefix + ".species.txt", ndmin=2)
    with open(prefix + ".yields.bin", "rb") as f:
        ns, nc = np.fromfile(f, dtype=np.int64, count=2)
        y = np.fromfile(f, dtype=np.float64).reshape(int(ns), int(nc))
    return lrf, sp, y


def golden_yields(only=None):
    for name, (mi, param, surf, over) in ONE_CELL.items():
        if only and name not in only:
            continue
        d = workdir(needs_22mom(param, over))
        case = os.path.join(d, "case")
        os.makedirs(case)
        shutil.copy(os.path.join(FIX, "music_input_" + mi), os.path.join(case, "music_input"))
        shutil.copy(os.path.join(FIX, surf), os.path.join(case, surf))
        run([os.path.join(REF, "ref_driver"), "yields", os.path.join(FIX, param), "case", surf,
             os.path.join(d, "out")] + over, d, os.path.join(d, "log"))
        lrf, sp, y = read_dump(os.path.join(d, "out"))
        np.savez_compressed(os.path.join(HERE, "yields_%s.npz" % name), lrf=lrf, species=sp, yields=y,
                            music_input=mi, param=param, surface=surf, overrides=np.array(over))
        print(name, lrf.shape, y.shape, "sum=%.17g" % y.sum())
        shutil.rmtree(d)
    for name, spec in SYNTH.items():
        if only and name not in only:
            continue
        d = workdir(needs_22mom(spec["param"], spec["over"]))
        g = dict(spec["gen"])
        cells = synthetic.make_case(os.path.join(d, "case"), **g)
        run([os.path.join(REF, "ref_driver"), "yields", os.path.join(FIX, spec["param"]), "case",
             "surface.dat", os.path.join(d, "out")] + spec["over"], d, os.path.join(d, "log"))
        lrf, sp, y = read_dump(os.path.join(d, "out"))
        np.savez_compressed(os.path.join(HERE, "yields_%s.npz" % name), lrf=lrf, species=sp, yields=y,
                            cells=cells, gen=np.array(sorted(g.items()), dtype=object).astype(str),
                            param=spec["param"], overrides=np.array(spec["over"]))
        print(name, lrf.shape, y.shape, "sum=%.17g" % y.sum())
        shutil.rmtree(d)


# statistical goldens: the reference's own sampler, fixed seed
STATS = {
    # one moving cell with shear stress, UrQMD list, CE delta-f (Viscous2 fixture scaled to a
    # smaller volume so that 10^4 events are cheap): cell line written below
    "cell_shear_ce": dict(music="9", param="iSS_parameters_CEdeltaf.dat", nev=20000, seed=1,
                          cell="viscous2_small", over=[]),
    "cell_bulk_ce_smash": dict(music="91", param="iSS_parameters_CEdeltaf.dat", nev=20000, seed=2,
                               cell="viscous1_small", over=[]),
    "cell_ideal_muB": dict(music="14", param="iSS_parameters_ideal.dat", nev=20000, seed=3,
                           cell="ideal4_small", over=["bulk_deltaf_kind=21"]),
    # Grad (14-moment) shear delta f with mu_B, mu_S, mu_Q != 0 (bulk off: the reference's kind-11
    # bulk table is uninitialised)
    "cell_shear_grad_muB": dict(music="14", param="iSS_parameters.dat", nev=20000, seed=6,
                                cell="viscous4_small",
                                over=["bulk_deltaf_kind=11", "include_deltaf_bulk=0",
                                      "include_deltaf_shear=1"]),
    # local charge conservation: positive species paired with their conjugates from the same cell
    "cell_lcc": dict(music="9", param="iSS_parameters_CEdeltaf.dat", nev=20000, seed=7,
                     cell="viscous2_small", over=["local_charge_conservation=1"]),
    # 22-moment delta f (kind 20, the reference CI default) on the synthetic table: one cell with
    # shear + bulk, and a 3+1D surface with n_B
    "cell_22mom": dict(music="14", param="iSS_parameters.dat", nev=20000, seed=8,
                       cell="viscous4_small", over=[]),
    "surf3d_22mom": dict(music=None, param="iSS_parameters.dat", nev=10000, seed=9,
                         gen=dict(ncell=2000, seed=2020, eos=14, rhob=1), over=[]),
    "surf3d_ce_diff": dict(music=None, param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=4,
                           gen=dict(ncell=2000, seed=2024, eos=14, rhob=1, diffusion=1, binary=1),
                           over=["include_deltaf_diffusion=1"]),
    "surf2d_ce_decay": dict(music=None, param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=5,
                            gen=dict(ncell=500, seed=12345, eos=9, boost_invariant=True),
                            over=["hydro_mode=1", "perform_decays=1", "y_LB=-2", "y_RB=2"]),
}


def small_cell(fixture, scale):
    """one-cell fixture with da0 scaled (volume = tau*da0)."""
    v = np.loadtxt(os.path.join(FIX, fixture)).ravel()
    v[4] *= scale
    return v


def golden_stats(only=None):
    for name, spec in STATS.items():
        if only and name not in only:
            continue
        d = workdir(needs_22mom(spec["param"], spec["over"]))
        case = os.path.join(d, "case")
        os.makedirs(case)
        extra = {}
        if spec["music"] is not None:
            shutil.copy(os.path.join(FIX, "music_input_" + spec["music"]),
                        os.path.join(case, "music_input"))
            fixture = {"viscous2_small": "testViscousOneFluidCell2.dat",
                       "viscous1_small": "testViscousOneFluidCell1.dat",
                       "ideal4_small": "testIdealOneFluidCell4.dat",
                       "viscous4_small": "testViscousOneFluidCell4.dat"}[spec["cell"]]
            v = small_cell(fixture, 0.01)
            np.savetxt(os.path.join(case, "surface.dat"), v[None, :], fmt="%.16e")
            extra["cell_line"] = v
        else:
            extra["cells"] = synthetic.make_case(case, **spec["gen"])
            extra["gen"] = np.array(sorted(spec["gen"].items()), dtype=object).astype(str)
        over = ["number_of_repeated_sampling=%d" % spec["nev"], "randomSeed=%d" % spec["seed"],
                "use_OSCAR_format=0", "use_gzip_format=0", "use_binary_format=1", "perform_checks=0",
                "output_samples_into_files=0"] + spec["over"]
        run([os.path.join(REF, "iSS.e"), os.path.join(FIX, spec["param"]), "case", "surface.dat"] + over,
            d, os.path.join(d, "log"))
        rec, off = obs.read_reference_bin(os.path.join(d, "particle_samples.bin"))
        s = obs.summarize(rec, off)
        np.savez_compressed(os.path.join(HERE, "stats_%s.npz" % name), param=spec["param"],
                            overrides=np.array(spec["over"]), music=str(spec["music"]), **extra, **s)
        print(name, "events", len(off) - 1, "hadrons", len(rec))
        shutil.rmtree(d)



# ---- legacy "conventional" sampler (MC_sampling = 2, EmissionFunctionArray) ---------------------
LEGACY_COMMON = ["MC_sampling=2", "store_samples_in_memory=1", "output_samples_into_files=0"]
LEGACY = {
    "l3d_shear": dict(gen=dict(ncell=60, seed=31, eos=9), param="iSS_parameters_CEdeltaf.dat",
                      over=["include_deltaf_shear=1", "include_deltaf_bulk=0", "bulk_deltaf_kind=1"]),
    "l3d_bulk1_diff": dict(gen=dict(ncell=60, seed=32, eos=14, rhob=1, diffusion=1, binary=1),
                           param="iSS_parameters_CEdeltaf.dat",
                           over=["include_deltaf_shear=1", "include_deltaf_bulk=1", "bulk_deltaf_kind=1",
                                 "include_deltaf_diffusion=1", "restrict_deltaf=1"]),
    "l2d_ideal_smash": dict(gen=dict(ncell=60, seed=33, eos=91, boost_invariant=True),
                            param="iSS_parameters_ideal.dat",
                            over=["hydro_mode=1", "bulk_deltaf_kind=1", "grouping_particles=0"]),
}
# further delta-f modes of the legacy class, pinned on the CPU only so far (tests/test_legacy_cpu.py)
LEGACY.update({
    "l3d_bulk2": dict(gen=dict(ncell=40, seed=41, eos=9), param="iSS_parameters_CEdeltaf.dat",
                      over=["include_deltaf_shear=1", "include_deltaf_bulk=1", "bulk_deltaf_kind=2"]),
    "l3d_bulk3_norestrict": dict(gen=dict(ncell=40, seed=42, eos=14, rhob=1),
                                 param="iSS_parameters_CEdeltaf.dat",
                                 over=["include_deltaf_shear=0", "include_deltaf_bulk=1",
                                       "bulk_deltaf_kind=3", "restrict_deltaf=0"]),
    "l3d_bulk4_boltzmann": dict(gen=dict(ncell=40, seed=43, eos=9), param="iSS_parameters_CEdeltaf.dat",
                                over=["include_deltaf_shear=1", "include_deltaf_bulk=1",
                                      "bulk_deltaf_kind=4", "quantum_statistics=0"]),
    # kind 0: 14-moment coefficients from BulkDf_Coefficients_Hadrons_s95p-v0-PCE.dat
    "l3d_bulk0": dict(gen=dict(ncell=40, seed=44, eos=9), param="iSS_parameters_CEdeltaf.dat",
                      over=["include_deltaf_shear=1", "include_deltaf_bulk=1", "bulk_deltaf_kind=0"]),
})
LEGACY_STATS = {
    # one moving cell with shear stress (Viscous2 fixture, volume scaled down), UrQMD list
    "cell_shear": dict(music="9", param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=11,
                       cell="testViscousOneFluidCell2.dat", scale=0.002,
                       over=["include_deltaf_shear=1", "include_deltaf_bulk=0", "bulk_deltaf_kind=1"]),
    # local charge conservation: negative species skipped, positive ones paired (:3326-3336, 3517-3546)
    "cell_lcc": dict(music="9", param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=13,
                     cell="testViscousOneFluidCell2.dat", scale=0.002,
                     over=["include_deltaf_shear=1", "include_deltaf_bulk=0", "bulk_deltaf_kind=1",
                           "local_charge_conservation=1"]),
    # bulk delta f of kind 3 (1/sqrt(E/T) form) + shear, CPU pin of the oracle only so far
    "cell_bulk3": dict(music="9", param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=14,
                       cell="testViscousOneFluidCell2.dat", scale=0.002,
                       over=["include_deltaf_shear=1", "include_deltaf_bulk=1", "bulk_deltaf_kind=3"]),
    # kind 0 on a surface with bulk pressure (the one-cell fixtures have Pi = 0)
    "surf3d_bulk0": dict(music=None, param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=15,
                         gen=dict(ncell=300, seed=2026, eos=9),
                         over=["include_deltaf_shear=1", "include_deltaf_bulk=1", "bulk_deltaf_kind=0"]),
    "surf3d_bulk1": dict(music=None, param="iSS_parameters_CEdeltaf.dat", nev=10000, seed=12,
                         gen=dict(ncell=300, seed=2025, eos=14, rhob=1, diffusion=1, binary=1),
                         over=["include_deltaf_shear=1", "include_deltaf_bulk=1", "bulk_deltaf_kind=1",
                               "include_deltaf_diffusion=1", "restrict_deltaf=1"]),
}


def golden_legacy(only=None):
    for name, spec in LEGACY.items():
        if only and name not in only:
            continue
        d = workdir()
        case = os.path.join(d, "case")
        os.makedirs(case)
        cells = synthetic.make_case(case, **spec["gen"])
        over = LEGACY_COMMON + spec["over"]
        run([os.path.join(REF, "ref_driver"), "legacy", os.path.join(FIX, spec["param"]), "case",
             "surface.dat", os.path.join(d, "out")] + over, d, os.path.join(d, "log"))
        pre = os.path.join(d, "out")
        with open(pre + ".lab.bin", "rb") as f:
            n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
            lab = np.fromfile(f, dtype=np.float32).reshape(n, 32)
        pos = np.fromfile(pre + ".pos.bin", dtype=np.float32).reshape(n, 4)
        sp = np.loadtxt(pre + ".species.txt", ndmin=2)
        with open(pre + ".yields.bin", "rb") as f:
            ns, nc = np.fromfile(f, dtype=np.int64, count=2)
            y = np.fromfile(f, dtype=np.float64).reshape(int(ns), int(nc))
        mx = np.fromfile(pre + ".max.bin", dtype=np.float64).reshape(int(ns), int(nc))
        np.savez_compressed(os.path.join(HERE, "legacy_%s.npz" % name), param=spec["param"],
                            overrides=np.array(over), cells=cells,
                            gen=np.array(sorted(spec["gen"].items()), dtype=object).astype(str),
                            lab=lab, pos=pos, species=sp, yields=y, maximum=mx)
        print(name, "cells", n, "species", int(ns), "sum", y.clip(0).sum())
        shutil.rmtree(d)
    for name, spec in LEGACY_STATS.items():
        if only and name not in only:
            continue
        d = workdir()
        case = os.path.join(d, "case")
        os.makedirs(case)
        extra = {}
        if spec["music"] is not None:
            shutil.copy(os.path.join(FIX, "music_input_" + spec["music"]),
                        os.path.join(case, "music_input"))
            v = small_cell(spec["cell"], spec["scale"])
            np.savetxt(os.path.join(case, "surface.dat"), v[None, :], fmt="%.16e")
            extra["cell_line"] = v
        else:
            extra["cells"] = synthetic.make_case(case, **spec["gen"])
            extra["gen"] = np.array(sorted(spec["gen"].items()), dtype=object).astype(str)
        over = LEGACY_COMMON + spec["over"]
        run_over = ["number_of_repeated_sampling=%d" % spec["nev"], "randomSeed=%d" % spec["seed"],
                    "use_OSCAR_format=0", "use_gzip_format=0", "use_binary_format=1",
                    "perform_checks=0"] + over
        run([os.path.join(REF, "iSS.e"), os.path.join(FIX, spec["param"]), "case", "surface.dat"]
            + run_over, d, os.path.join(d, "log"))
        rec, off = obs.read_reference_bin(os.path.join(d, "particle_samples.bin"))
        s_ = obs.summarize(rec, off)
        # the lab-frame cells and the species order the reference sampled from
        run([os.path.join(REF, "ref_driver"), "legacy", os.path.join(FIX, spec["param"]), "case",
             "surface.dat", os.path.join(d, "out")] + over, d, os.path.join(d, "log2"))
        pre = os.path.join(d, "out")
        with open(pre + ".lab.bin", "rb") as f:
            n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
            extra["lab"] = np.fromfile(f, dtype=np.float32).reshape(n, 32)
        extra["pos"] = np.fromfile(pre + ".pos.bin", dtype=np.float32).reshape(n, 4)
        extra["species"] = np.loadtxt(pre + ".species.txt", ndmin=2)
        np.savez_compressed(os.path.join(HERE, "legacy_stats_%s.npz" % name), param=spec["param"],
                            overrides=np.array(over), music=str(spec["music"]),
                            **extra, **s_)
        print(name, "events", len(off) - 1, "hadrons", len(rec))
        shutil.rmtree(d)


MOMENTUM = [  # (mass, T, mu, sign)
    (0.13957, 0.15, 0.0, -1), (0.13957, 0.12, 0.05, -1), (0.49367, 0.155, 0.02, -1),
    (0.93827, 0.15, 0.3, 1), (0.93827, 0.15, -0.3, 1), (1.232, 0.14, 0.0, 1),
    (2.25, 0.16, 0.1, 1), (0.5479, 0.15, 0.0, 0), (5.0, 0.15, 0.0, -1), (8.0, 0.15, 0.0, 1),
]
P_EDGES = np.linspace(0.0, 5.0, 101)


def golden_momentum():
    d = workdir()
    hists = []
    n = 2000000
    for i, (m, T, mu, sign) in enumerate(MOMENTUM):
        out = os.path.join(d, "p%d.bin" % i)
        run([os.path.join(REF, "ref_driver"), "momentum", repr(m), repr(T), repr(mu), str(sign), str(n),
             str(100 + i), out], d, os.path.join(d, "log"))
        p = np.fromfile(out, dtype=np.float64)
        hists.append(np.histogram(p, P_EDGES)[0])
    np.savez_compressed(os.path.join(HERE, "momentum_sampler.npz"), cases=np.array(MOMENTUM),
                        edges=P_EDGES, hist=np.array(hists), n=n)
    print("momentum", np.array(hists).sum(axis=1))
    shutil.rmtree(d)


DECAY_PIDS = [113, 213, 223, 313, 2214, 3114, 221, 331, 333, 10213, 20223]


def golden_decay():
    d = workdir()
    res = {}
    n = 200000
    for pid in DECAY_PIDS:
        out = os.path.join(d, "d%d.bin" % pid)
        run([os.path.join(REF, "ref_driver"), "decay", "iSS_tables", "1", str(pid), str(n), "7", out],
            d, os.path.join(d, "log"))
        raw = np.fromfile(out, dtype=np.uint8)
        pos = 0
        nd_hist = np.zeros(8, dtype=np.int64)
        pids = {}
        esum = np.zeros(4)
        e_hist = np.zeros(50, dtype=np.int64)
        while pos < len(raw):
            nd = int(raw[pos:pos + 4].view("<i4")[0])
            pos += 4
            nd_hist[nd] += 1
            rec = raw[pos:pos + 40*nd].view(np.dtype([("pid", "<i4"), ("f", "<f4", 9)]))
            pos += 40*nd
            key = tuple(sorted(int(x) for x in rec["pid"]))
            pids[key] = pids.get(key, 0) + 1
            for r in rec:
                esum += r["f"][1:5]     # E px py pz  (iSS_Hadron: mass,E,px,py,pz,t,x,y,z)
                e_hist[min(49, int(r["f"][1]/0.05))] += 1
        res["nd_%d" % pid] = nd_hist
        keys = sorted(pids)
        res["chan_%d" % pid] = np.array([list(k) + [0]*(5 - len(k)) for k in keys], dtype=np.int64)
        res["chan_count_%d" % pid] = np.array([pids[k] for k in keys], dtype=np.int64)
        res["p4sum_%d" % pid] = esum
        res["ehist_%d" % pid] = e_hist
        print("decay", pid, nd_hist[:5], len(keys), "channels")
    np.savez_compressed(os.path.join(HERE, "decay.npz"), pids=np.array(DECAY_PIDS), n=n, **res)
    shutil.rmtree(d)


def writer_hadrons():
    """a small fixed hadron list: 5 events (one of them empty), odd values to exercise the formats"""
    rng = np.random.default_rng(31415)
    counts = [7, 0, 3, 12, 1]
    n = sum(counts)
    dt = np.dtype([("pid", "<i4"), ("mass", "<f4"), ("E", "<f4"), ("px", "<f4"), ("py", "<f4"),
                   ("pz", "<f4"), ("t", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
    h = np.zeros(n, dtype=dt)
    h["pid"] = rng.choice([211, -211, 2212, -2212, 3122, 111, 9000221, -3334], n)
    h["mass"] = rng.choice([0.13957, 0.93827, 1.11568, 0.13498], n)
    for k in ("px", "py", "pz"):
        h[k] = rng.normal(0, 0.7, n)
    h["E"] = np.sqrt(h["mass"].astype(np.float64)**2 + h["px"].astype(np.float64)**2
                     + h["py"].astype(np.float64)**2 + h["pz"].astype(np.float64)**2)
    h["t"] = rng.uniform(0.6, 15, n)
    h["x"], h["y"] = rng.uniform(-10, 10, n), rng.uniform(-10, 10, n)
    h["z"] = rng.normal(0, 5, n)
    h["t"][0] = 1e10            # stable-particle "life time" of the decay code
    h["px"][1] = 0.0
    h["pz"][2] = -1.2345678e-7
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return h, off


def golden_writers():
    d = workdir()
    case = os.path.join(d, "case")
    os.makedirs(case)
    shutil.copy(os.path.join(FIX, "music_input_9"), os.path.join(case, "music_input"))
    shutil.copy(os.path.join(FIX, "testIdealOneFluidCell2.dat"), case)
    h, off = writer_hadrons()
    with open(os.path.join(d, "hadrons.bin"), "wb") as f:
        np.int64(len(off) - 1).tofile(f)
        off.tofile(f)
        h.tofile(f)
    run([os.path.join(REF, "ref_driver"), "writers", os.path.join(FIX, "iSS_parameters_ideal.dat"),
         "case", "testIdealOneFluidCell2.dat", "hadrons.bin", "bulk_deltaf_kind=21"], d,
        os.path.join(d, "log"))
    import gzip
    np.savez_compressed(os.path.join(HERE, "writers.npz"), hadrons=h, offsets=off,
                        oscar=np.frombuffer(open(os.path.join(d, "OSCAR.DAT"), "rb").read(), dtype=np.uint8),
                        binary=np.frombuffer(open(os.path.join(d, "particle_samples.bin"), "rb").read(), dtype=np.uint8),
                        gz_text=np.frombuffer(gzip.open(os.path.join(d, "particle_samples.gz"), "rb").read(), dtype=np.uint8))
    print("writers", len(h), "hadrons")
    shutil.rmtree(d)


# smooth Cooper-Frye spectra of the legacy EmissionFunctionArray:
# name -> generator kwargs, parameter file, overrides, species (Monte-Carlo ids)
SPECTRA_SPECIES = [211, 321, 2212, -2212, 3122, 113, 2224]
SPECTRA = {
    "sp3d_shear": dict(gen=dict(ncell=150, seed=31, eos=9), param="iSS_parameters_CEdeltaf.dat",
                       over=["include_deltaf_bulk=0"]),
    "sp3d_bulk1_diff": dict(gen=dict(ncell=150, seed=32, eos=14, rhob=1, diffusion=1, binary=1),
                            param="iSS_parameters_CEdeltaf.dat",
                            over=["bulk_deltaf_kind=1", "include_deltaf_diffusion=1"]),
    "sp3d_bulk2_norestrict": dict(gen=dict(ncell=100, seed=33, eos=9), param="iSS_parameters_CEdeltaf.dat",
                                  over=["bulk_deltaf_kind=2", "restrict_deltaf=0"]),
    "sp3d_bulk3_pos": dict(gen=dict(ncell=100, seed=34, eos=9), param="iSS_parameters_CEdeltaf.dat",
                           over=["bulk_deltaf_kind=3", "use_pos_dN_only=1", "deltaf_max_ratio=0.5"]),
    "sp3d_bulk4": dict(gen=dict(ncell=100, seed=35, eos=9), param="iSS_parameters_CEdeltaf.dat",
                       over=["bulk_deltaf_kind=4"]),
    "sp3d_bulk0_quirk": dict(gen=dict(ncell=60, seed=36, eos=9), param="iSS_parameters_CEdeltaf.dat",
                             over=["bulk_deltaf_kind=0"]),
    "sp2d_ideal_boltzmann": dict(gen=dict(ncell=100, seed=37, eos=9, boost_invariant=True),
                                 param="iSS_parameters_ideal.dat",
                                 over=["hydro_mode=1", "quantum_statistics=0", "bulk_deltaf_kind=21"]),
}


def golden_spectra(only=None):
    for name, spec in SPECTRA.items():
        if only and name not in only:
            continue
        d = workdir()
        g = dict(spec["gen"])
        cells = synthetic.make_case(os.path.join(d, "case"), **g)
        out = os.path.join(d, "out")
        run([os.path.join(REF, "ref_driver"), "spectra", os.path.join(FIX, spec["param"]), "case",
             "surface.dat", out, "species=" + ",".join(str(m) for m in SPECTRA_SPECIES)] + spec["over"],
            d, os.path.join(d, "log"))
        with open(out + ".lab.bin", "rb") as f:
            n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
            lab = np.fromfile(f, dtype=np.float32).reshape(n, 32)
        sp = np.loadtxt(out + ".species.txt", ndmin=2)
        with open(out + ".dN.bin", "rb") as f:
            ns, npt, nphi = (int(v) for v in np.fromfile(f, dtype=np.int64, count=3))
            dn = np.fromfile(f, dtype=np.float64).reshape(ns, 2, npt, nphi)
        vndiff = np.array([np.loadtxt(out + ".vndiff.%d.dat" % m) for m in SPECTRA_SPECIES])
        vninte = np.array([np.loadtxt(out + ".vninte.%d.dat" % m) for m in SPECTRA_SPECIES])
        text = open(out + ".vndiff.%d.dat" % SPECTRA_SPECIES[0], "rb").read()
        np.savez_compressed(os.path.join(HERE, "spectra_%s.npz" % name), lab=lab, species=sp,
                            dN=dn[:, 0], dN_max=dn[:, 1], vndiff=vndiff, vninte=vninte,
                            vndiff_text0=np.frombuffer(text, dtype=np.uint8), cells=cells,
                            gen=np.array(sorted(g.items()), dtype=object).astype(str),
                            param=spec["param"], overrides=np.array(spec["over"]))
        print(name, lab.shape, dn.shape, "sum=%.17g" % dn[:, 0].sum())
        shutil.rmtree(d)


# whole-program golden of the spectra mode: the reference's iSS.e with MC_sampling = 0 and
# calculate_vn = 1 on a short chosen-particle list (tables folder with only that file replaced)
FLOW_CHOSEN = [211, -211, 111, 321, 2212, -2212, 2112, 3122, 113, 213, 223, 2224, 2214]
FLOW_RUNS = {
    "new": ["use_historic_flow_output_format=0", "calculate_dN_dphi=1"],
    "old": ["use_historic_flow_output_format=1"],
}


def golden_flows():
    for name, over in FLOW_RUNS.items():
        d = tempfile.mkdtemp(prefix="iss_golden_")
        os.makedirs(os.path.join(d, "iSS_tables"))
        for f in os.listdir(REF_TABLES):
            if f != "chosen_particles_SMASH.dat":
                os.symlink(os.path.join(REF_TABLES, f), os.path.join(d, "iSS_tables", f))
        # (EOS 14 with afterburner_type = 2 in the parameter file: the SMASH lists are used)
        with open(os.path.join(d, "iSS_tables", "chosen_particles_SMASH.dat"), "w") as f:
            f.write("".join("%d\n" % m for m in FLOW_CHOSEN))
        g = dict(ncell=40, seed=41, eos=14, rhob=1, diffusion=1, binary=1)
        cells = synthetic.make_case(os.path.join(d, "case"), **g)
        over = ["MC_sampling=0", "calculate_vn=1", "bulk_deltaf_kind=1", "include_deltaf_diffusion=1",
                "calculate_vn_to_order=4", "perform_checks=0"] + over
        run([os.path.join(REF, "iSS.e"), os.path.join(FIX, "iSS_parameters_CEdeltaf.dat"), "case",
             "surface.dat"] + over, d, os.path.join(d, "log"))
        files = {}
        for f in sorted(os.listdir(os.path.join(d, "case"))):
            if f.startswith(("thermal_", "dN_", "v2data")):
                files[f] = np.frombuffer(open(os.path.join(d, "case", f), "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, "flows_%s.npz" % name), cells=cells,
                            gen=np.array(sorted(g.items()), dtype=object).astype(str),
                            param="iSS_parameters_CEdeltaf.dat", overrides=np.array(over),
                            chosen=np.array(FLOW_CHOSEN), names=np.array(sorted(files)),
                            **{"file_%d" % i: files[k] for i, k in enumerate(sorted(files))})
        print("flows", name, sorted(files)[:4], "...", len(files), "files")
        shutil.rmtree(d)


if __name__ == "__main__":
    what = sys.argv[1:] or ["yields", "stats", "momentum", "decay", "writers", "spectra"]
    if "spectra" in what:
        golden_spectra([w for w in what if w in SPECTRA] or None)
        golden_flows()
    if "yields" in what:
        golden_yields([w for w in what if w in ONE_CELL or w in SYNTH] or None)
    if "momentum" in what:
        golden_momentum()
    if "decay" in what:
        golden_decay()
    if "stats" in what:
        golden_stats([w for w in what if w in STATS] or None)
    if "writers" in what:
        golden_writers()
    if "legacy" in what:
        golden_legacy([w for w in what if w in LEGACY or w in LEGACY_STATS] or None)
