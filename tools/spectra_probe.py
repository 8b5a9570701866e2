"""Ad-hoc probe (not a test): throughput of the smooth-spectra kernel (iss_cuda_spectra) on a
synthetic lab-frame surface.  PROBE_CELLS cells x PROBE_SPECIES species x 15 x 48 x 81 points."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import spectra_cases as sc
from iss_b200 import capi

ncell = int(os.environ.get("PROBE_CELLS", "100000"))
ns = int(os.environ.get("PROBE_SPECIES", "32"))
g = sc.load("sp3d_bulk1_diff")
rng = np.random.default_rng(1)
lab = g["lab"][rng.integers(0, len(g["lab"]), size=ncell)].copy()
lab[:, 9] *= rng.uniform(0.97, 1.03, size=ncell).astype(np.float32)      # spread T a little
base = sc.species_of(g)
sp = np.zeros(ns, dtype=capi.SPECIES_DTYPE)
for k in range(ns):
    b = base[k % len(base)]
    sp[k]["pid"], sp[k]["mass"], sp[k]["gspin"] = b["pid"], b["mass"]*(1 + 0.01*(k//len(base))), int(b["gspin"])
    sp[k]["baryon"], sp[k]["strange"], sp[k]["charge"], sp[k]["sign"] = b["baryon"], b["strange"], b["charge"], int(b["sign"])
pT, phi, eta = sc.bin_tables()
e = capi.Engine(0)
e.upload_table(capi.TABLE_KAPPA_B, sc.kappa_table(), 150, 100, [0.05, 0.001, 0.0, 0.007892])
e.upload_surface_lab(lab)
peak = e.fp64_peak()
for name, kw in (("shear only", dict(include_deltaf_bulk=0, include_deltaf_diffusion=0)),
                 ("shear+bulk1+diff", dict(include_deltaf_bulk=1, bulk_deltaf_kind=1, include_deltaf_diffusion=1)),
                 ("ideal, no restrict", dict(include_deltaf_shear=0, include_deltaf_bulk=0, restrict_deltaf=0))):
    e.spectra(sp[:2], pT[:, 0], phi[:, 0], eta[:, 0], eta[:, 1], **kw)
    t0 = time.perf_counter()
    dN, _ = e.spectra(sp, pT[:, 0], phi[:, 0], eta[:, 0], eta[:, 1], **kw)
    wall = time.perf_counter() - t0
    n, ms = e.spectra_stats()
    print("%-20s cells=%d species=%d evals=%.3e kernel %.1f ms (wall %.1f ms) -> %.3e evals/s; "
          "FP64 DFMA peak %.1f TFLOP/s" % (name, ncell, ns, n, ms, 1e3*wall, n/(ms*1e-3), peak))

# CPU reference beside it: the unmodified reference (oracle/_ref/ref_driver spectra, one core) on a
# 300-cell sample of the same kind of surface, 7 species, bulk kind 1 + diffusion
import json, shutil, subprocess, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
drv = os.path.join(REPO, "oracle", "_ref", "ref_driver")
if os.path.exists(drv) and os.environ.get("PROBE_CPU", "1") == "1":
    from iss_b200 import synthetic
    d = tempfile.mkdtemp()
    os.symlink(os.path.join(REPO, "iSS_tables"), os.path.join(d, "iSS_tables"))
    synthetic.make_case(os.path.join(d, "case"), ncell=300, seed=32, eos=14, rhob=1, diffusion=1, binary=1)
    t0 = time.perf_counter()
    subprocess.run([drv, "spectra", os.path.join(REPO, "tests", "fixtures", "iSS_parameters_CEdeltaf.dat"),
                    "case", "surface.dat", os.path.join(d, "out"), "species=211,321,2212,-2212,3122,113,2224",
                    "bulk_deltaf_kind=1", "include_deltaf_diffusion=1"], cwd=d, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    wall = time.perf_counter() - t0
    log_evals = 300*81*15*48*7
    print(json.dumps({"cpu_reference": "ref_driver spectra (EmissionFunctionArray::calculate_dN_pTdpTdphidy), 1 core, "
                      "300 cells x 7 species, bulk kind 1 + diffusion", "evals": log_evals, "wall_s": round(wall, 3),
                      "evals_per_s_per_core": log_evals/wall, "note": "wall includes reading the tables (~0.1 s)"}))
    shutil.rmtree(d)
