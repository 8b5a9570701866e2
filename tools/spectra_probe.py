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
