"""Ad-hoc probe (not a test): BASELINE.json configs[4] size -- 10^7-cell surface on one GPU."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from iss_b200 import capi
ncell = int(os.environ.get("C5_CELLS", "10000000"))
work = tempfile.mkdtemp()
t0 = time.time(); bench.make_case(work, ncell); t1 = time.time()
fd = os.dup(1); os.dup2(2, 1)
s = capi.Sampler(work, bench.PARAM, "surface.dat", **dict(bench.OVERRIDES, number_of_repeated_sampling=100))
s.read_in_FO_surface(); t2 = time.time()
s.set_random_seed(1); s.prepare_sampler(); t3 = time.time()
e = s.engine()
e.timing(enable=True, reset=True)
dN = e.compute_yields()
nev = 100
c = e.sample(5, 0, nev)
mult = e.multiplicities(nev); off = e.event_offsets(nev)
had = e.fetch_all()
ms, _ = e.timing(enable=False)
os.dup2(fd, 1)
E, px, py, pz, m = (had[k].astype(np.float64) for k in ("E", "px", "py", "pz", "mass"))
shell = np.abs(E*E - px*px - py*py - pz*pz - m*m).max()/np.max(E*E)
print("cells kept %d  gen %.1fs ingest %.1fs prepare(upload) %.2fs" % (e.ncell, t1-t0, t2-t1, t3-t2))
print("hadrons %d (mult sum %d, offsets %d)  dN sum %.1f per event %.1f  mass-shell dev %.2e" % (
    len(had), mult.sum(), off[-1], dN.sum(), len(had)/nev, shell))
print("kernel ms:", {k: round(v, 2) for k, v in ms.items()})
assert len(had) == mult.sum() == off[-1]
assert abs(len(had)/nev - dN.sum()) < 6*np.sqrt(dN.sum()/nev)
assert shell < 1e-4
print("C5 OK")
