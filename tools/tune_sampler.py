"""Ad-hoc tuning probe (not a test): sampler/yields kernel time on the C4 bench surface."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from iss_b200 import capi

ncell = int(os.environ.get("TUNE_CELLS", "1000000"))
E = int(os.environ.get("TUNE_EVENTS", "1000"))
work = tempfile.mkdtemp()
bench.make_case(work, ncell)
fd = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1)
s = capi.Sampler(work, bench.PARAM, "surface.dat", **dict(bench.OVERRIDES, number_of_repeated_sampling=E))
s.read_in_FO_surface(); s.set_random_seed(1); s.prepare_sampler()
e = s.engine()
e.compute_yields(); e.sample(1, 0, E)
e.timing(enable=True, reset=True)
n = 0
for k in range(3):
    e.compute_yields()
    c = e.sample(1, (k+1)*E, (k+2)*E); n += c.n_hadrons
    e.L.iss_cuda_histograms(e.h, capi._ptr(np.asarray([211, 2212], dtype=np.int32)), 2, 0)
ms, cnt = e.timing(enable=False)
os.dup2(fd, 1)
print("cells=%d ev=%d hadrons/step=%d tries/hadron=%.3f  ms/step: %s  sampler %.3e hadrons/s" % (
    ncell, E, n//3, c.n_tries/c.n_hadrons,
    {k: round(v/3, 3) for k, v in ms.items()}, n/(ms["sample"]*1e-3)))
