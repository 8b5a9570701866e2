"""Ad-hoc GPU probe (not a test): one-cell closure + timing breakdown."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from iss_b200 import capi
import cases

def one(name, nev):
    g = cases.load(name)
    d = tempfile.mkdtemp()
    param, surf, over = cases.materialise(g, d)
    over.update(number_of_repeated_sampling=nev, perform_checks=1)
    s = capi.Sampler(d, param, surf, **over)
    s.read_in_FO_surface(); s.set_random_seed(1)
    t0 = time.time(); s.generate_samples(); t1 = time.time()
    h, off = s.hadrons()
    dN = s.species_dN()
    print(name, "events", len(off)-1, "hadrons", len(h), "per event", len(h)/(len(off)-1), "dN sum", dN.sum(), "wall", t1-t0)
    os.chdir(d); s.perform_checks()
    data = np.loadtxt("checkReconstructedTmunu.dat")
    print(name, "closure mean|diff| =", np.mean(np.abs(data[:, 2])), "(reference CI bound 1e-3)")
    e = s.engine()
    print("timing", e.timing())
    s.close()

e = capi.Engine(); print("fp64 peak TFLOP/s", e.fp64_peak()); e.close()
for n in ["ideal1", "viscous1", "viscous2", "ideal4"]:
    one(n, 200)
