"""Ad-hoc probe (not a test): decay kernel throughput on a boost-invariant 1e5-cell surface."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from iss_b200 import capi, synthetic
work = tempfile.mkdtemp()
synthetic.make_case(work, ncell=100000, seed=12345, eos=9, boost_invariant=True)
fd = os.dup(1); os.dup2(2, 1)
over = dict(bench.OVERRIDES, hydro_mode=1, include_deltaf_diffusion=0, perform_decays=1, number_of_repeated_sampling=20, y_LB=-2.0, y_RB=2.0)
s = capi.Sampler(work, bench.PARAM, "surface.dat", **over)
s.read_in_FO_surface(); s.set_random_seed(3); s.prepare_sampler()
e = s.engine(); e.compute_yields()
nev = int(os.environ.get("NEV", "400"))
for rep in range(3):
    e.timing(enable=True, reset=True)
    c = e.sample(3, rep*nev, (rep+1)*nev)
    nprim = c.n_hadrons
    c2 = e.decay(3)
    ms, n = e.timing(enable=False)
    os.dup2(fd, 1)
    print("rep %d primaries %d finals %d  setup %.3f propose %.3f decay %.3f ms (%d launches) -> %.3e decayed primaries/s" % (
        rep, nprim, c2.n_hadrons, ms["setup"], ms["sample"], ms["decay"], n["decay"], nprim/(ms["decay"]*1e-3)))
    os.dup2(2, 1)
