"""Ad-hoc probe (not a test): host-side phase times of iSS::generate_samples() on the bench surface."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ISS_PROFILE"] = "1"
import bench
from iss_b200 import capi
work = tempfile.mkdtemp()
bench.make_case(work, 1000000)
fd = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1)
s = capi.Sampler(work, bench.PARAM, "surface.dat", **dict(bench.OVERRIDES, number_of_repeated_sampling=1000))
s.read_in_FO_surface(); s.set_random_seed(1)
for i in range(int(os.environ.get("E2E_CALLS", "3"))):
    t0 = time.perf_counter(); s.generate_samples(); t1 = time.perf_counter()
    sys.stderr.write("=== generate_samples %d: %.1f ms\n" % (i, 1e3*(t1 - t0)))
