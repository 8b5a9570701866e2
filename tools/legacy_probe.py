"""Ad-hoc probe (not a test): throughput of the legacy conventional sampler (MC_sampling = 2,
EmissionFunctionArray) on a synthetic 3+1D surface of 1e5 cells with shear delta f, UrQMD list,
next to the reference's own MC_sampling = 2 run on one host core (oracle/_ref/iSS.e, when it is in
the snapshot) on a 1/50 sample of the same surface."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from iss_b200 import capi, synthetic
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCELL = int(os.environ.get("NCELL", "100000"))
NEV = int(os.environ.get("NEV", "200"))
OVER = dict(MC_sampling=2, include_deltaf_shear=1, include_deltaf_bulk=0, bulk_deltaf_kind=1,
            include_deltaf_diffusion=0, perform_decays=0, perform_checks=0, use_OSCAR_format=0,
            use_gzip_format=0, use_binary_format=0, store_samples_in_memory=1,
            output_samples_into_files=0)
work = tempfile.mkdtemp()
synthetic.make_case(work, ncell=NCELL, seed=12345, eos=9)
fd = os.dup(1); os.dup2(2, 1)
s = capi.Sampler(work, bench.PARAM, "surface.dat", number_of_repeated_sampling=NEV, **OVER)
s.read_in_FO_surface(); s.set_random_seed(3); s.prepare_sampler()
e = s.engine()
for rep in range(3):
    e.timing(enable=True, reset=True)
    e.L.iss_cuda_legacy_compute_yields(e.h, None, None, None)
    c = e.sample(3, rep*NEV, (rep + 1)*NEV)
    ms, n = e.timing(enable=False)
    os.dup2(fd, 1)
    print("rep %d cells %d species %d events %d hadrons %d tries/hadron %.1f  yields %.3f scan %.3f "
          "sample %.3f ms -> %.3e hadrons/s, %.3e tries/s, %.3e cell x species yields/s" % (
              rep, NCELL, e.nspecies if hasattr(e, "nspecies") else -1, NEV, c.n_hadrons,
              c.n_tries/max(1, c.n_hadrons), ms["yields"], ms["scan"], ms["sample"],
              c.n_hadrons/(ms["sample"]*1e-3), c.n_tries/(ms["sample"]*1e-3),
              NCELL*len(s.species())/(ms["yields"]*1e-3)))
    os.dup2(2, 1)
s.close()
ref = os.path.join(REPO, "oracle", "_ref", "iSS.e")
if os.path.exists(ref):
    w2 = tempfile.mkdtemp()
    os.makedirs(os.path.join(w2, "case"))
    synthetic.make_case(os.path.join(w2, "case"), ncell=NCELL//50, seed=12345, eos=9)
    os.symlink(os.path.join(REPO, "iSS_tables"), os.path.join(w2, "iSS_tables"))
    nev_ref = 20
    args = [ref, bench.PARAM, "case", "surface.dat", "number_of_repeated_sampling=%d" % nev_ref,
            "randomSeed=1"] + ["%s=%g" % kv for kv in OVER.items()]
    t0 = time.time()
    out = subprocess.run(args, cwd=w2, capture_output=True, text=True).stdout
    wall = time.time() - t0
    nh = 0
    for line in out.splitlines():
        if "dN=" in line:
            try:
                nh += float(line.split("dN=")[1].split("...")[0])*nev_ref
            except ValueError:
                pass
    os.dup2(fd, 1)
    print("reference MC_sampling=2, one core, %d cells, %d events: ~%.0f hadrons in %.2f s wall -> %.3e hadrons/s"
          % (NCELL//50, nev_ref, nh, wall, nh/max(wall, 1e-9)))
