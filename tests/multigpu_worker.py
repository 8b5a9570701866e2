"""Worker of tests/test_multigpu_gpu.py: one rank of a one-process-per-GPU job driving the drop-in
facade (class iSS) with the QA block reduced over the ranks through the C ABI
(iss_cuda_nccl_init / iss_cuda_histograms_allreduce).  argv: case name, events per rank, out.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import cases  # noqa: E402
from iss_b200 import capi  # noqa: E402


def main():
    name, nev, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rank = int(os.environ["RANK"])
    work = os.path.join(os.path.dirname(out), "work_r%d" % rank)
    g = cases.load(name)
    param, surf, over = cases.materialise(g, work)
    over.update(number_of_repeated_sampling=nev, first_event_index=rank*nev, perform_checks=1,
                use_OSCAR_format=0, use_gzip_format=0, use_binary_format=0,
                reduce_checks_over_ranks=1)
    s = capi.Sampler(work, param, surf, table_path=cases.tables_for(g), **over)
    assert s.read_in_FO_surface() == 0
    s.set_random_seed(77)
    assert s.generate_samples() == 0
    h, off = s.hadrons()
    cwd = os.getcwd()
    os.chdir(work)
    s.perform_checks()
    os.chdir(cwd)
    np.savez(out, qa=s.qa_block(), n_hadrons=len(h), n_events=len(off) - 1,
             tmunu=np.loadtxt(os.path.join(work, "checkReconstructedTmunu.dat")),
             spectra=np.loadtxt(os.path.join(work, "check_211_spectra.dat")))
    s.close()


if __name__ == "__main__":
    main()
