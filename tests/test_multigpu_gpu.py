"""The one collective of the path behind the C ABI, and the sharding identities on real GPUs
(SURVEY.md section 8(e)).  Tests that need more than one GPU skip on a single-GPU box; the
world-size-2 logic is also covered on CPU with gloo (tests/test_sharding_cpu.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def gpu_count():
    import torch
    return torch.cuda.device_count()


def test_qa_allreduce_without_communicator_is_local(built, tmp_path):
    """iss_cuda_histograms_allreduce(h, NULL) on a handle that never joined a job: a no-op, the QA
    block stays this rank's (single-GPU hosts need no NCCL)."""
    capi = built
    g = cases.load("s3d_ce")
    param, surf, over = cases.materialise(g, str(tmp_path))
    s = capi.Sampler(str(tmp_path), param, surf, **over)
    try:
        assert s.read_in_FO_surface() == 0
        s.set_random_seed(1)
        assert s.prepare_sampler() == 0
        e = s.engine()
        e.compute_yields()
        e.sample(3, 0, 200)
        qa = e.histograms([211, 2212]).copy()
        assert capi.cuda_lib().iss_cuda_histograms_allreduce(e.h, None) == 0
        qa2 = np.zeros_like(qa)
        e.check(e.L.iss_cuda_qa_fetch(e.h, capi._ptr(qa2)), "qa_fetch")
        assert np.array_equal(qa, qa2)
    finally:
        s.close()


def test_facade_reduces_checks_over_two_ranks(built, tmp_path):
    """Two processes, one GPU each, events [0, n) and [n, 2n) of the same seed, parameter
    reduce_checks_over_ranks = 1: both ranks hold the QA block of all 2n events, perform_checks
    writes the job-wide files on every rank, and the block equals the one a single process
    accumulates over the 2n events (same hadrons: the streams are keyed by the event index)."""
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    capi = built
    name, nev = "viscous2", 40
    idfile = str(tmp_path/"nccl_id")
    procs, logs = [], []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), ISS_NCCL_ID_FILE=idfile)
        logs.append(open(tmp_path/("rank%d.log" % r), "w"))      # files: a full pipe would stall a rank
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "multigpu_worker.py"), name,
                                       str(nev), str(tmp_path/("rank%d.npz" % r))], env=env,
                                      stdout=logs[-1], stderr=subprocess.STDOUT))
    try:
        for p in procs:
            p.wait(timeout=240)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        for f in logs:
            f.close()
    for r, p in enumerate(procs):
        assert p.returncode == 0, open(tmp_path/("rank%d.log" % r)).read()[-3000:]
    r0, r1 = (np.load(tmp_path/("rank%d.npz" % r)) for r in range(2))
    assert np.array_equal(r0["qa"], r1["qa"])
    assert r0["qa"][0] == 2*nev and r0["n_events"] == nev
    assert r0["qa"][25] == r0["n_hadrons"] + r1["n_hadrons"]
    assert np.array_equal(r0["tmunu"], r1["tmunu"]) and np.array_equal(r0["spectra"], r1["spectra"])
    # single process over the same 2n events
    g = cases.load(name)
    d = tmp_path/"single"
    param, surf, over = cases.materialise(g, str(d))
    over.update(number_of_repeated_sampling=2*nev, perform_checks=1, use_OSCAR_format=0)
    s = capi.Sampler(str(d), param, surf, table_path=cases.tables_for(g), **over)
    try:
        s.read_in_FO_surface()
        s.set_random_seed(77)
        s.generate_samples()
        qa = s.qa_block().copy()
        cwd = os.getcwd()
        os.chdir(d)
        s.perform_checks()
        os.chdir(cwd)
        tm = np.loadtxt(d/"checkReconstructedTmunu.dat")
    finally:
        s.close()
    counts = np.r_[0, 25:29]
    assert np.array_equal(qa[counts], r0["qa"][counts])
    assert np.allclose(qa, r0["qa"], rtol=1e-10, atol=1e-9)
    assert np.allclose(tm, r0["tmunu"], rtol=1e-6, atol=1e-9)


def test_surface_chunks_over_nccl_reproduce_the_whole_surface(built):
    """tools/chunk_probe.py under torchrun on all GPUs of the box (up to 8): every rank holds 1/N of
    the cells; species totals are identical on all ranks and equal to the whole-surface run, the
    ranks' hadron counts, tries and additive QA entries (NCCL all-reduce) add up to it."""
    n = min(gpu_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(REPO, "tools", "chunk_probe.py"), "--cells", "300000", "--events", "100",
           "--steps", "1"]
    r = subprocess.run(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=420)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    line = json.loads([ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == n
    assert all(line["checks"].values()), line["checks"]
    assert line["hadrons_all_ranks"] > 1e6
