"""Surface ingest on the device (iss_cuda_ingest_music_binary, the default for binary surfaces behind
class iSS) against (a) dumps of the unmodified reference and (b) the host reader of this repo on a
larger surface.  The kernels restate the host arithmetic expression by expression without
multiply-add contraction, so the float32 records are required to be bit-identical; only cosh/sinh of
the space-time rapidity come from a different math library (rounded to float: a difference needs the
double result within ~2 ulp of a float rounding boundary, ~1e-8 per value), so a handful of 1-ulp
differences per million cells is tolerated."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from iss_b200 import capi, synthetic

pytestmark = pytest.mark.gpu

BINARY_CASES = [n for n in cases.SYNTH if "binary" in dict(cases.load(n)["gen"])
                and int(dict(cases.load(n)["gen"])["binary"]) == 1]


def dump(tmp_path, param, surf, over, mode):
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iss_host_dump")
    args = [exe, param, "case", surf, "out_" + mode] + ["%s=%r" % kv for kv in over.items()]
    env = {k: v for k, v in os.environ.items() if k != "ISS_INGEST"}
    if mode == "host":
        env["ISS_INGEST"] = "host"
    r = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    with open(tmp_path/("out_%s.lrf.bin" % mode), "rb") as f:
        n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        lrf = np.fromfile(f, dtype=np.float32).reshape(n, 28)
    tmunu = [l for l in r.stdout.decode().splitlines() if "] = " in l and "GeV/fm^3" in l]
    return lrf, tmunu


def test_binary_cases_exist():
    assert len(BINARY_CASES) >= 2


@pytest.mark.parametrize("name", BINARY_CASES)
def test_device_ingest_matches_reference_dump(name, tmp_path):
    g = cases.load(name)
    param, surf, over = cases.materialise(g, str(tmp_path/"case"))
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    lrf, _ = dump(tmp_path, param, surf, over, "device")
    assert lrf.shape == g["lrf"].shape
    assert np.array_equal(lrf.view(np.uint32), g["lrf"].view(np.uint32))


@pytest.mark.parametrize("kw", [dict(eos=14, rhob=1, diffusion=1), dict(eos=9),
                                dict(eos=91, boost_invariant=True)])
def test_device_ingest_equals_host_ingest_large(kw, tmp_path):
    """3 x 10^5 cells: same kept cells in the same order, float32 records bit-identical up to a few
    1-ulp differences, same T^{mu nu} log lines (sequential float sums of device tensors)"""
    n = 300000
    synthetic.make_case(str(tmp_path/"case"), ncell=n, seed=99, binary=1, **kw)
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    param = os.path.join(cases.FIX, "iSS_parameters_CEdeltaf.dat")
    over = dict(hydro_mode=1) if kw.get("boost_invariant") else {}
    dev, tm_dev = dump(tmp_path, param, "surface.dat", over, "device")
    host, tm_host = dump(tmp_path, param, "surface.dat", over, "host")
    assert dev.shape == host.shape and dev.shape[0] > 0.9*n
    diff = dev.view(np.uint32) != host.view(np.uint32)
    assert diff.sum() <= 8, diff.sum()
    if diff.any():
        a, b = dev[diff].astype(np.float64), host[diff].astype(np.float64)
        assert np.all(np.abs(a - b) <= 1.3e-7*np.abs(b))        # one float ulp
    else:
        assert tm_dev == tm_host and len(tm_dev) == 16


def test_device_ingest_filters_and_status():
    """T <= 0.01 GeV cells and u.dsigma < 0 cells through the C ABI (no EOS regulation)"""
    raw = synthetic.generate_cells(2000, seed=5).astype(np.float32)
    raw[::50, 13] = 0.5*0.01/0.197327053            # cold cells: dropped by the T filter
    raw[7::50, 4] = -5.0                             # dsigma_tau strongly negative: u.dsigma < 0
    e = capi.Engine(0)
    try:
        lrf, tm, st = e.ingest_music_binary(raw)
    finally:
        e.close()
    assert (st[::50] & 1).all() and not (st[1::50] & 1).any()
    assert (st[7::50] & 4).all()
    assert len(tm) == len(raw) - len(raw[::50])
    assert len(lrf) == int(((st & 5) == 0).sum())
    # kept cells keep the file order: tau, x, y are copied through
    kept = raw[(st & 5) == 0]
    assert np.array_equal(lrf[:, :3], kept[:, :3])
