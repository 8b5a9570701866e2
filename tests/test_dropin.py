"""Source-level drop-in check: a host that uses only the reference's public class API compiles and
links against this repository's headers and libiSS.so (CPU), and runs on the GPU."""
import os
import subprocess

import pytest

import cases
from iss_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))


def build_host(tmp_path):
    libdir = os.path.dirname(capi.host_lib_path())
    exe = str(tmp_path/"dropin_host")
    cmd = ["g++", "-O1", "-std=c++17", os.path.join(HERE, "dropin_host.cpp"),
           "-I" + os.path.join(libdir, "host"), "-I" + os.path.join(capi.REPO, "include"),
           "-L" + libdir, "-liSS", "-liss_cuda", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()
    return exe


def test_reference_style_host_compiles_and_links(built, tmp_path):
    build_host(tmp_path)


@pytest.mark.gpu
def test_reference_style_host_runs(built, tmp_path):
    exe = build_host(tmp_path)
    g = cases.load("ideal2")
    case = tmp_path/"case"
    param, surf, over = cases.materialise(g, str(case))
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    args = [exe, "case", param, surf, "50", "9"] + ["%s=%r" % kv for kv in over.items()]
    r = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-1500:]
    line = [l for l in out.splitlines() if l.startswith("DROPIN")][0].split()
    nev, hadrons = int(line[2]), int(line[4])
    dN = float(g["yields"].sum())
    assert nev == 50
    assert abs(hadrons/nev - dN) < 6*(dN/nev)**0.5
