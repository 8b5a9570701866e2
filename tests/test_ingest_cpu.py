"""CPU tests of the host ingest: music_input / surface parsing, HRG regulation, particle table,
local-rest-frame transform and species order, against dumps of the unmodified reference
(tests/golden/yields_*.npz: `lrf`, `species`).  Bit-exact float32 records are required because
Sigma_LRF and T feed every yield at 1e-6 (reference iSS.cpp:170-293, readindata.cpp:626-842)."""
import os
import subprocess

import numpy as np
import pytest

import cases
from iss_b200 import capi


@pytest.mark.parametrize("name", cases.ONE_CELL + cases.SYNTH)
def test_ingest_bit_exact(name, built, tmp_path):
    g = cases.load(name)
    case = tmp_path/"case"
    param, surf, over = cases.materialise(g, str(case))
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iss_host_dump")
    args = [exe, param, "case", surf, "out"] + ["%s=%r" % kv for kv in over.items()]
    # the HOST reader is what this CPU test checks; binary surfaces go through the device by
    # default (tests/test_ingest_gpu.py) and never fall back silently
    r = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=dict(os.environ, ISS_INGEST="host"))
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    with open(tmp_path/"out.lrf.bin", "rb") as f:
        n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        lrf = np.fromfile(f, dtype=np.float32).reshape(n, 28)
    assert lrf.shape == g["lrf"].shape
    assert np.array_equal(lrf.view(np.uint32), g["lrf"].view(np.uint32))
    sp = np.loadtxt(tmp_path/"out.species.txt", ndmin=2)
    assert np.array_equal(sp, g["species"])


def test_libraries_export_declared_symbols(built):
    L = capi.cuda_lib()
    H = capi.host_lib()
    import re
    inc = os.path.join(capi.REPO, "include")
    declared = set(re.findall(r"\b(iss_cuda_\w+)\s*\(", open(os.path.join(inc, "iss_cuda.h")).read()))
    assert declared == set(capi.CUDA_SYMBOLS)
    for sname in declared:
        assert hasattr(L, sname), sname
    declared_h = set(re.findall(r"\b(iss_host_\w+)\s*\(", open(os.path.join(inc, "iss_host.h")).read()))
    assert declared_h == set(capi.HOST_SYMBOLS)
    for sname in declared_h:
        assert hasattr(H, sname), sname


def test_no_gpu_fails_loudly(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.IssError):
        capi.Engine()
    # a binary surface is ingested on the device: without one the facade exits with a message
    # instead of quietly using the host reader
    g = cases.load("s3d_ce_diff")
    param, surf, over = cases.materialise(g, str(tmp_path/"case"))
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iss_host_dump")
    env = {k: v for k, v in os.environ.items() if k != "ISS_INGEST"}
    r = subprocess.run([exe, param, "case", surf, "out"] + ["%s=%r" % kv for kv in over.items()],
                       cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    assert r.returncode != 0
    assert b"no usable CUDA device" in r.stdout


def _dump_lrf(tmp_path, param, surf, over, threads, out):
    exe = os.path.join(os.path.dirname(capi.host_lib_path()), "iss_host_dump")
    args = [exe, param, "case", surf, out] + ["%s=%r" % kv for kv in over.items()]
    r = subprocess.run(args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=dict(os.environ, ISS_INGEST="host", ISS_HOST_THREADS=str(threads)))
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    with open(tmp_path/(out + ".lrf.bin"), "rb") as f:
        n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        return np.fromfile(f, dtype=np.float32).reshape(n, 28), r.stdout.decode()


@pytest.mark.parametrize("boost_invariant", [False, True])
def test_text_surface_parallel_parse(boost_invariant, built, tmp_path):
    """Text surfaces are cut at newlines and parsed on several threads with std::from_chars: the
    records must not depend on the number of threads, a file whose cells straddle lines (allowed
    by the reference's stream reader in 3+1D, readindata.cpp:692-749) must give the same records
    through the single-piece fallback, and cold cells are reported and dropped in file order."""
    import bench
    from iss_b200 import synthetic
    case = tmp_path/"case"
    synthetic.make_case(str(case), ncell=30000, seed=99, eos=14, rhob=1, diffusion=1, binary=0,
                        boost_invariant=boost_invariant)
    os.symlink(capi.TABLES, tmp_path/"iSS_tables")
    over = dict(bench.OVERRIDES, hydro_mode=1 if boost_invariant else 2)
    text = open(case/"surface.dat").read()
    lines = text.splitlines()
    assert len(lines) == 30000 and os.path.getsize(case/"surface.dat") > 8*(1 << 20)
    # two cells below the temperature cut, far apart (different pieces)
    for k in (5, 29990):
        cols = lines[k].split()
        cols[13] = "1.0e-03"            # T in fm^-1: 2e-4 GeV
        lines[k] = " ".join(cols)
    open(case/"surface.dat", "w").write("\n".join(lines) + "\n")
    one, log1 = _dump_lrf(tmp_path, bench.PARAM, "surface.dat", over, 1, "t1")
    many, log8 = _dump_lrf(tmp_path, bench.PARAM, "surface.dat", over, 8, "t8")
    assert len(one) > 20000
    assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    assert log1.count("Discard surf elem") == 2 == log8.count("Discard surf elem")
    # a last cell that is not followed by a newline is dropped by both readers of the reference: the
    # 3+1D stream has eofbit set after extracting the last number (`if (!surfdat.eof())`,
    # readindata.cpp:752; checked against oracle/_ref: 49 -> 48 cells), the boost-invariant reader
    # keeps a line only when the stream is not at eof after it (readindata.cpp:531-541)
    open(case/"surface.dat", "w").write("\n".join(lines))
    cut, _ = _dump_lrf(tmp_path, bench.PARAM, "surface.dat", over, 8, "t8b")
    assert len(one) - 1 <= len(cut) + 0 <= len(one)     # (the dropped cell may have failed u.dsigma >= 0 anyway)
    assert np.array_equal(cut.view(np.uint32), one[:len(cut)].view(np.uint32))
    full_last, _ = _dump_lrf(tmp_path, bench.PARAM, "surface.dat", over, 1, "t1b")
    assert np.array_equal(cut.view(np.uint32), full_last.view(np.uint32))
    if boost_invariant:
        return
    open(case/"surface.dat", "w").write("\n".join(lines[:-1]) + "\n")
    without, _ = _dump_lrf(tmp_path, bench.PARAM, "surface.dat", over, 8, "t8d")
    assert np.array_equal(cut.view(np.uint32), without.view(np.uint32))
    # same numbers, seven per line: cells straddle lines
    nums = " ".join(lines).split()
    reflow = "\n".join(" ".join(nums[i:i + 7]) for i in range(0, len(nums), 7)) + "\n"
    open(case/"surface.dat", "w").write(reflow)
    odd, _ = _dump_lrf(tmp_path, bench.PARAM, "surface.dat", over, 8, "t8c")
    assert np.array_equal(odd.view(np.uint32), one.view(np.uint32))


def test_ctypes_mirrors_match_the_c_header(built, tmp_path):
    """The ctypes structures of iss_b200/capi.py (tests, bench) and of the oracle have the size and
    field offsets of the C structs in include/iss_cuda.h (compiled here with gcc)."""
    import ctypes as C
    pairs = [("iss_species", capi.Species), ("iss_options", capi.Options),
             ("iss_decay_species", capi.DecaySpecies), ("iss_decay_channel", capi.DecayChannel),
             ("iss_spectra_options", capi.SpectraOptions), ("iss_legacy_options", capi.LegacyOptions),
             ("iss_ingest_options", capi.IngestOptions), ("iss_ingest_result", capi.IngestResult),
             ("iss_counts", capi.Counts)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "iss_cuda.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('printf("\\n");')
    lines += ['printf("hadron %zu\\n", sizeof(iss_hadron));', 'return 0; }']
    src = tmp_path/"abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path/"abi"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(capi.REPO, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in out if l.strip()}
    for cname, cls in pairs:
        want = [C.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
        assert got[cname] == want, cname
    assert got["hadron"] == [40] == [capi.HADRON_DTYPE.itemsize]
