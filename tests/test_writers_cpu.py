"""Row O of SURVEY.md section 8(a): OSCAR / binary / gz sample files are byte-compatible with the
files the reference's own writers (FSSW::combine_samples_to_OSCAR/_gzip_file/_binary_file,
FSSW.cpp:365-561) produce for the same hadron list (golden: tests/golden/writers.npz, written by
the compiled reference through oracle/ref_driver.cpp `writers`).  CPU only."""
import gzip
import os

import numpy as np

import cases
from iss_b200 import capi


def test_sample_files_are_byte_identical_to_the_reference(built, tmp_path):
    g = np.load(os.path.join(cases.GOLDEN, "writers.npz"))
    h = g["hadrons"]
    off = g["offsets"]
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for fmt in ("oscar", "gzip", "binary"):
            capi.write_samples(fmt, h, off)
        assert open("OSCAR.DAT", "rb").read() == g["oscar"].tobytes()
        assert open("particle_samples.bin", "rb").read() == g["binary"].tobytes()
        assert gzip.open("particle_samples.gz", "rb").read() == g["gz_text"].tobytes()
    finally:
        os.chdir(cwd)


def test_oscar_large_event_is_ordered(built, tmp_path):
    """the multi-threaded OSCAR formatter keeps the particle order (events above the threading
    threshold)"""
    n = 20000
    h = np.zeros(n, dtype=capi.HADRON_DTYPE)
    h["pid"] = 211
    h["px"] = np.arange(n, dtype=np.float32)
    h["E"] = 1.0
    off = np.array([0, n], dtype=np.int64)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        capi.write_samples("oscar", h, off)
        lines = open("OSCAR.DAT").read().splitlines()
        body = lines[-n:]
        idx = np.array([int(l.split()[0]) for l in body])
        px = np.array([float(l.split()[2]) for l in body])
        assert np.array_equal(idx, np.arange(1, n + 1))
        assert np.array_equal(px, np.arange(n))
    finally:
        os.chdir(cwd)
