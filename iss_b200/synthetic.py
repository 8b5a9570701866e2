"""Synthetic freeze-out hypersurfaces in MUSIC format.

Column order and units are those of the reference reader
(reference src/readindata.cpp:626-765): 34 numbers per cell,
tau x y eta | da0..3 | u0..3 | e T muB muS muQ (e+P)/T | pi00 01 02 03 11 12 13 22 23 33 |
Pi | rhoB | q0..3, with e, T, mu, pi, Pi in fm^-n units (multiplied by hbarc on read).
Binary files hold all 34 float32 per cell; text files hold 28 + [Pi] + [rhoB] + [q0..3]
columns depending on the music_input flags.
"""
import os

import numpy as np

HBARC = 0.197327053


def generate_cells(ncell, seed, boost_invariant=False, rhob=False, diffusion=False):
    """Returns a float64 array [ncell, 34] in file units."""
    rng = np.random.default_rng(seed)
    c = np.zeros((ncell, 34))
    c[:, 0] = rng.uniform(0.6, 12.0, ncell)
    c[:, 1] = rng.uniform(-10, 10, ncell)
    c[:, 2] = rng.uniform(-10, 10, ncell)
    eta = rng.uniform(-4, 4, ncell)
    c[:, 3] = 0.0 if boost_invariant else eta
    c[:, 4] = rng.uniform(0, 0.05, ncell)
    c[:, 5:8] = rng.normal(0, 0.01, (ncell, 3))
    u = np.empty((ncell, 3))
    u[:, 0:2] = rng.normal(0, 0.4, (ncell, 2))
    u[:, 2] = rng.normal(0, 0.2, ncell)
    if boost_invariant:
        c[:, 7] = 0.0
        u[:, 2] = 0.0
    c[:, 9:12] = u
    c[:, 8] = np.sqrt(1.0 + (u**2).sum(axis=1))
    e = 1.5228525*rng.uniform(0.6, 1.2, ncell)
    T = rng.uniform(0.140, 0.160, ncell)/HBARC
    c[:, 12] = e
    c[:, 13] = T
    c[:, 17] = 1.15*e/T
    c[:, 18:28] = rng.normal(0, 0.01, (ncell, 10))
    c[:, 28] = -np.abs(rng.normal(0, 0.003, ncell))
    rb = rng.uniform(0, 0.08, ncell)
    q = rng.normal(0, 0.002, (ncell, 4))
    if rhob:
        c[:, 29] = rb
    if diffusion:
        c[:, 30:34] = q
    return c


def write_music_input(folder, eos, bulk=1, rhob=0, diffusion=0, binary=0):
    with open(os.path.join(folder, "music_input"), "w") as f:
        f.write("EOS_to_use  %d\n" % eos)
        f.write("Include_Bulk_Visc_Yes_1_No_0  %d\n" % bulk)
        f.write("Include_Rhob_Yes_1_No_0  %d\n" % rhob)
        f.write("turn_on_baryon_diffusion  %d\n" % diffusion)
        f.write("freeze_surface_in_binary  %d\n" % binary)
        f.write("EndOfData\n")


def write_surface(path, cells, binary, bulk=1, rhob=0, diffusion=0):
    if binary:
        cells.astype(np.float32).tofile(path)
        return
    cols = list(range(28))
    if bulk:
        cols.append(28)
    if rhob:
        cols.append(29)
    if diffusion:
        cols += [30, 31, 32, 33]
    np.savetxt(path, cells[:, cols], fmt="%.10e")


def make_case(folder, ncell, seed, eos, boost_invariant=False, rhob=0, diffusion=0, binary=0,
              bulk=1, surface_name="surface.dat"):
    """Writes <folder>/music_input and <folder>/<surface_name>; returns the cell array."""
    os.makedirs(folder, exist_ok=True)
    cells = generate_cells(ncell, seed, boost_invariant, bool(rhob), bool(diffusion))
    write_music_input(folder, eos, bulk, rhob, diffusion, binary)
    write_surface(os.path.join(folder, surface_name), cells, binary, bulk, rhob, diffusion)
    return cells
