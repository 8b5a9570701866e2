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


# ---- synthetic 22-moment delta-f table ---------------------------------------------------------
# The reference loads <tables>/deltaf_tables/{urqmd,smash}/NEoSBQS_22mom_deltafCoeff.dat
# (reference src/FSSW.cpp:1341-1376) for bulk_deltaf_kind = 20, the kind its own CI parameter file
# selects (tests/iSS_parameters.dat:19), but the blob is not part of the reference tree.  The
# loader fixes the format: one header line, then 200 (e) x 200 (n_B) rows of 8 numbers
# "e  n_B  c_shear  c1 .. c5" with n_B running fastest, a uniform e grid and per-e uniform n_B grids
# starting at 0.  This generator writes such a file with smooth analytic columns, so that the
# compiled reference and the engine can be run on identical input.
def table_22mom():
    """[200*200][8] float64, rounded to the 7 significant digits the file carries."""
    ne = nb = 200
    e = 0.01*(np.arange(ne) + 1.0)                    # GeV/fm^3, like the CE table's grid
    out = np.zeros((ne, nb, 8))
    for i in range(ne):
        nB_max = 0.0097*(e[i]/0.01)**0.92               # widening n_B range with e
        nB = nB_max*np.arange(nb)/(nb - 1.0)
        x = nB/nB_max
        T = 0.150*(e[i]/0.30)**0.25                     # GeV (conformal-like)
        out[i, :, 0] = e[i]
        out[i, :, 1] = nB
        out[i, :, 2] = 1.0/(2.0*T*T*1.15*e[i])*(1.0 + 0.15*x)        # shear: ~1/(2 T^2 (e+P))
        out[i, :, 3] = 14.0/(e[i] + 0.2)*(1.0 - 0.20*x*x)            # c1 (p0^2 term, with -c2)
        out[i, :, 4] = 9.0/(e[i] + 0.2)*(1.0 + 0.10*x)               # c2 (m^2 term)
        out[i, :, 5] = -2.5*x/(e[i] + 0.3)                           # c3 (baryon)
        out[i, :, 6] = 1.2*x*(1.0 - 0.5*x)/(e[i] + 0.3)              # c4 (strangeness)
        out[i, :, 7] = -0.6*x/(e[i] + 0.5)                           # c5 (charge)
    flat = out.reshape(ne*nb, 8)
    # the file is the contract: keep exactly what "%.6e" preserves
    return np.array([[float("%.6e" % v) for v in row] for row in flat])


def write_22mom_table(path):
    tab = table_22mom()
    with open(path, "w") as f:
        f.write("# e(GeV/fm^3)  rho_B(1/fm^3)  c_shear  c1  c2  c3  c4  c5   (synthetic, "
                "iss_b200/synthetic.py)\n")
        for row in tab:
            f.write("  " + "  ".join("%.6e" % v for v in row) + "\n")


def tables_with_22mom(dst, src_tables):
    """A table folder that mirrors `src_tables` through symlinks and adds the synthetic 22-moment
    table for both hadron lists.  Returns dst."""
    os.makedirs(dst, exist_ok=True)
    for name in os.listdir(src_tables):
        if name != "deltaf_tables":
            os.symlink(os.path.join(src_tables, name), os.path.join(dst, name))
    sd = os.path.join(src_tables, "deltaf_tables")
    dd = os.path.join(dst, "deltaf_tables")
    os.makedirs(dd)
    for name in os.listdir(sd):
        if name in ("urqmd", "smash"):
            os.makedirs(os.path.join(dd, name))
            for f in os.listdir(os.path.join(sd, name)):
                os.symlink(os.path.join(sd, name, f), os.path.join(dd, name, f))
            write_22mom_table(os.path.join(dd, name, "NEoSBQS_22mom_deltafCoeff.dat"))
        else:
            os.symlink(os.path.join(sd, name), os.path.join(dd, name))
    return dst
