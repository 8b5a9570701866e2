"""Materialises a golden fixture (tests/golden/*.npz) as a work folder the engine can read:
<dir>/music_input + <dir>/surface.dat, plus the parameter file and overrides to use."""
import os
import shutil

import numpy as np

from iss_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
FIX = os.path.join(HERE, "fixtures")

# viscous3 / viscous4 and the *_22mom cases run bulk_deltaf_kind = 20 (22-moment), the kind the
# reference's own CI parameter file selects; its table blob is not in the reference tree, so both
# the compiled reference (make_golden.py) and the engine read the synthetic table of
# iss_b200/synthetic.py (tables_for()).
ONE_CELL = ["ideal1", "ideal2", "ideal3", "ideal4", "viscous1", "viscous2", "viscous3", "viscous4"]
SYNTH = ["s3d_ce", "s3d_ce_diff", "s3d_14mom", "s2d_smash_ce", "s3d_ideal_b", "s3d_bulk1",
         "s3d_boltzmann", "s2d_urqmd_bin", "s3d_22mom", "s3d_22mom_diff"]
REPO_TABLES = os.path.join(os.path.dirname(HERE), "iSS_tables")


def load(name, kind="yields"):
    return np.load(os.path.join(GOLDEN, "%s_%s.npz" % (kind, name)), allow_pickle=False)


def overrides_of(g):
    out = {}
    for kv in g["overrides"]:
        k, v = str(kv).split("=")
        if k.startswith("oracle_"):
            continue
        out[k] = float(v)
    return out


def bulk_kind_of(g):
    kind = None
    for line in open(os.path.join(FIX, str(g["param"]))):
        line = line.split("#")[0]
        if "=" in line and line.split("=")[0].strip() == "bulk_deltaf_kind":
            kind = int(float(line.split("=")[1]))
    return int(overrides_of(g).get("bulk_deltaf_kind", kind))


_tables_22mom = []


def tables_for(g):
    """Table folder for a fixture: the repo's iSS_tables, or (kind 20) a per-process temporary
    mirror of it with the synthetic 22-moment table added."""
    if bulk_kind_of(g) != 20:
        return REPO_TABLES
    if not _tables_22mom:
        import atexit
        import tempfile
        d = tempfile.mkdtemp(prefix="iss_tables22_")
        atexit.register(shutil.rmtree, d, True)
        _tables_22mom.append(synthetic.tables_with_22mom(os.path.join(d, "iSS_tables"), REPO_TABLES))
    return _tables_22mom[0]


def materialise(g, folder):
    """Returns (param_file, surface_name, overrides dict)."""
    os.makedirs(folder, exist_ok=True)
    param = os.path.join(FIX, str(g["param"]))
    if "cells" in g.files:
        gen = {k: v for k, v in g["gen"]}
        kw = dict(eos=int(gen["eos"]), bulk=1, rhob=int(gen.get("rhob", 0)),
                  diffusion=int(gen.get("diffusion", 0)), binary=int(gen.get("binary", 0)))
        synthetic.write_music_input(folder, **kw)
        synthetic.write_surface(os.path.join(folder, "surface.dat"), g["cells"], kw["binary"],
                                kw["bulk"], kw["rhob"], kw["diffusion"])
        return param, "surface.dat", overrides_of(g)
    if "cell_line" in g.files:
        shutil.copy(os.path.join(FIX, "music_input_" + str(g["music"])),
                    os.path.join(folder, "music_input"))
        np.savetxt(os.path.join(folder, "surface.dat"), g["cell_line"][None, :], fmt="%.16e")
        return param, "surface.dat", overrides_of(g)
    shutil.copy(os.path.join(FIX, "music_input_" + str(g["music_input"])),
                os.path.join(folder, "music_input"))
    surf = str(g["surface"])
    shutil.copy(os.path.join(FIX, surf), os.path.join(folder, surf))
    return param, surf, overrides_of(g)
