"""Materialises a golden fixture (tests/golden/*.npz) as a work folder the engine can read:
<dir>/music_input + <dir>/surface.dat, plus the parameter file and overrides to use."""
import os
import shutil

import numpy as np

from iss_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
FIX = os.path.join(HERE, "fixtures")

ONE_CELL = ["ideal1", "ideal2", "ideal3", "ideal4", "viscous1", "viscous2"]
SYNTH = ["s3d_ce", "s3d_ce_diff", "s3d_14mom", "s2d_smash_ce", "s3d_ideal_b", "s3d_bulk1",
         "s3d_boltzmann", "s2d_urqmd_bin"]


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
