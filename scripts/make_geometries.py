#!/usr/bin/env python
"""Freeze the seeded synthetic stand-in geometries as .xyz fixtures (occ_b200/data/geometries)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import molecules as m  # noqa: E402
from occ_b200.aobasis import ELEMENT_SYMBOLS  # noqa: E402


def write(name, Z, pos, comment):
    os.makedirs(m.GEOM_DIR, exist_ok=True)
    with open(os.path.join(m.GEOM_DIR, name), "w") as f:
        f.write(f"{len(Z)}\n{comment}\n")
        for z, p in zip(Z, pos):
            f.write(f"{ELEMENT_SYMBOLS[z]:2s} {p[0]: .10f} {p[1]: .10f} {p[2]: .10f}\n")
    print(name, len(Z), "atoms")


Z, pos = m.generate_water_cluster(seed=32)
write("h2o_32.xyz", Z, pos, "(H2O)32 synthetic cluster, jittered 4x4x2 lattice 3.1 A, numpy default_rng(32)")
Z, pos = m.generate_globule(54, 90, 6, 18, 9.0, 168)
write("valinomycin_standin.xyz", Z, pos, "C54H90N6O18 composition-exact synthetic stand-in for valinomycin, default_rng(168)")
Z, pos = m.generate_globule(47, 51, 1, 14, 8.0, 113)
write("taxol_standin.xyz", Z, pos, "C47H51NO14 composition-exact synthetic stand-in for taxol, default_rng(113)")
