#!/usr/bin/env python
"""Developer target for ncu: one XC-gradient call after a warm-up (valinomycin B3LYP/def2-TZVP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid
mol, bas, method = (sys.argv[1:4] + ["valinomycin", "def2-tzvp", "b3lyp"][len(sys.argv) - 1:])[:3]
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas)
h = grid.MolecularGrid(b).hierarchy()
D2 = molecules.synthetic_density(b, 1)
d = dft.DFT(method, b, h, gradient=True)
d.gradient(D2)
g = d.gradient(D2)
print(g.sum(0))
d.close()
