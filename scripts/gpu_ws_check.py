#!/usr/bin/env python
"""Developer check: Vxc through the warp-specialised kernel (OCC_XC_VXC_WS) against the default kernel, same handle inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid
for mol, bas, method in (("water", "def2-svp", "pbe"), ("benzene", "def2-tzvp", "pbe"), ("h2o_32", "def2-svp", "b3lyp")):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    h = grid.MolecularGrid(b).hierarchy()
    D2 = molecules.synthetic_density(b, 1)
    Ks = []
    for ws in (False, True):
        if ws:
            os.environ["OCC_XC_VXC_WS"] = "1"
        else:
            os.environ.pop("OCC_XC_VXC_WS", None)
        d = dft.DFT(method, b, h)
        K = d.build(D2)
        K2 = d.build(D2)
        Ks.append(K)
        print(mol, "ws" if ws else "default", "exc", d.exchange_correlation_energy(), "repeat diff", float(np.abs(K - K2).max()))
        d.close()
    diff = np.abs(Ks[0] - Ks[1])
    i, j = np.unravel_index(np.argmax(diff), diff.shape)
    print(mol, "max diff", float(diff.max()), "at", i, j, "K there", Ks[0][i, j], Ks[1][i, j], "nnz diff", int((diff > 1e-10).sum()), "of", diff.size)
