#!/usr/bin/env python
"""Small end-to-end exercise of every entry point, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid, nlc
which = sys.argv[1] if len(sys.argv) > 1 else "all"
Z, pos = molecules.MOLECULES["water"]()
b = aobasis.AOBasis.load(Z, pos, "def2-svp")
h = grid.MolecularGrid(b).hierarchy()
leaves = h.leaves[::8]                      # a thin sample keeps the sanitizer run short
hs = dft.GridHierarchy(h.sorted_points, h.sorted_weights, leaves)
for method, nspin in (("b3lyp", 1), ("tpss", 2), ("svwn", 1)):
    D2 = molecules.synthetic_density(b, nspin)
    d = dft.DFT(method, b, hs, dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED, gradient=True)
    K = d.build(D2)
    g = d.gradient(D2)
    print(method, nspin, float(np.abs(K).max()), float(np.abs(g).max()))
    d.close()
f = nlc.NonLocalCorrelationFunctional()
f.set_integration_grid(b, grid.GridSettings(50, 26, 1e-5, False, True))
print("vv10", f.vv10(molecules.synthetic_density(b, 1))[0])
f.close()
rho = dft.evaluate_density_on_grid(b, molecules.synthetic_density(b, 1), np.random.default_rng(1).uniform(-3, 3, (300, 3)), 2)
print("rho", float(rho[:, 0].max()))
