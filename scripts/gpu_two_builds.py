#!/usr/bin/env python
"""Developer target for ncu: two builds of one workload (first = warm-up), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid
mol, bas, method, nspin = (sys.argv[1:5] + ["valinomycin", "def2-tzvp", "b3lyp", "1"][len(sys.argv) - 1:])[:4]
debug = int(sys.argv[5]) if len(sys.argv) > 5 else 0
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas)
h = grid.MolecularGrid(b).hierarchy()
D2 = molecules.synthetic_density(b, int(nspin))
d = dft.DFT(method, b, h, dft.UNRESTRICTED if int(nspin) == 2 else dft.RESTRICTED, debug=debug)
d.build(D2)
d.build(D2)
print("exc", d.exchange_correlation_energy())
d.close()
