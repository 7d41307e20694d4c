#!/usr/bin/env python
"""Developer target for ncu: two builds of one workload (first = warm-up), nothing else.
   argv: molecule basis method nspin [debug] [mo]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid
a = sys.argv[1:]
mol, bas, method, nspin = (a[:4] + ["valinomycin", "def2-tzvp", "b3lyp", "1"][len(a[:4]):])[:4]
debug = int(a[4]) if len(a) > 4 else 0
mo = len(a) > 5 and a[5] == "mo"
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas)
h = grid.MolecularGrid(b).hierarchy()
D2 = molecules.synthetic_density(b, int(nspin))
Cs = molecules.synthetic_orbitals(b, int(nspin))
d = dft.DFT(method, b, h, dft.UNRESTRICTED if int(nspin) == 2 else dft.RESTRICTED, debug=debug,
            nocc=Cs[0].shape[1] if mo else None)
for _ in range(2):
    K = d.build_mo(D2, Cs) if mo else d.build(D2)
print("exc", d.exchange_correlation_energy())
d.close()
