#!/usr/bin/env python
"""Developer target for ncu: one VV10 energy call and one VV10 gradient call after a warm-up.  usage: gpu_vv10_once.py [mol basis]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, nlc
a = sys.argv[1:]
mol, bas = (a + ["h2o_32", "def2-svp"][len(a):])[:2]
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas)
D2 = molecules.synthetic_density(b, 1)
f = nlc.NonLocalCorrelationFunctional()
f.set_integration_grid(b, screening_threshold=1e-10, gradient=True)
print(f.vv10(D2)[0], abs(f.vv10_gradient(D2)[1]).max())
f.close()
