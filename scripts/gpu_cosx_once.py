#!/usr/bin/env python
"""Developer target for ncu: one COSX compute_K after a warm-up.  usage: gpu_cosx_once.py mol basis pure angular"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, cosx
a = sys.argv[1:]
mol, bas, pure, ang = (a + ["h2o_32", "def2-svp", "0", "50"][len(a):])[:4]
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas, pure=pure == "1")
D = molecules.synthetic_density(b, 1) * 0.5
sgx = cosx.SemiNumericalExchange(b, cosx.sgx_grid_settings(int(ang)), projector=None)
sgx.compute_K(D)
K = sgx.compute_K(D)
print(abs(K).max())
sgx.close()
