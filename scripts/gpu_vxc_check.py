#!/usr/bin/env python
"""Developer check: resident-B Vxc kernel vs the legacy gbuild + k_vxc pair (debug bit 4) — results and phase times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid

def both(mol, bas, method, nspin):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    h = grid.MolecularGrid(b).hierarchy()
    D2 = molecules.synthetic_density(b, nspin)
    out = {}
    for dbg in (4, 0):
        d = dft.DFT(method, b, h, dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED, debug=dbg)
        K = d.build(D2).copy()
        d.set_profiling(True); d.build(D2); ph = d.phase_ms(); d.set_profiling(False)
        out[dbg] = (K, d.exchange_correlation_energy(), ph)
        d.close()
    dk = np.abs(out[0][0] - out[4][0]).max()
    print(f"{mol} {method} nspin={nspin}: max|K_new - K_legacy| = {dk:.2e}  legacy gbuild+vxc = "
          f"{out[4][2]['gbuild'] + out[4][2]['vxc']:.2f} ms, resident = {out[0][2]['gbuild'] + out[0][2]['vxc']:.2f} ms; "
          f"density {out[0][2]['density']:.2f}", flush=True)

if __name__ == "__main__":
    both("water", "def2-svp", "b3lyp", 1)
    both("benzene", "def2-tzvp", "r2scan", 2)
    both("benzene", "def2-tzvp", "svwn", 1)
    both("valinomycin", "def2-tzvp", "b3lyp", 1)
    both("taxol", "def2-tzvp", "r2scan", 1)
