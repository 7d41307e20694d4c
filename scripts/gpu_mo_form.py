#!/usr/bin/env python
"""Developer experiment: D-form vs MO-form build, phase times (device events) for the BASELINE configs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid

def run(mol, bas, method, nspin):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    h = grid.MolecularGrid(b).hierarchy()
    Cs = molecules.synthetic_orbitals(b, nspin)
    D2 = molecules.synthetic_density(b, nspin)
    d = dft.DFT(method, b, h, dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED, nocc=Cs[0].shape[1])
    Kd = d.build(D2).copy(); ed = d.exchange_correlation_energy()
    Km = d.build_mo(D2, Cs).copy(); em = d.exchange_correlation_energy()
    d.set_profiling(True)
    d.build(D2); pd = d.phase_ms()
    d.build_mo(D2, Cs); pm = d.phase_ms()
    d.set_profiling(False)
    info = d.info()
    print(f"{mol} {method}/{bas} nspin={nspin} nocc={Cs[0].shape[1]} nbatches={info['nbatches']} dK={np.abs(Kd-Km).max():.1e} dE={abs(ed-em):.1e}")
    print("   D-form :", " ".join(f"{k}={v:.1f}" for k, v in pd.items()), f"sum={sum(pd.values()):.1f}")
    print("   MO-form:", " ".join(f"{k}={v:.1f}" for k, v in pm.items()), f"sum={sum(pm.values()):.1f}",
          f"density flops ratio={info['mo_density_flops'] / (0.5 * info['flops_per_build']):.3f}", flush=True)
    d.close()

if __name__ == "__main__":
    run("valinomycin", "def2-tzvp", "b3lyp", 1)
    run("taxol", "def2-tzvp", "r2scan", 1)
    run("h2o_32", "def2-svp", "b3lyp", 2)
    run("benzene", "def2-tzvp", "pbe", 1)
