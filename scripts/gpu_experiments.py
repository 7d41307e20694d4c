#!/usr/bin/env python
"""Developer timing experiments on the GPU box (phase times for batch sizes / debug switches)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid

def run(mol, bas, method, nspin, batch_mb, debug=0, reps=3, lanes=2):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    key = (mol, bas)
    if key not in run.cache:
        run.cache[key] = grid.MolecularGrid(b).hierarchy()
    h = run.cache[key]
    D2 = molecules.synthetic_density(b, nspin)
    d = dft.DFT(method, b, h, dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED, batch_bytes=batch_mb << 20, debug=debug, lanes=lanes)
    d.build(D2)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); d.build(D2); ts.append(time.perf_counter() - t0)
    d.set_profiling(True); d.build(D2); ph = d.phase_ms(); d.set_profiling(False)
    info = d.info()
    print(f"{mol} {method} nspin={nspin} batch={batch_mb}MB debug={debug} lanes={lanes} nbatches={info['nbatches']} "
          f"build={min(ts)*1e3:.1f}ms TF={info['flops_per_build']/min(ts)/1e12:.2f} "
          + " ".join(f"{k}={v:.1f}" for k, v in ph.items()), flush=True)
    d.close()
run.cache = {}

if __name__ == "__main__":
    for mb in (256, 1024, 4096, 16384):
        run("valinomycin", "def2-tzvp", "b3lyp", 1, mb)
    run("valinomycin", "def2-tzvp", "b3lyp", 1, 4096, debug=1)
    run("valinomycin", "def2-tzvp", "b3lyp", 1, 4096, debug=2)
    for mb in (64, 256, 2048):
        run("benzene", "def2-tzvp", "pbe", 1, mb)
    run("h2o_32", "def2-svp", "b3lyp", 2, 4096)
    run("taxol", "def2-tzvp", "r2scan", 1, 4096)
