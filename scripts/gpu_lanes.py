#!/usr/bin/env python
"""Whole-build time (host wall clock around occ_xc_build_mo, pinned-less host buffers) for options.lanes = 1 / 2."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from occ_b200 import dft, grid, molecules
import bench

for wl in sys.argv[1:] or ["valinomycin_b3lyp_tzvp"]:
    basis, method, nspin, quality, D2 = bench.make_inputs(wl)
    h = grid.MolecularGrid(basis).hierarchy()
    Cs = molecules.synthetic_orbitals(basis, nspin)
    kind = dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED
    for lanes in (1, 2):
        for mb in (4096, 16384):
            d = dft.DFT(method, basis, h, kind, nocc=[c.shape[1] for c in Cs], lanes=lanes, batch_bytes=mb << 20)
            d.build_mo(D2, Cs)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); d.build_mo(D2, Cs); ts.append(time.perf_counter() - t0)
            print(json.dumps({"workload": wl, "lanes": lanes, "batch_mb": mb, "host_to_host_ms": 1e3 * min(ts),
                              "exc": d.exchange_correlation_energy()}), flush=True)
            d.close()
