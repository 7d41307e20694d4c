#!/usr/bin/env python
"""Phase times (CUDA events, profiling mode of the handle) of the small-nb BASELINE configurations, D form and factorised."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from occ_b200 import aobasis, dft, grid, molecules
import bench

for wl in sys.argv[1:] or ["water_b3lyp_svp", "benzene_pbe_tzvp", "h2o32_b3lyp_svp_uks"]:
    basis, method, nspin, quality, D2 = bench.make_inputs(wl)
    h = grid.MolecularGrid(basis).hierarchy()
    Cs = molecules.synthetic_orbitals(basis, nspin)
    kind = dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED
    d = dft.DFT(method, basis, h, kind, nocc=[c.shape[1] for c in Cs],
                batch_bytes=int(os.environ.get("OCC_XC_BATCH_MB", "0")) << 20,
                debug=int(os.environ.get("OCC_XC_DEBUG_SWITCHES", "0")))
    info = d.info()
    for form in os.environ.get("OCC_XC_FORMS", "D,mo").split(","):
        call = (lambda: d.build(D2)) if form == "D" else (lambda: d.build_mo(D2, Cs))
        call()
        d.set_profiling(True)
        ph = {}
        for _ in range(5):
            call()
            for k, v in d.phase_ms().items():
                ph[k] = ph.get(k, 0.0) + v / 5
        d.set_profiling(False)
        print(json.dumps({"workload": wl, "form": form, "debug": int(os.environ.get("OCC_XC_DEBUG_SWITCHES", "0")), "phases_ms": {k: round(v, 4) for k, v in ph.items()},
                          "sum_ms": round(sum(ph.values()), 4), "nb_min_max": [info["nb_min"], info["nb_max"]],
                          "ntiles": info["ntiles"], "nbatches": info["nbatches"], "flops": info["flops_per_build"],
                          "mo_density_flops": info["mo_density_flops"], "vxc_flops": info["vxc_flops"]}), flush=True)
    d.close()
