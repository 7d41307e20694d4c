#!/usr/bin/env python
"""occ_xc_multi_build (several GPUs driven from ONE process, the layout the reference's single-process SCF needs):
host-to-host ms per build on 1..N devices, pinned and pageable host buffers, D form and factorised form.
The call is synchronous (returns when K is on the host), so it is timed on the host clock: median of `reps` calls.

  python scripts/gpu_multi_single_process.py [workload] [reps]"""
import ctypes as C
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from occ_b200 import dft, grid, molecules

wl = sys.argv[1] if len(sys.argv) > 1 else "valinomycin_b3lyp_tzvp"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
basis, method, nspin, quality, D2 = bench.make_inputs(wl)
h = grid.MolecularGrid(basis, grid.GridSettings.from_grid_quality(quality) if quality else None).hierarchy()
Cs = molecules.synthetic_orbitals(basis, nspin)
kind = dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED
nbf = basis.nbf
n = nspin * nbf * nbf
Cp = np.concatenate([c.ravel(order="F") for c in Cs])
hD = torch.from_numpy(np.ascontiguousarray(D2.ravel(order="F"))).pin_memory()
hC = torch.from_numpy(np.ascontiguousarray(Cp)).pin_memory()
hK = torch.empty(n, dtype=torch.float64).pin_memory()
pD, pC, pK = np.array(hD.numpy()), np.array(hC.numpy()), np.empty(n)       # pageable copies
ndev = torch.cuda.device_count()
K1 = None
for N in [x for x in (1, 2, 4, 8) if x <= ndev]:
    t0 = time.perf_counter()
    m = dft.MultiGPUDFT(method, basis, h, kind, devices=list(range(N)), nocc=[c.shape[1] for c in Cs])
    t_create = time.perf_counter() - t0
    row = {"workload": wl, "n_gpus": N, "create_s": round(t_create, 2), "grid_points": len(h.sorted_weights), "nbf": nbf}
    for name, args in (("D_form_pinned", (hD.data_ptr(), 0, hK.data_ptr())),
                       ("mo_form_pinned_C_only", (0, hC.data_ptr(), hK.data_ptr())),
                       ("mo_form_pageable_C_only", (0, pC.ctypes.data, pK.ctypes.data)),
                       ("D_form_pageable", (pD.ctypes.data, 0, pK.ctypes.data))):
        for _ in range(2):
            m.build_raw(*args)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            exc, na, nb = m.build_raw(*args)
            ts.append(time.perf_counter() - t0)
        row[name + "_ms"] = round(1e3 * statistics.median(ts), 3)
        row[name + "_min_ms"] = round(1e3 * min(ts), 3)
    row["exc"] = exc
    Kn = hK.numpy().copy()
    m.build_raw(hD.data_ptr(), 0, hK.data_ptr())
    Kd = hK.numpy().copy()
    if K1 is None:
        K1 = Kd
    row["max_abs_dK_vs_1gpu"] = float(np.abs(Kd - K1).max())
    row["grid_points_per_s_mo"] = round(len(h.sorted_weights) / (1e-3 * row["mo_form_pinned_C_only_ms"]))
    m.close()
    print(json.dumps(row), flush=True)
