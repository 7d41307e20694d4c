#!/usr/bin/env python
"""COSX on 1 / 2 / 4 / 8 GPUs driven from ONE process (occ_cosx_multi_*): host D, C -> host K per call.
usage: gpu_cosx_multi.py [max_gpus [mol basis angular]]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, cosx

ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mol, bas, ang = (sys.argv[2:5] + ["valinomycin", "def2-tzvp", "50"][len(sys.argv) - 2 if len(sys.argv) > 2 else 0:])[:3]
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas, pure=False)
D = molecules.synthetic_density(b, 1) * 0.5
C = molecules.synthetic_orbitals(b, 1)[0]
one = cosx.SemiNumericalExchange(b, cosx.sgx_grid_settings(int(ang), 1e-5))
K1 = one.compute_K_mo(D, C)
grids, P = (one.atom_offsets, one.points, one.weights), one.overlap_projector
one.close()
n = 1
while n <= ngpu:
    t0 = time.perf_counter()
    m = cosx.MultiGPUSemiNumericalExchange(b, atom_grids=grids, devices=list(range(n)), projector=P)
    t_create = time.perf_counter() - t0
    K = m.compute_K(D, C)
    ts, td = [], []
    for _ in range(3):
        t0 = time.perf_counter(); K = m.compute_K(D, C); ts.append(time.perf_counter() - t0)
    for _ in range(2):
        t0 = time.perf_counter(); m.compute_K(D); td.append(time.perf_counter() - t0)
    print(json.dumps({"row": "f-2 COSX, one process", "workload": f"{mol}/{bas} RKS, sgx grid {ang} angular", "n_gpus": n,
                      "grid_points": len(grids[2]), "mo_form_ms_per_call": 1e3 * min(ts), "d_form_ms_per_call": 1e3 * min(td),
                      "max_abs_diff_vs_one_handle": float(np.abs(K - K1).max()), "create_s": t_create}), flush=True)
    m.close()
    n *= 2
