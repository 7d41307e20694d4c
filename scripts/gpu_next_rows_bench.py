#!/usr/bin/env python
"""Measurement of the SURVEY 8(f) rows next to their CPU oracle (GPU box): XC nuclear gradient, grid producer,
density on grid.  One JSON line per row -> profiles/r01_next_rows_bench.jsonl."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid
from oracle import oracle as O

def names(method):
    return [n for i in dft.get_dft_method(method) for n, v in O.FUNC_IDS.items() if v == i]

def main(mol="valinomycin", bas="def2-tzvp", method="b3lyp"):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    ob = O.Basis(b.as_arrays())
    # ---- grid producer (f-3): LMG + Lebedev + Becke/TA weights (CUDA kernel) + Morton hierarchy ----
    t0 = time.perf_counter(); mg = grid.MolecularGrid(b); h = mg.hierarchy(); t_first = time.perf_counter() - t0
    t0 = time.perf_counter(); mg2 = grid.MolecularGrid(b); h2 = mg2.hierarchy(); t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter(); pts, w, _ = O.molecular_grid(Z, pos, ob); O.hierarchy(pts, w); t_cpu = time.perf_counter() - t0
    print(json.dumps({"row": "f-3 grid producer", "workload": f"{mol}/{bas}", "grid_points": len(h.sorted_weights),
                      "natoms": len(Z), "ms_gpu_path": 1e3 * t_gpu, "ms_first_call": 1e3 * t_first,
                      "s_cpu_oracle": t_cpu, "cpu_cores": os.cpu_count()}), flush=True)
    D2 = molecules.synthetic_density(b, 1)
    # ---- XC nuclear gradient (f-1) ----
    d = dft.DFT(method, b, h, gradient=True)
    g = d.gradient(D2)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); g = d.gradient(D2); ts.append(time.perf_counter() - t0)
    info = d.info()
    d.close()
    stride = 40
    sample = np.concatenate([h.leaves[::stride], h.leaves[-1:]])
    t0 = time.perf_counter()
    O.xc_gradient(ob, h.sorted_points, h.sorted_weights, sample, names(method), 1, D2, len(Z))
    t_cpu = time.perf_counter() - t0
    npts_sample = int(sample["count"].sum())
    print(json.dumps({"row": "f-1 XC nuclear gradient", "workload": f"{mol} {method}/{bas} RKS", "grid_points": int(info["npts"]),
                      "ms_per_call": 1e3 * min(ts), "grid_points_per_s": info["npts"] / min(ts),
                      "dmma_flops": 1.5 * info["flops_per_build"],   # density + two A.D^T products, 2 n nb^2 each = 6 sum n nb^2
                      "cpu_oracle_points_per_s": npts_sample / t_cpu, "cpu_cores": os.cpu_count(),
                      "cpu_sample": f"every {stride}-th Morton leaf ({len(sample)} leaves, {npts_sample} points), {t_cpu:.1f} s",
                      "sum_gradient": [float(x) for x in g.sum(0)]}), flush=True)
    # ---- density on grid (f-4) ----
    P = np.asfortranarray(D2)
    t0 = time.perf_counter(); rho = dft.evaluate_density_on_grid(b, P, h.sorted_points, 1, screening_threshold=1e-10)
    t1 = time.perf_counter() - t0
    print(json.dumps({"row": "f-4 density on grid", "workload": f"{mol}/{bas} rho + grad rho", "points": len(rho),
                      "ms_total_incl_hierarchy_and_handle": 1e3 * t1,
                      "n_electrons_check": float((rho[:, 0] * h.sorted_weights).sum())}), flush=True)

if __name__ == "__main__":
    main()
