#!/usr/bin/env python
"""VV10 timing on the GPU box: whole occ_xc_vv10 call and the CPU oracle on the same NLC grid."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, grid, nlc

def run(mol, bas, thr, cpu=False):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    D2 = molecules.synthetic_density(b, 1)
    f = nlc.NonLocalCorrelationFunctional()
    t0 = time.perf_counter(); f.set_integration_grid(b, screening_threshold=thr); t_setup = time.perf_counter() - t0
    e, K = f.vv10(D2)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); e, K = f.vv10(D2); ts.append(time.perf_counter() - t0)
    n = f.num_grid_points
    out = {"workload": f"vv10 {mol}/{bas}", "nbf": b.nbf, "nlc_grid_points": n, "pairs": n * (n - 1) // 2,
           "screening_threshold": thr, "ms_per_call": 1e3 * min(ts), "pairs_per_s": n * (n - 1) / 2 / min(ts),
           "energy": e, "setup_s": t_setup}
    if cpu:
        from oracle import oracle as O
        mg = grid.MolecularGrid(b, grid.GridSettings(110, 50, 1e-7, False, True))
        pts, w = mg.points()
        t0 = time.perf_counter(); r = O.vv10(O.Basis(b.as_arrays()), pts, w, D2); tc = time.perf_counter() - t0
        out["cpu_oracle_s"] = tc
        out["cpu_cores"] = os.cpu_count()
        out["dE_vs_oracle"] = abs(e - r["energy"])
        out["dK_vs_oracle"] = float(np.abs(K - r["K"]).max())
    print(json.dumps(out), flush=True)
    f.close()

if __name__ == "__main__":
    run("benzene", "def2-tzvp", 0.0, cpu=True)
    run("h2o_32", "def2-svp", 1e-10, cpu=True)
    run("valinomycin", "def2-tzvp", 1e-10, cpu=False)
