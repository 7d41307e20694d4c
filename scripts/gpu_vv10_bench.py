#!/usr/bin/env python
"""VV10 timing on the GPU box: whole occ_xc_vv10 call and the CPU oracle on the same NLC grid."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, grid, nlc

def run(mol, bas, thr, cpu=False, devices=None):
    """devices: list of CUDA ordinals -> occ_xc_multi_vv10 / occ_xc_multi_vv10_gradient (one process, several GPUs)"""
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    D2 = molecules.synthetic_density(b, 1)
    f = nlc.NonLocalCorrelationFunctional()
    t0 = time.perf_counter()
    f.set_integration_grid(b, screening_threshold=thr, gradient=True, devices=devices)
    t_setup = time.perf_counter() - t0
    e, K = f.vv10(D2)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); e, K = f.vv10(D2); ts.append(time.perf_counter() - t0)
    eg, g = f.vv10_gradient(D2)
    tg = []
    for _ in range(2):
        t0 = time.perf_counter(); eg, g = f.vv10_gradient(D2); tg.append(time.perf_counter() - t0)
    n = f.num_grid_points
    out = {"workload": f"vv10 {mol}/{bas}", "n_gpus": len(devices) if devices else 1, "nbf": b.nbf, "nlc_grid_points": n,
           "pairs": n * (n - 1) // 2, "screening_threshold": thr, "ms_per_call": 1e3 * min(ts),
           "pairs_per_s": n * (n - 1) / 2 / min(ts), "energy": e, "setup_s": t_setup,
           "gradient_ms_per_call": 1e3 * min(tg), "gradient_absmax": float(np.abs(g).max()),
           "gradient_energy_minus_energy": eg - e}
    if cpu:
        from oracle import oracle as O
        mg = grid.MolecularGrid(b, grid.GridSettings(110, 50, 1e-7, True, True))
        pts, w = mg.points()
        t0 = time.perf_counter(); r = O.vv10(O.Basis(b.as_arrays()), pts, w, D2); tc = time.perf_counter() - t0
        out["cpu_oracle_s"] = tc
        out["cpu_cores"] = os.cpu_count()
        out["dE_vs_oracle"] = abs(e - r["energy"])
        out["dK_vs_oracle"] = float(np.abs(K - r["K"]).max())
        if n < 150000:   # the oracle's O(N^2) grid term on all host cores
            off = np.concatenate([[0], np.cumsum([mg.atom_info(a)[0] for a in range(mg.natoms)])]).astype(np.uint64)
            t0 = time.perf_counter(); rg = O.vv10_gradient(O.Basis(b.as_arrays()), pts, w, off, D2)
            out["cpu_oracle_gradient_s"] = time.perf_counter() - t0
            out["dgrad_vs_oracle"] = float(np.abs(g - rg["gradient"]).max())
    print(json.dumps(out), flush=True)
    f.close()

if __name__ == "__main__":
    ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    if ngpu == 1:
        run("benzene", "def2-tzvp", 0.0, cpu=True)
        run("h2o_32", "def2-svp", 1e-10, cpu=True)
        run("valinomycin", "def2-tzvp", 1e-10, cpu=False)
    else:
        n = 1
        while n <= ngpu:
            run("valinomycin", "def2-tzvp", 1e-10, cpu=False, devices=list(range(n)))
            n *= 2
