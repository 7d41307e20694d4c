#!/usr/bin/env python
"""Whole-grid CPU-oracle builds of the three large BASELINE.json configurations, reduced to a digest the GPU tests
compare against on the box (tests/test_fullsize_parity.py): Exc, N_alpha, N_beta, the Frobenius norm / sum / trace of K
and 256 seeded entries of K per spin block.  The oracle is the CPU restatement of the reference algorithm (oracle/);
inputs are the seeded synthetic densities of occ_b200.molecules on the oracle's own grid (bit-identical to the
product's grid producer, tests/test_gpu_parity.py::test_product_grid_producer_matches_oracle).

  python scripts/make_fullsize_digest.py [workload ...]      -> tests/golden/fullsize_digest.json
Takes minutes of CPU per configuration (valinomycin ~1 min, taxol at the 590 grid ~5 min on 8 cores)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from occ_b200 import aobasis, molecules  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {  # name: (molecule, basis, method, nspin, veryfine grid)
    "h2o32_b3lyp_svp_uks": ("h2o_32", "def2-svp", "b3lyp", 2, False),
    "valinomycin_b3lyp_tzvp": ("valinomycin", "def2-tzvp", "b3lyp", 1, False),
    "taxol_r2scan_tzvp": ("taxol", "def2-tzvp", "r2scan", 1, True),
}
OUT = os.path.join(ROOT, "tests", "golden", "fullsize_digest.json")


def sample_indices(nbf, nspin, seed=2025, count=256):
    rng = np.random.default_rng(seed)
    return [(int(s), int(i), int(j)) for s in range(nspin) for i, j in
            zip(rng.integers(nbf, size=count), rng.integers(nbf, size=count))]


def digest(name):
    mol, bas, method, nspin, veryfine = CASES[name]
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas)
    ob = O.Basis(b.as_arrays())
    s = O.default_grid_settings(max_angular_points=590, min_angular_points=302) if veryfine else O.default_grid_settings()
    t0 = time.perf_counter()
    pts, w, _ = O.molecular_grid(Z, pos, ob, s)
    spts, sw, perm, leaves = O.hierarchy(pts, w)
    t_grid = time.perf_counter() - t0
    D2 = molecules.synthetic_density(b, nspin)
    t0 = time.perf_counter()
    r = O.xc_build(ob, spts, sw, leaves, O.METHODS[method], nspin, D2)
    t_build = time.perf_counter() - t0
    K, nbf = r["K"], b.nbf
    idx = sample_indices(nbf, nspin)
    d = {"molecule": mol, "basis": bas, "method": method, "nspin": nspin, "grid": "veryfine" if veryfine else "default",
         "nbf": nbf, "npts": int(len(sw)), "nleaves": int(len(leaves)), "weight_sum": float(sw.sum()),
         "exc": r["exc"], "n_alpha": r["n_alpha"], "n_beta": r["n_beta"],
         "K_fro": float(np.sqrt((K * K).sum())), "K_sum": float(K.sum()),
         "K_trace": [float(np.trace(K[sp * nbf:(sp + 1) * nbf])) for sp in range(nspin)],
         "K_absmax": float(np.abs(K).max()),
         "K_asym": float(max(np.abs(K[sp * nbf:(sp + 1) * nbf] - K[sp * nbf:(sp + 1) * nbf].T).max() for sp in range(nspin))),
         "sample_seed": 2025, "K_samples": [float(K[sp * nbf + i, j]) for sp, i, j in idx],
         "sum_n_nb2": r["stats"]["sum_n_nb2"], "nb_min": r["stats"]["nb_min"], "nb_max": r["stats"]["nb_max"],
         "oracle_seconds": {"grid": t_grid, "build": t_build, "threads": os.cpu_count()}}
    print(name, "npts", d["npts"], "exc", d["exc"], "N", d["n_alpha"], "grid %.1fs build %.1fs" % (t_grid, t_build), flush=True)
    return d


def main():
    names = sys.argv[1:] or list(CASES)
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for n in names:
        out[n] = digest(n)
        json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
