#!/usr/bin/env python
"""Generate tests/golden/xc_golden.npz: small-case outputs of the CPU oracle (the reference itself cannot be built
here: Eigen/libxc/TBB are network-fetched), so the GPU tests have committed vectors that do not depend on the oracle
being rebuilt on the box.  Inputs are regenerated from seeds by the tests; only outputs are stored.
Also stores the reference's own known answers (tests/dft_tests.cpp:71-111)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from occ_b200 import aobasis, molecules  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [("water", "def2-svp", "b3lyp", 1, False), ("water", "def2-svp", "pbe", 2, False),
         ("water", "def2-svp", "r2scan", 1, True), ("water", "def2-svp", "svwn", 1, False),
         ("water", "def2-svp", "tpss", 2, False)]


def main():
    out = {}
    for mol, bas, method, nspin, pure in CASES:
        Z, pos = molecules.MOLECULES[mol]()
        b = aobasis.AOBasis.load(Z, pos, bas, pure=pure)
        ob = O.Basis(b.as_arrays())
        pts, w, _ = O.molecular_grid(Z, pos, ob)
        spts, sw, perm, leaves = O.hierarchy(pts, w)
        D2 = molecules.synthetic_density(b, nspin)
        r = O.xc_build(ob, spts, sw, leaves, O.METHODS[method], nspin, D2)
        key = f"{mol}_{bas}_{method}_{nspin}_{int(pure)}"
        out[key + "_K"] = r["K"]
        out[key + "_scalars"] = np.array([r["exc"], r["n_alpha"], r["n_beta"], len(sw), w.sum()])
    # pointwise functional table on a seeded sample (unpolarized + polarized), every implemented id
    rng = np.random.default_rng(2024)
    n = 64
    rho = 10 ** rng.uniform(-6, 0.7, size=(n, 2))
    g = rng.normal(size=(n, 2, 3)) * rho[..., None] ** (4 / 3)
    sig = np.stack([(g[:, 0] ** 2).sum(1), (g[:, 0] * g[:, 1]).sum(1), (g[:, 1] ** 2).sum(1)], 1)
    tau = np.stack([(g[:, k] ** 2).sum(1) / (8 * rho[:, k]) * rng.uniform(0.5, 4.0, n) for k in range(2)], 1)
    out["fx_rho"], out["fx_sigma"], out["fx_tau"] = rho, sig, tau
    for name in O.FUNC_IDS:
        u = O.xc_eval(name, False, rho.sum(1), sig[:, 0] + 2 * sig[:, 1] + sig[:, 2], tau.sum(1))
        p = O.xc_eval(name, True, rho, sig, tau)
        out[f"fx_{name}_u"] = np.concatenate([u["exc"][:, None], u["vrho"], u["vsigma"], u["vtau"]], 1)
        out[f"fx_{name}_p"] = np.concatenate([p["exc"][:, None], p["vrho"], p["vsigma"], p["vtau"]], 1)
    out["pbe_x_kat_rho"] = np.array([0.030097001731591706, 0.030097001731591706, 0.15085916775597902, 0.17751228531944346])
    out["pbe_x_kat_grad"] = np.array([[-0.063735860841572081, 0.0, 0.018656553309952275],
                                      [-0.063735860841572081, 0.0, 0.018656553309952275],
                                      [0.0, 0.0, 0.13035149660035189], [0.0, 0.0, 0.078426011080503066]])
    out["pbe_x_kat_exc"] = np.array([-0.27851489, -0.27851489, -0.39899553, -0.41654061])
    # tests/dft_tests.cpp:138-191: TPSS known answers (energies of x and c, potentials of c)
    out["tpss_kat_rows"] = np.array([
        [0.06019401627252133, -0.12747174847145987, 0.0, 0.03731311973522952, 0.09472176570942804, 0.036634279981580156],
        [0.06019401627252133, 0.0, -0.12747174847145987, 0.03731311973522952, 0.09472176570942804, 0.036634279981580156],
        [0.30171849914696935, 0.0, 0.0, 0.26070301592021455, -2.594981941107205, 0.02815789498408782],
        [0.3550246251878335, 0.0, 0.0, 0.15685208520532756, -6.307746108179051, 0.008662278222335566]])
    out["tpss_kat_x_exc"] = np.array([-0.33175882, -0.33175882, -0.56154382, -0.59257418])
    out["tpss_kat_c"] = np.array([[-0.02276044, -0.02276044, -0.04168478, -0.04268936],
                                  [-5.57565777e-02, -5.57565777e-02, -8.00593613e-02, -8.53775165e-02],
                                  [3.91761083e-02, 3.91761083e-02, 1.47458233e-01, 5.51608074e-01],
                                  [1.10283011e-03, 1.10283011e-03, -3.50002323e-01, -1.56487879e+00]])
    p = os.path.join(ROOT, "tests", "golden", "xc_golden.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes", len(out), "arrays")


if __name__ == "__main__":
    main()
