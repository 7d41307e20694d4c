#!/usr/bin/env python
"""Small end-to-end exercise of every entry point, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft, grid, nlc
which = sys.argv[1] if len(sys.argv) > 1 else "all"
Z, pos = molecules.MOLECULES["water"]()
b = aobasis.AOBasis.load(Z, pos, "def2-svp")
h = grid.MolecularGrid(b).hierarchy()
leaves = h.leaves[::8]                      # a thin sample keeps the sanitizer run short
hs = dft.GridHierarchy(h.sorted_points, h.sorted_weights, leaves)
for method, nspin in (("b3lyp", 1), ("tpss", 2), ("svwn", 1), ("b97", 2), ("b3pw91", 1), ("bp86", 2),
                      ("wb97x-v", 1), ("b3p86", 2), ("b86bpbe", 1)):
    D2 = molecules.synthetic_density(b, nspin)
    d = dft.DFT(method, b, hs, dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED, gradient=True)
    K = d.build(D2)
    g = d.gradient(D2)
    print(method, nspin, float(np.abs(K).max()), float(np.abs(g).max()))
    if d.have_nonlocal_correlation():             # wB97X-V: VV10 energy and gradient through the DFT mirror
        d.set_nlc_grid(grid.GridSettings(50, 26, 1e-5, False, True))
        print("nlc", d.post_scf_nlc_correction(0.5 * D2), float(np.abs(d.compute_xc_gradient(0.5 * D2)).max()))
    d.close()
for devices in (None, [0, 0]):                      # single handle, then two shards of one process (peer-memory paths)
    f = nlc.NonLocalCorrelationFunctional()
    f.set_integration_grid(b, grid.GridSettings(50, 26, 1e-5, False, True), gradient=True, devices=devices)
    print("vv10", devices, f.vv10(molecules.synthetic_density(b, 1))[0])
    print("vv10 gradient", float(np.abs(f.vv10_gradient(molecules.synthetic_density(b, 1))[1]).max()))
    f.close()
# multi-handle build (MO and D form), deterministic variant
Cs = molecules.synthetic_orbitals(b, 1)
m = dft.MultiGPUDFT("b3lyp", b, hs, devices=[0, 0], nocc=Cs[0].shape[1])
print("multi", float(np.abs(m.build(molecules.synthetic_density(b, 1), Cs)).max()))
m.close()
m = dft.MultiGPUDFT("pbe", b, hs, dft.UNRESTRICTED, devices=[0, 0], gradient=True)
print("multi gradient", float(np.abs(m.gradient(molecules.synthetic_density(b, 2))).max()))
m.close()
m = dft.MultiGPUDFT("b3lyp", b, hs, devices=[0, 0], deterministic=True)
print("multi det", float(np.abs(m.build(molecules.synthetic_density(b, 1))).max()))
m.close()
# COSX: s / p / d (def2-SVP) and up to g shells (def2-QZVP), ORCA GRIDX 1 grid, D and MO forms
from occ_b200 import cosx
for bas in ("def2-svp", "def2-qzvp"):
    bs = aobasis.AOBasis.load(Z, pos, bas, pure=True)
    sgx = cosx.SemiNumericalExchange(bs, settings=cosx.cosx_grid_settings(1))
    D = 0.5 * molecules.synthetic_density(bs, 1)
    print("cosx", bas, float(np.abs(sgx.compute_K(D)).max()),
          float(np.abs(sgx.compute_K_mo(D, molecules.synthetic_orbitals(bs, 1)[0])).max()))
    sgx.close()
    if bas == "def2-svp":
        mg = cosx.MultiGPUSemiNumericalExchange(bs, settings=cosx.cosx_grid_settings(1), devices=[0, 0])
        print("cosx multi", float(np.abs(mg.compute_K(D, molecules.synthetic_orbitals(bs, 1)[0])).max()))
        mg.close()
rho = dft.evaluate_density_on_grid(b, molecules.synthetic_density(b, 1), np.random.default_rng(1).uniform(-3, 3, (300, 3)), 2)
print("rho", float(rho[:, 0].max()))
