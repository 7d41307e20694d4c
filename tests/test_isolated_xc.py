"""The reference's own end-to-end pins for this path: tests/isolated_xc_tests.cpp:71-319 — exchange-correlation
energies of the converged RHF/3-21G (spherical) water density against PySCF:
    Slater+VWN3 -8.786388687644, PBE -9.262866622721, B88+LYP -9.325386381715, r2SCAN -9.313583890330,
    wB97X -6.577307030585.
The density comes from scripts/rhf_sp.py (a 300-line s/p Hartree-Fock, E = -75.585401384 Eh) and is committed as
tests/golden/water_321g_rhf.npz.

What the reference really asserts.  `REQUIRE(xc_energy == Catch::Approx(g).margin(1e-8))` is NOT a 1e-8 test: the
reference pins Catch2 v3.5.2 (tests/CMakeLists.txt:1-5), whose Approx is constructed with
epsilon = 100 * std::numeric_limits<float>::epsilon() = 1.19e-5 and compares
    |x - g| <= margin  ||  |x - g| <= epsilon * (scale + |g|)          (catch_approx.cpp, Approx::equalityComparisonImpl)
so the effective tolerance of the reference's assertions is 1.19e-5 * |g| ~ 1.0e-4 Eh (7.8e-5 for wB97X).  That is why
the reference's test passes although its default grid integrates N_el only to 2e-6, and why its own comment can
report a 1.7e-5 deviation for r2SCAN on the VeryFine grid under a nominal margin of 2e-5 / 1e-6.
`catch_approx` below is that rule; the tests apply it with the reference's margins — and then pin the functionals
much harder than the reference does:

  * on an emulation of the quadrature PySCF used for those numbers ("Level 3 grid": Treutler-Ahlrichs M4 radial grids
    of 75 / 50 points scaled by the Treutler xi, NWChem pruning {50, 86, 266, 302, 266}, Becke partition with
    Treutler's radius adjustment; 33 698 points against the "29,200-33,700" the reference quotes at :268) all five
    energies agree with PySCF to 1.1e-7 (r2SCAN 4.4e-7), with a COMMON offset of +1.1e-7 that is the convergence of
    PySCF's own HF density;
  * on a converged unpruned grid (N_el to 4e-9) LDA / PBE / BLYP agree to 1e-7;
  * the 1.2e-6 .. 2.1e-6 deviations on OCC's default 302-point pruned grid are therefore quadrature error of that
    grid (it integrates N_el to 2.2e-6), not a deviation of the grid restatement: the VeryFine grid reproduces the
    reference's own point count (38 652) and its reported r2SCAN deviation (+1.73e-5 vs "~1.7e-5").
  * wB97X discriminates libxc's PW92 parameter set: with the 1992 constants the PySCF-grid energy is +0.9e-7 from
    the golden, with the "modified" constants +4.6e-6.
"""
import os

import numpy as np
import pytest

from occ_b200 import aobasis, molecules

GOLD = {"lda": -8.786388687644, "pbe": -9.262866622721, "blyp": -9.325386381715, "r2scan": -9.313583890330,
        "wb97x": -6.577307030585, "wb97x-v": -6.539400300836}
# margins of the reference's assertions (isolated_xc_tests.cpp:94,138,182,228,274,319,364)
REF_MARGIN_DEFAULT = {"lda": 1e-8, "pbe": 1e-8, "blyp": 1e-8, "r2scan": 1e-6, "wb97x": 1e-8, "wb97x-v": 1e-8}
HERE = os.path.dirname(os.path.abspath(__file__))
BOHR_PYSCF = 0.52917721092
BRAGG = {1: 0.35 / BOHR_PYSCF, 8: 0.60 / BOHR_PYSCF}
TREUTLER_XI = {1: 0.8, 8: 0.9}          # Treutler & Ahlrichs, J. Chem. Phys. 102, 346 (1995), Table 1


def catch_approx(x: float, target: float, margin: float = 0.0, epsilon: float = 100.0 * float(np.finfo(np.float32).eps),
                 scale: float = 0.0) -> bool:
    """Catch::Approx(target).margin(margin) == x in Catch2 v3.5.2 (the version the reference's tests build against)."""
    d = abs(x - target)
    return d <= margin or d <= epsilon * (scale + abs(target))


def _setup():
    Z, pos = molecules.water_isolated_xc()
    b = aobasis.AOBasis.load(Z, pos, "3-21g", pure=True)
    g = np.load(os.path.join(HERE, "golden", "water_321g_rhf.npz"))
    return Z, pos, b, np.asfortranarray(g["P"]), g["S"], float(g["energy"])


def _becke_cells(pts, Z, pos):
    """Becke cell functions with Treutler's radius adjustment a_ij = (1/chi - chi)/4, chi = sqrt(R_i/R_j)"""
    d = np.linalg.norm(pts[:, None, :] - pos[None, :, :], axis=2)
    cell = np.ones((len(pts), len(pos)))
    rad = np.sqrt(np.array([BRAGG[z] for z in Z]))
    for i in range(len(pos)):
        for j in range(i):
            mu = (d[:, i] - d[:, j]) / np.linalg.norm(pos[i] - pos[j])
            a = min(max(0.25 * (rad[j] / rad[i] - rad[i] / rad[j]), -0.5), 0.5)
            mu = mu - a * (mu * mu - 1.0)
            for _ in range(3):
                mu = (3.0 - mu * mu) * mu * 0.5
            cell[:, i] *= 0.5 * (1.0 - mu)
            cell[:, j] *= 0.5 * (1.0 + mu)
    return cell


def pyscf_level3_grid(Z, pos, lebedev):
    """Test-side emulation of the quadrature behind the goldens (PySCF grid level 3 for H / O; see the module text)."""
    pos = np.asarray(pos, float).reshape(-1, 3)
    orders = np.array([38, 50, 74, 86, 110, 146, 170, 194, 230, 266, 302])
    alphas = {1: (0.25, 0.5, 1.0, 4.5), 8: (0.1667, 0.5, 0.9, 3.5)}
    pts_all, w_all = [], []
    for a, z in enumerate(Z):
        n = 50 if z <= 2 else 75
        step = np.pi / (n + 1)
        i = np.arange(1, n + 1)
        x = np.cos(i * step)
        ln2 = 1.0 / np.log(2.0)
        r = (-ln2 * (1 + x) ** 0.6 * np.log((1 - x) / 2))[::-1] * TREUTLER_XI[z]
        dr = (step * np.sin(i * step) * ln2 * (1 + x) ** 0.6 * (-0.6 / (1 + x) * np.log((1 - x) / 2) + 1 / (1 - x)))[::-1] \
            * TREUTLER_XI[z]
        place = ((r / BRAGG[z]).reshape(-1, 1) > np.array(alphas[z])).sum(axis=1)
        angs = orders[np.array([1, 3, 9, 10, 9])[place]]
        p, w = [], []
        for k in range(n):
            t = lebedev[f"n{angs[k]}"]
            p.append(r[k] * t[:, :3])
            w.append(4 * np.pi * r[k] ** 2 * dr[k] * t[:, 3])
        p = np.vstack(p) + pos[a]
        c = _becke_cells(p, Z, pos)
        pts_all.append(p)
        w_all.append(np.concatenate(w) * c[:, a] / c.sum(1))
    return np.ascontiguousarray(np.vstack(pts_all)), np.ascontiguousarray(np.concatenate(w_all))


def converged_grid(Z, pos, lebedev, nrad=120, nang=590):
    """Unpruned Gauss-Chebyshev (2nd kind, Becke mapping) x Lebedev grid: N_el of the RHF density to 4e-9"""
    pos = np.asarray(pos, float).reshape(-1, 3)
    t = lebedev[f"n{nang}"]
    pts_all, w_all = [], []
    for a, z in enumerate(Z):
        i = np.arange(1, nrad + 1)
        x = np.cos(i * np.pi / (nrad + 1))
        rm = BRAGG[z]
        r = rm * (1 + x) / (1 - x)
        wr = np.pi / (nrad + 1) * np.sin(i * np.pi / (nrad + 1)) ** 2 * 2 * rm / ((1 - x) ** 2 * np.sqrt(1 - x * x))
        p = (r[:, None, None] * t[None, :, :3]).reshape(-1, 3) + pos[a]
        c = _becke_cells(p, Z, pos)
        w = ((wr * r * r * 4 * np.pi)[:, None] * t[None, :, 3]).reshape(-1) * c[:, a] / c.sum(1)
        keep = w > 1e-30
        pts_all.append(p[keep])
        w_all.append(w[keep])
    return np.ascontiguousarray(np.vstack(pts_all)), np.ascontiguousarray(np.concatenate(w_all))


def _lebedev():
    return np.load(os.path.join(HERE, "golden", "lebedev_tables.npz"))


def test_catch_approx_rule():
    # margin alone would fail, the default epsilon of Catch2 accepts: the reference's default-grid assertions
    assert not abs(-8.786390 - GOLD["lda"]) <= 1e-8
    assert catch_approx(-8.786390, GOLD["lda"], margin=1e-8)
    assert not catch_approx(-8.7865, GOLD["lda"], margin=1e-8)        # 1.1e-4 away: rejected
    assert catch_approx(1.0 + 5e-9, 1.0, margin=1e-8, epsilon=0.0)


def test_rhf_fixture_is_a_converged_closed_shell_density():
    Z, pos, b, P, S, e = _setup()
    assert b.nbf == 13 and abs(np.sum(P * S) - 10.0) < 1e-10
    assert np.abs(P @ S @ P - 2.0 * P).max() < 1e-9          # idempotent: P S P = 2 P
    assert abs(e - (-75.5854013843)) < 1e-8


@pytest.mark.parametrize("quality", ["default", "veryfine"])
def test_oracle_reproduces_the_reference_assertions_on_occ_grids(oracle, quality):
    Z, pos, b, P, S, e = _setup()
    ob = oracle.Basis(b.as_arrays())
    s = oracle.default_grid_settings() if quality == "default" else \
        oracle.default_grid_settings(max_angular_points=590, min_angular_points=302)
    pts, w, _ = oracle.molecular_grid(Z, pos, ob, s)
    spts, sw, perm, leaves = oracle.hierarchy(pts, w)
    for m in GOLD:
        r = oracle.xc_build(ob, spts, sw, leaves, oracle.METHODS[m], 1, P)
        diff = r["exc"] - GOLD[m]
        if quality == "default":
            assert catch_approx(r["exc"], GOLD[m], margin=REF_MARGIN_DEFAULT[m]), (m, diff)  # the reference's REQUIRE
            assert abs(r["n_alpha"] - 10.0) < 2.5e-6          # quadrature error of the 302-point pruned grid ...
            assert abs(diff) < (1.5e-5 if m == "r2scan" else 4e-6), (m, diff)   # ... which bounds the energy error
        elif m == "r2scan":
            assert len(w) == 38652                            # isolated_xc_tests.cpp:271
            assert catch_approx(r["exc"], GOLD[m], margin=2e-5)
            assert 1.5e-5 < diff < 1.9e-5                     # the deviation the reference reports for its own build


def test_oracle_energies_on_the_pyscf_quadrature_and_on_a_converged_grid(oracle):
    Z, pos, b, P, S, e = _setup()
    ob = oracle.Basis(b.as_arrays())
    pts, w = pyscf_level3_grid(Z, pos, _lebedev())
    assert len(w) == 33698                                    # "29,200-33,700 points", isolated_xc_tests.cpp:268
    spts, sw, perm, leaves = oracle.hierarchy(pts, w)
    for m in GOLD:
        r = oracle.xc_build(ob, spts, sw, leaves, oracle.METHODS[m], 1, P)
        assert abs(r["exc"] - GOLD[m]) < (6e-7 if m == "r2scan" else 2e-7), (m, r["exc"] - GOLD[m])
        assert abs(r["n_alpha"] - 10.0) < 2e-8
    pts, w = converged_grid(Z, pos, _lebedev())
    spts, sw, perm, leaves = oracle.hierarchy(pts, w)
    for m in ("lda", "pbe", "blyp"):
        r = oracle.xc_build(ob, spts, sw, leaves, oracle.METHODS[m], 1, P)
        assert abs(r["exc"] - GOLD[m]) < 2e-7, (m, r["exc"] - GOLD[m])
        assert abs(r["n_alpha"] - 10.0) < 1e-8


@pytest.mark.gpu
def test_cuda_xc_energies_against_pyscf_goldens():
    from occ_b200 import dft, grid
    Z, pos, b, P, S, e = _setup()
    for quality in ("default", "veryfine"):
        h = grid.MolecularGrid(b, None if quality == "default" else grid.GridSettings.from_grid_quality("veryfine")).hierarchy()
        for m in GOLD:
            d = dft.DFT(m, b, h)
            d.compute_K_dft(0.5 * P)                          # mo.D = P / 2 (dft.h:205 doubles it)
            x = d.exchange_correlation_energy()
            if quality == "default":
                assert catch_approx(x, GOLD[m], margin=REF_MARGIN_DEFAULT[m]), (m, x - GOLD[m])
                assert abs(x - GOLD[m]) < (1.5e-5 if m == "r2scan" else 4e-6), (m, x - GOLD[m])
            elif m == "r2scan":
                assert catch_approx(x, GOLD[m], margin=2e-5) and 1.5e-5 < x - GOLD[m] < 1.9e-5
            d.close()


@pytest.mark.gpu
def test_cuda_xc_energies_on_the_pyscf_quadrature():
    """the product path on the quadrature that produced the goldens: all five energies to 2e-7 (r2SCAN 6e-7)"""
    from occ_b200 import dft, grid
    Z, pos, b, P, S, e = _setup()
    pts, w = pyscf_level3_grid(Z, pos, _lebedev())
    mg = grid.MolecularGrid.from_points(pts, w)
    h = mg.hierarchy()
    mg.close()
    for m in GOLD:
        d = dft.DFT(m, b, h)
        d.compute_K_dft(0.5 * P)
        assert abs(d.exchange_correlation_energy() - GOLD[m]) < (6e-7 if m == "r2scan" else 2e-7), m
        assert abs(d.total_density[0] - 10.0) < 2e-8
        d.close()
