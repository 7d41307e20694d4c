"""Pins of the CPU oracle (test infrastructure) against everything the reference offers for this path:
its own gau2grid compiled into oracle/_ref, the PBE-x known-answer test of tests/dft_tests.cpp:71-111, the grid size
quoted at tests/isolated_xc_tests.cpp:271, the Lebedev tables produced by its lebedev.cpp, and the Morton-hierarchy
invariants of tests/dft_tests.cpp:1287-1462."""
import os

import numpy as np
import pytest

from occ_b200 import aobasis, molecules


# ---------------------------------------------------------------- collocation vs the reference's gau2grid
@pytest.mark.parametrize("pure", [0, 1])
@pytest.mark.parametrize("l", [0, 1, 2, 3, 4])
def test_collocation_matches_reference_gau2grid(oracle, l, pure):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libocc_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(100 + l)
    pts = rng.normal(size=(67, 3)) * 1.5
    exps = np.array([4.1, 1.3, 0.35])
    coefs = rng.normal(size=3)
    center = np.array([0.1, -0.2, 0.3])
    for deriv in (0, 1):
        ours = oracle.collocate_shell(l, pure, pts, coefs, exps, center, deriv)
        ref = oracle.ref_collocate_shell(l, pure, pts, coefs, exps, center, deriv)
        for a, b in zip(ours, ref):
            scale = max(1.0, np.abs(b).max())
            assert np.abs(a - b).max() <= 2e-14 * scale


def test_shell_normalisation_matches_host_mirror(oracle):
    """src/gto/shell.cpp:20-64,180-186 restated twice (oracle C++, occ_b200.aobasis Python)"""
    exps = np.array([130.70932, 23.808861, 6.4436083])
    c = np.array([0.15432897, 0.53532814, 0.44463454])
    for l in range(4):
        a = oracle.incorporate_shell_norm(l, exps, c)
        b = aobasis.incorporate_shell_norm(l, exps, c)
        assert np.allclose(a, b, rtol=1e-14, atol=0)


# ---------------------------------------------------------------- functionals
def test_pbe_exchange_known_answers(oracle):
    """tests/dft_tests.cpp:71-111 (restricted): exc of gga_x_pbe at four (rho, grad rho) points, 1e-6"""
    rho = np.array([0.030097001731591706, 0.030097001731591706, 0.15085916775597902, 0.17751228531944346])
    grad = np.array([[-0.063735860841572081, 0.0, 0.018656553309952275], [-0.063735860841572081, 0.0, 0.018656553309952275],
                     [0.0, 0.0, 0.13035149660035189], [0.0, 0.0, 0.078426011080503066]])
    sigma = (grad ** 2).sum(1)
    r = oracle.xc_eval("gga_x_pbe", False, rho, sigma)
    expected = np.array([-0.27851489, -0.27851489, -0.39899553, -0.41654061])
    assert np.abs(r["exc"] - expected).max() < 1e-6


# tests/dft_tests.cpp:138-191: rho, grad rho (3), laplacian, tau at four points; TPSS known answers (1e-6)
TPSS_KAT_ROWS = np.array([
    [0.06019401627252133, -0.12747174847145987, 0.0, 0.03731311973522952, 0.09472176570942804, 0.036634279981580156],
    [0.06019401627252133, 0.0, -0.12747174847145987, 0.03731311973522952, 0.09472176570942804, 0.036634279981580156],
    [0.30171849914696935, 0.0, 0.0, 0.26070301592021455, -2.594981941107205, 0.02815789498408782],
    [0.3550246251878335, 0.0, 0.0, 0.15685208520532756, -6.307746108179051, 0.008662278222335566]])
TPSS_KAT = {
    "exc_x": [-0.33175882, -0.33175882, -0.56154382, -0.59257418],
    "exc_c": [-0.02276044, -0.02276044, -0.04168478, -0.04268936],
    "vrho_c": [-5.57565777e-02, -5.57565777e-02, -8.00593613e-02, -8.53775165e-02],
    "vsigma_c": [3.91761083e-02, 3.91761083e-02, 1.47458233e-01, 5.51608074e-01],
    "vtau_c": [1.10283011e-03, 1.10283011e-03, -3.50002323e-01, -1.56487879e+00],
}


def tpss_kat_inputs():
    rho = TPSS_KAT_ROWS[:, 0].copy()
    sigma = (TPSS_KAT_ROWS[:, 1:4] ** 2).sum(1)
    tau = TPSS_KAT_ROWS[:, 5].copy()
    return rho, sigma, tau


def test_tpss_known_answers_energies_and_potentials(oracle):
    """tests/dft_tests.cpp:138-191: mgga_x_tpss eps; mgga_c_tpss eps AND vrho / vsigma / vtau — the one place the
    reference pins libxc POTENTIALS, so it also pins the dual-number derivative machinery of the oracle"""
    rho, sigma, tau = tpss_kat_inputs()
    x = oracle.xc_eval("mgga_x_tpss", False, rho, sigma, tau)
    c = oracle.xc_eval("mgga_c_tpss", False, rho, sigma, tau)
    assert np.abs(x["exc"] - TPSS_KAT["exc_x"]).max() < 1e-6
    assert np.abs(c["exc"] - TPSS_KAT["exc_c"]).max() < 1e-6
    assert np.abs(c["vrho"][:, 0] - TPSS_KAT["vrho_c"]).max() < 1e-6
    assert np.abs(c["vsigma"][:, 0] - TPSS_KAT["vsigma_c"]).max() < 1e-6
    assert np.abs(c["vtau"][:, 0] - TPSS_KAT["vtau_c"]).max() < 1e-6


FUNCS = ["lda_x", "lda_c_vwn", "lda_c_vwn_rpa", "lda_c_vwn_3", "gga_x_pbe", "gga_x_b88", "gga_c_pbe", "gga_c_lyp",
         "hyb_gga_xc_b3lyp", "mgga_x_r2scan", "mgga_c_r2scan", "lda_c_pw", "lda_c_pw_mod", "mgga_x_tpss",
         "mgga_c_tpss", "hyb_gga_xc_b3lyp5", "hyb_gga_xc_pbeh", "hyb_gga_xc_bhandh", "hyb_gga_xc_bhandhlyp",
         "hyb_gga_xc_b97", "hyb_gga_xc_b3pw91", "gga_c_pw91", "gga_c_p86", "lda_c_pz", "hyb_gga_xc_wb97x",
         "hyb_gga_xc_wb97x_v", "hyb_gga_xc_b3p86", "gga_x_b86_mgc"]


def _sample(n, seed, polarized):
    rng = np.random.default_rng(seed)
    s = 2 if polarized else 1
    rho = 10 ** rng.uniform(-3, 0.5, size=(n, s))
    g = rng.normal(size=(n, s, 3))
    g *= (rng.uniform(0.3, 2.0, size=(n, s)) * rho ** (4 / 3) / np.linalg.norm(g, axis=2))[..., None]
    if polarized:
        sigma = np.stack([(g[:, 0] ** 2).sum(1), (g[:, 0] * g[:, 1]).sum(1), (g[:, 1] ** 2).sum(1)], 1)
    else:
        sigma = (g[:, 0] ** 2).sum(1)[:, None]
    # tau above the von Weizsaecker bound so the Fermi-hole clamp is inactive
    tw = np.stack([(g[:, k] ** 2).sum(1) / (8 * rho[:, k]) for k in range(s)], 1)
    tau = tw * (1.0 + rng.uniform(0.05, 3.0, size=(n, s)))
    return rho, sigma, tau


@pytest.mark.parametrize("func", FUNCS)
def test_potentials_are_derivatives_of_the_energy_density(oracle, func):
    """vrho/vsigma/vtau vs central finite differences of rho*exc (unpolarized) — the potentials are not pinned by
    any reference test, so they are secured against the pinned energies"""
    rho, sigma, tau = _sample(40, 7, False)
    r0 = oracle.xc_eval(func, False, rho, sigma, tau)
    fam = r0["family"]

    def F(rh, sg, ta):
        return oracle.xc_eval(func, False, rh, sg, ta)["exc"] * rh[:, 0]

    h = 1e-5
    d = (F(rho * (1 + h), sigma, tau) - F(rho * (1 - h), sigma, tau)) / (2 * h * rho[:, 0])
    assert np.allclose(d, r0["vrho"][:, 0], rtol=2e-6, atol=1e-9)
    if fam >= 1:
        d = (F(rho, sigma * (1 + h), tau) - F(rho, sigma * (1 - h), tau)) / (2 * h * sigma[:, 0])
        assert np.allclose(d, r0["vsigma"][:, 0], rtol=2e-5, atol=1e-9)
    if fam >= 2:
        d = (F(rho, sigma, tau * (1 + h)) - F(rho, sigma, tau * (1 - h))) / (2 * h * tau[:, 0])
        assert np.allclose(d, r0["vtau"][:, 0], rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("func", FUNCS)
def test_polarized_at_zero_polarisation_equals_unpolarized(oracle, func):
    rho, sigma, tau = _sample(40, 11, False)
    ru = oracle.xc_eval(func, False, rho, sigma, tau)
    rp = oracle.xc_eval(func, True, np.repeat(rho / 2, 2, 1), np.repeat(sigma / 4, 3, 1), np.repeat(tau / 2, 2, 1))
    assert np.allclose(rp["exc"], ru["exc"], rtol=1e-12, atol=1e-14)
    assert np.allclose(rp["vrho"][:, 0], ru["vrho"][:, 0], rtol=1e-10, atol=1e-12)
    assert np.allclose(rp["vrho"][:, 1], ru["vrho"][:, 0], rtol=1e-10, atol=1e-12)
    if ru["family"] >= 1:
        # d/d sigma_total = (vaa + vab + vbb) / 4 * ... : total-sigma derivative from the spin-resolved ones
        vs = (rp["vsigma"][:, 0] + rp["vsigma"][:, 1] + rp["vsigma"][:, 2]) / 4.0
        assert np.allclose(vs, ru["vsigma"][:, 0], rtol=1e-9, atol=1e-12)
    if ru["family"] >= 2:
        assert np.allclose(rp["vtau"][:, 0], ru["vtau"][:, 0], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("func", FUNCS)
def test_polarized_potentials_by_finite_differences(oracle, func):
    rho, sigma, tau = _sample(30, 13, True)
    r0 = oracle.xc_eval(func, True, rho, sigma, tau)
    fam = r0["family"]

    def F(rh, sg, ta):
        return oracle.xc_eval(func, True, rh, sg, ta)["exc"] * rh.sum(1)

    h = 1e-5
    for k in range(2):
        up, dn = rho.copy(), rho.copy()
        up[:, k] *= 1 + h
        dn[:, k] *= 1 - h
        d = (F(up, sigma, tau) - F(dn, sigma, tau)) / (2 * h * rho[:, k])
        assert np.allclose(d, r0["vrho"][:, k], rtol=5e-6, atol=1e-8)
    if fam >= 1:
        for k in range(3):
            up, dn = sigma.copy(), sigma.copy()
            step = h * (sigma[:, k] if k != 1 else 0.5 * (sigma[:, 0] + sigma[:, 2]))
            up[:, k] += step
            dn[:, k] -= step
            d = (F(rho, up, tau) - F(rho, dn, tau)) / (2 * step)
            # FD round-off floor: a few ulp of F divided by the step (matters when one spin channel is tiny)
            noise = 3e-14 * np.abs(F(rho, sigma, tau)) / step
            assert (np.abs(d - r0["vsigma"][:, k]) <= 5e-5 * np.abs(r0["vsigma"][:, k]) + 1e-8 + noise).all()
    if fam >= 2:
        for k in range(2):
            up, dn = tau.copy(), tau.copy()
            up[:, k] *= 1 + h
            dn[:, k] *= 1 - h
            d = (F(rho, sigma, up) - F(rho, sigma, dn)) / (2 * h * tau[:, k])
            assert np.allclose(d, r0["vtau"][:, k], rtol=5e-5, atol=1e-8)


@pytest.mark.parametrize("method,nspin", [("svwn", 1), ("pbe", 1), ("pbe", 2), ("tpss", 1), ("tpss", 2), ("b3lyp", 1),
                                          ("b3lyp", 2), ("wb97x", 1), ("wb97x", 2), ("r2scan", 1)])
def test_vxc_matrix_is_the_derivative_of_exc_with_respect_to_the_density_matrix(oracle, case, method, nspin):
    """dExc = sum_ij K_ij dD2_ij for a symmetric perturbation of the (doubled) density matrix, restricted and unrestricted:
    the ASSEMBLY of the Vxc matrix (quadrature weights, the 1/2 and 2 of the gradient terms, K = Ktmp + Ktmp^T, the tau
    term, the shell screening) is pinned against the energies, which the reference's goldens pin.  An error common to
    the oracle and the CUDA path in that assembly would break this identity."""
    c = case("water", "def2-svp")
    b = c["basis"]
    n = b.nbf
    D2 = molecules.synthetic_density(b, nspin)
    rng = np.random.default_rng(3)
    dD = rng.normal(size=(nspin * n, n))
    for s in range(nspin):
        dD[s * n:(s + 1) * n] = 0.5 * (dD[s * n:(s + 1) * n] + dD[s * n:(s + 1) * n].T)
    dD *= 1e-3 * np.abs(D2).max()
    names = oracle.METHODS[method]

    def build(D):
        return oracle.xc_build(c["obasis"], c["spts"], c["sw"], c["leaves"], names, nspin, np.asfortranarray(D))

    eps = 1e-3
    fd = (build(D2 + eps * dD)["exc"] - build(D2 - eps * dD)["exc"]) / (2 * eps)
    an = float((build(D2)["K"] * dD).sum())
    assert abs(fd - an) < 2e-8 * abs(an), (fd, an)


def test_b3lyp_is_its_libxc_mixture(oracle):
    rho, sigma, tau = _sample(25, 5, False)
    mix = {"lda_x": 0.08, "gga_x_b88": 0.72, "lda_c_vwn_rpa": 0.19, "gga_c_lyp": 0.81}
    tot = oracle.xc_eval("hyb_gga_xc_b3lyp", False, rho, sigma)
    acc = sum(c * oracle.xc_eval(f, False, rho, sigma)["exc"] for f, c in mix.items())
    assert np.allclose(tot["exc"], acc, rtol=1e-13)
    assert oracle.lib().orc_xc_exx(402) == 0.20


def test_unknown_functional_raises(oracle):
    with pytest.raises(RuntimeError):
        oracle.xc_eval(12345, False, np.ones(2))


# ---------------------------------------------------------------- grids
def test_lebedev_tables_are_the_references(oracle):
    t = np.load(os.path.join(os.path.dirname(__file__), "golden", "lebedev_tables.npz"))
    assert {"n110", "n302", "n266", "n590", "n50", "n74"} <= set(t.files)
    for k in t.files:
        a = t[k]
        assert a.shape == (int(k[1:]), 4)       # tests/dft_tests.cpp:196-200 pins the 110 shape
        assert abs(a[:, 3].sum() - 1.0) < 1e-13
        assert np.allclose((a[:, :3] ** 2).sum(1), 1.0, atol=1e-14)
    if oracle.ref_lib() is not None:
        import ctypes
        R = oracle.ref_lib()
        for n in (110, 302):
            a = np.zeros((n, 4))
            R.occ_ref_lebedev(ctypes.c_int(n), a.ctypes.data_as(ctypes.c_void_p))
            assert np.array_equal(a, t["n%d" % n])


def test_veryfine_water_grid_size_quoted_by_the_reference(oracle):
    """tests/isolated_xc_tests.cpp:271: 38 652 points for water/3-21G (spherical) at GridQuality::VeryFine"""
    Z, pos = molecules.water_isolated_xc()
    b = aobasis.AOBasis.load(Z, pos, "3-21g", pure=True)
    ob = oracle.Basis(b.as_arrays())
    s = oracle.default_grid_settings(max_angular_points=590, min_angular_points=302)  # grid_settings.cpp:50-54
    pts, w, info = oracle.molecular_grid(Z, pos, ob, s)
    assert len(w) == 38652
    assert info["atom_npts"] == [13732, 12460, 12460]
    assert info["atom_nradial"] == [86, 74, 74]


def test_default_grid_sizes_of_the_baseline_configs(oracle):
    for mol, bas, n in (("water", "def2-svp", 30172), ("benzene", "def2-tzvp", 139236)):
        Z, pos = molecules.MOLECULES[mol]()
        b = aobasis.AOBasis.load(Z, pos, bas)
        pts, w, info = oracle.molecular_grid(Z, pos, oracle.Basis(b.as_arrays()))
        assert len(w) == n


def test_morton_hierarchy_invariants(oracle, case):
    """tests/dft_tests.cpp:1287-1462: all points preserved, sum of weights preserved, leaf sizes, spheres contain points"""
    c = case("water", "def2-svp")
    spts, sw, perm, leaves = c["spts"], c["sw"], c["perm"], c["leaves"]
    assert sorted(perm.tolist()) == list(range(len(sw)))
    assert np.array_equal(spts, c["pts"][perm.astype(np.int64)])
    assert abs(sw.sum() - c["w"].sum()) < 1e-12 * abs(c["w"].sum())
    assert leaves["count"].sum() == len(sw)
    assert (leaves["count"][:-1] == 128).all() and 1 <= leaves["count"][-1] <= 256
    assert (leaves["offset"] == np.concatenate([[0], np.cumsum(leaves["count"])[:-1]])).all()
    for lf in leaves[:: max(1, len(leaves) // 40)]:
        p = spts[int(lf["offset"]): int(lf["offset"] + lf["count"])]
        assert np.allclose(p.mean(0), lf["center"], atol=1e-12)
        assert np.linalg.norm(p - lf["center"], axis=1).max() <= lf["radius"] + 1e-10


def test_basis_function_counts_of_the_baseline_configs():
    for mol, bas, cart, sph in (("water", "def2-svp", 25, 24), ("benzene", "def2-tzvp", 252, 222),
                                ("h2o_32", "def2-svp", 800, 768), ("valinomycin", "def2-tzvp", 3348, 2958),
                                ("taxol", "def2-tzvp", 2538, 2228)):
        Z, pos = molecules.MOLECULES[mol]()
        assert aobasis.AOBasis.load(Z, pos, bas).nbf == cart
        assert aobasis.AOBasis.load(Z, pos, bas, pure=True).nbf == sph


# ---------------------------------------------------------------- polarized branches: the hydrogen atom
def hydrogen_atom_quadrature():
    """rho_alpha = exp(-2r)/pi, rho_beta = 0 on a radial Gauss-Legendre grid: sigma_aa = 4 rho^2, tau = tau_W = rho/2"""
    x, wq = np.polynomial.legendre.leggauss(400)
    r, w = 20.0 * (x + 1.0), 20.0 * wq
    ra = np.exp(-2.0 * r) / np.pi
    z = np.zeros_like(ra)
    return (np.stack([ra, z], 1), np.stack([4.0 * ra ** 2, z, z], 1), np.stack([0.5 * ra, z], 1),
            ra * 4.0 * np.pi * r ** 2 * w)


# Published values for the (fully spin-polarized, one-electron) hydrogen atom.  LSD / PBE: Table I of Perdew, Burke,
# Ernzerhof, PRL 77, 3865 (1996): Ex(LSD) -0.2680, Ex(PBE) -0.3059, Ec(LSD) -0.0222, Ec(PBE) -0.0060; B88: Becke 1988,
# -0.310.  Exact constraints: LYP, TPSS-c and r2SCAN-c are one-electron self-correlation free (Ec = 0); TPSS-x and
# (r2)SCAN-x reproduce the exact hydrogen exchange energy -0.3125 by construction.
HYDROGEN = {"lda_x": (-0.2680, 5e-5), "gga_x_pbe": (-0.3059, 5e-5), "gga_x_b88": (-0.310, 5e-4),
            "lda_c_pw_mod": (-0.0222, 5e-5), "gga_c_pbe": (-0.0060, 5e-5), "gga_c_lyp": (0.0, 1e-10),
            "mgga_x_tpss": (-0.3125, 2e-6), "mgga_c_tpss": (0.0, 1e-10), "mgga_x_r2scan": (-0.3125, 5e-6),
            "mgga_c_r2scan": (0.0, 1e-10),
            # PW91 correlation: Table I of the PBE paper lists -0.0065 for hydrogen (this form gives -0.00659);
            # P86: -0.002 .. -0.003 in Table I of Perdew, Phys. Rev. B 33, 8822 (1986)
            "gga_c_pw91": (-0.0065, 1.5e-4), "gga_c_p86": (-0.0025, 6e-4)}


@pytest.mark.parametrize("func", sorted(HYDROGEN))
def test_polarized_branches_on_the_hydrogen_atom(oracle, func):
    """No reference test pins a POLARIZED functional value; the hydrogen atom does, through published numbers and the
    exact one-electron constraints the functionals were built to satisfy"""
    rho, sigma, tau, qw = hydrogen_atom_quadrature()
    o = oracle.xc_eval(func, True, rho, sigma, tau)
    want, tol = HYDROGEN[func]
    assert abs(float((o["exc"] * qw).sum()) - want) < tol


def test_vwn_ferromagnetic_limits_on_the_hydrogen_atom(oracle):
    """zeta = 1: VWN5 and VWN3 both reduce to the Monte-Carlo ferromagnetic fit (identical), which fits the same
    Ceperley-Alder data as PW92's ferromagnetic branch (pinned above by PBE's Table I): agreement to 1e-4 Eh"""
    rho, sigma, tau, qw = hydrogen_atom_quadrature()
    e = {f: float((oracle.xc_eval(f, True, rho, sigma, tau)["exc"] * qw).sum())
         for f in ("lda_c_vwn", "lda_c_vwn_3", "lda_c_pw_mod", "lda_c_pw")}
    assert abs(e["lda_c_vwn"] - e["lda_c_vwn_3"]) < 1e-12
    assert abs(e["lda_c_vwn"] - e["lda_c_pw_mod"]) < 1e-4
    assert abs(e["lda_c_pw"] - e["lda_c_pw_mod"]) < 1e-6


def test_unpolarized_lda_correlation_fits_agree_with_ceperley_alder(oracle):
    """uniform gas, zeta = 0: eps_c from Ceperley-Alder QMC (Ry -> Ha): rs = 1, 2, 5, 10 -> -0.0598, -0.0451, -0.0282,
    -0.0186; VWN5 and PW92 are both fits of these data (5e-4), and agree with each other to 3e-4 (PW92 at rs = 2
    is the -0.04476 of its own Table)"""
    rs = np.array([1.0, 2.0, 5.0, 10.0])
    rho = 3.0 / (4.0 * np.pi * rs ** 3)
    ca = np.array([-0.0598, -0.0451, -0.0282, -0.0186])
    pw = oracle.xc_eval("lda_c_pw_mod", False, rho)["exc"]
    vwn = oracle.xc_eval("lda_c_vwn", False, rho)["exc"]
    assert np.abs(pw - ca).max() < 5e-4 and np.abs(vwn - ca).max() < 5e-4
    assert np.abs(pw - vwn).max() < 3e-4
    # VWN3 == VWN5 at zeta = 0; the RPA fit lies below both (RPA overcorrelates)
    assert np.abs(oracle.xc_eval("lda_c_vwn_3", False, rho)["exc"] - vwn).max() < 1e-14
    assert (oracle.xc_eval("lda_c_vwn_rpa", False, rho)["exc"] < vwn - 0.01).all()


# ---------------------------------------------------------------- published atomic energies (He, Li, Be)
# Table I of Perdew, Burke, Ernzerhof, PRL 77, 3865 (1996) (LSD = Slater + PW92, PBE), Becke, PRA 38, 3098 (1988) (B88)
# Lee, Yang, Parr / Miehlich et al. (LYP) and Table II of Tao, Perdew, Staroverov, Scuseria, PRL 91, 146401 (2003)
# (TPSS correlation: its zeta / xi dependent C(zeta, xi) is exercised by lithium), all on Hartree-Fock densities.  Lithium is spin-polarized with
# 0 < zeta < 1 in the valence region: it pins the zeta interpolation of PW92 and phi(zeta) of PBE correlation, the
# spin-scaling of the exchange functionals and the open-shell LYP form — none of which any reference test covers.
ATOMS = {
    "He": (2, 1, 1, {"lda_x": -0.8840, "gga_x_pbe": -1.0136, "lda_c_pw_mod": -0.1125, "gga_c_pbe": -0.0420,
                     "gga_x_b88": -1.0255, "gga_c_lyp": -0.0438, "mgga_c_tpss": -0.0430}),
    "Li": (3, 2, 1, {"lda_x": -1.5379, "gga_x_pbe": -1.7572, "lda_c_pw_mod": -0.1508, "gga_c_pbe": -0.0514,
                     "gga_x_b88": -1.7753, "gga_c_lyp": -0.0534, "mgga_c_tpss": -0.0493}),
    "Be": (4, 2, 2, {"lda_x": -2.3124, "gga_x_pbe": -2.6358, "lda_c_pw_mod": -0.2240, "gga_c_pbe": -0.0856,
                     "gga_x_b88": -2.6578, "gga_c_lyp": -0.0946, "mgga_c_tpss": -0.0869}),
}


def atom_quadrature(name):
    import atoms_uhf
    Z, na, nb, _ = ATOMS[name]
    res = atoms_uhf.uhf(Z, na, nb)
    x, wq = np.polynomial.legendre.leggauss(600)
    r, w = (1 + x) / (1 - x), wq * 2 / (1 - x) ** 2
    rho, sigma, tau = atoms_uhf.radial_densities(res, r)
    return res, rho, sigma, tau, 4 * np.pi * r ** 2 * w


# Correlation energies printed to three digits in the original papers: PW91 (Perdew et al., Phys. Rev. B 46, 6671 (1992),
# Table II: He 0.046, Be 0.094) and P86 (Perdew, Phys. Rev. B 33, 8822 (1986), Table I: He 0.044, Be 0.094).  No reference
# test holds a value of either functional (bp86 / b3pw91 only appear in the method table, src/dft/dft_method.cpp:16,30).
ATOMS_3DIGIT = {"He": {"gga_c_pw91": -0.046, "gga_c_p86": -0.044}, "Be": {"gga_c_pw91": -0.094, "gga_c_p86": -0.094}}


@pytest.mark.parametrize("atom", sorted(ATOMS_3DIGIT))
def test_pw91_and_p86_correlation_energies_of_closed_shell_atoms(oracle, atom):
    res, rho, sigma, tau, vol = atom_quadrature(atom)
    for func, want in ATOMS_3DIGIT[atom].items():
        o = oracle.xc_eval(func, True, rho, sigma, tau)
        got = float((o["exc"] * rho.sum(1) * vol).sum())
        assert abs(got - want) < 6e-4, (atom, func, got, want)
    # P86 = PZ81 + gradient term; PZ81 agrees with the PW92 fit of the same Ceperley-Alder data to 1e-3 on these atoms
    pz = oracle.xc_eval("lda_c_pz", True, rho, sigma, tau)
    pw = oracle.xc_eval("lda_c_pw", True, rho, sigma, tau)
    assert abs(float(((pz["exc"] - pw["exc"]) * rho.sum(1) * vol).sum())) < 1e-3


def test_b97_and_b3pw91_are_their_published_mixtures(oracle):
    """B97 shares every line with the pinned wB97X except the omega = 0 exchange and its nine coefficients; its leading
    terms reduce to 0.8094 LSDA exchange + (0.1737 same-spin, 0.9454 opposite-spin) PW92 at zero gradient.  B3PW91 is
    libxc's 0.08 / 0.72 / 0.19 / 0.81 mixture."""
    rho, sigma, tau = _sample(30, 5, True)
    z = np.zeros_like(sigma)
    b97 = oracle.xc_eval("hyb_gga_xc_b97", True, rho, z, tau)["exc"]
    x = oracle.xc_eval("lda_x", True, rho, z, tau)["exc"]
    pw = oracle.xc_eval("lda_c_pw", True, rho, z, tau)["exc"] * rho.sum(1)
    ferro = [oracle.xc_eval("lda_c_pw", True, np.stack([rho[:, k], np.zeros(len(rho))], 1), z, tau)["exc"] * rho[:, k]
             for k in range(2)]
    want = 0.8094 * x * rho.sum(1) + 0.1737 * (ferro[0] + ferro[1]) + 0.9454 * (pw - ferro[0] - ferro[1])
    assert np.allclose(b97 * rho.sum(1), want, rtol=1e-12, atol=1e-14)
    mix = oracle.xc_eval("hyb_gga_xc_b3pw91", True, rho, sigma, tau)
    parts = [oracle.xc_eval(n, True, rho, sigma, tau) for n in ("lda_x", "gga_x_b88", "lda_c_pw", "gga_c_pw91")]
    for k in ("exc", "vrho", "vsigma"):
        assert np.allclose(mix[k], sum(c * q[k] for c, q in zip((0.08, 0.72, 0.19, 0.81), parts)), rtol=1e-12, atol=1e-14)
    assert oracle.lib().orc_xc_exx(407) == pytest.approx(0.1943) and oracle.lib().orc_xc_exx(401) == pytest.approx(0.20)
    # libxc hyb_gga_xc_b3p86: the same weights over {lda_x, gga_x_b88, lda_c_vwn_rpa, gga_c_p86}
    mix = oracle.xc_eval("hyb_gga_xc_b3p86", True, rho, sigma, tau)
    parts = [oracle.xc_eval(n, True, rho, sigma, tau) for n in ("lda_x", "gga_x_b88", "lda_c_vwn_rpa", "gga_c_p86")]
    for k in ("exc", "vrho", "vsigma"):
        assert np.allclose(mix[k], sum(c * q[k] for c, q in zip((0.08, 0.72, 0.19, 0.81), parts)), rtol=1e-12, atol=1e-14)
    assert oracle.lib().orc_xc_exx(403) == pytest.approx(0.20) and oracle.lib().orc_xc_exx(466) == pytest.approx(1.0)


def test_b86b_exchange_limits_and_enhancement_factor(oracle):
    """gga_x_b86_mgc (Becke 1986b): F(x) = 1 + (beta / C_x) x^2 / (1 + gamma x^2)^(4/5); LSDA exchange at zero gradient,
    the closed form at finite gradient, and the small-gradient coefficient against B88's (0.0042) ratio"""
    rho, sigma, tau = _sample(30, 11, True)
    lda = oracle.xc_eval("lda_x", True, rho, 0 * sigma, tau)["exc"]
    assert np.allclose(oracle.xc_eval("gga_x_b86_mgc", True, rho, 0 * sigma, tau)["exc"], lda, rtol=1e-14)
    got = oracle.xc_eval("gga_x_b86_mgc", True, rho, sigma, tau)["exc"] * rho.sum(1)
    cx = 0.75 * (6.0 / np.pi) ** (1.0 / 3.0)
    want = 0.0
    for k, sk in ((0, 0), (1, 2)):
        r43 = rho[:, k] ** (4.0 / 3.0)
        x2 = sigma[:, sk] / r43 ** 2
        want = want - r43 * (cx + 0.00375 * x2 / (1.0 + 0.007 * x2) ** 0.8)
    assert np.allclose(got, want, rtol=1e-13)


@pytest.mark.parametrize("atom", sorted(ATOMS))
def test_published_atomic_exchange_and_correlation_energies(oracle, atom):
    res, rho, sigma, tau, vol = atom_quadrature(atom)
    hf_limit = {"He": -2.86168, "Li": -7.43273, "Be": -14.57302}[atom]
    assert abs(res["energy"] - hf_limit) < 5e-5                     # the helper's own density is at the HF limit
    assert abs(float((rho.sum(1) * vol).sum()) - ATOMS[atom][0]) < 1e-9
    for func, want in ATOMS[atom][3].items():
        o = oracle.xc_eval(func, True, rho, sigma, tau)
        got = float((o["exc"] * rho.sum(1) * vol).sum())
        assert abs(got - want) < 1.5e-4, (atom, func, got, want)
