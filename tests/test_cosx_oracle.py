"""Pins of the COSX oracle (oracle/src/cosx.cpp), SURVEY 8(f) rank 2.

The reference holds no golden K for SemiNumericalExchange (tests/dft_tests.cpp:322-369 only prints |K - K_exact|), so
the restatement is pinned piece by piece:
  * Boys table / E coefficients / Hermite Coulomb integrals: against the reference's own header-only code
    (include/occ/ints/{boys,ecoeffs,rints}.h) compiled into oracle/_ref/libocc_ref.so;
  * ESP integrals: closed forms for s Gaussians (full Coulomb and erf-attenuated) and numerical quadrature of the
    collocated AO products (the same AOs the XC path is pinned on) for every l <= 3, Cartesian and spherical;
  * analytic overlap: against quadrature (the reference prints the same "Max error |Sn - S|");
  * assembled K: against the exact exchange matrix of an s-only basis (closed-form two-electron integrals), the
    comparison the reference's own test makes.
"""
import math

import numpy as np
import pytest

from occ_b200 import aobasis, molecules


def _ref_or_skip(O):
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return O


def test_boys_matches_reference_headers(oracle):
    O = _ref_or_skip(oracle)
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(0, 36, 300)) + list(rng.uniform(36, 300, 50)) + [0.0, 1e-9, 35.999999, 36.0, 36.000001]
    worst = 0.0
    for x in xs:
        for m in range(13):
            a, b = O.cosx_boys(x, m), O.ref_boys(x, m)
            worst = max(worst, abs(a - b) / abs(b))
    assert worst < 1e-14
    # the table against the Maclaurin / erf reference values (tests/ints_tests.cpp:40-60 use 1e-8 / 1e-14)
    for x in rng.uniform(0, 36, 100):
        for m in range(8):
            assert O.cosx_boys(x, m) == pytest.approx(O.cosx_boys_reference(x, m), rel=2e-13)


def test_ecoeffs_match_reference_headers(oracle):
    O = _ref_or_skip(oracle)
    rng = np.random.default_rng(1)
    for _ in range(200):
        la, lb = (int(v) for v in rng.integers(0, 6, 2))
        a, b = rng.uniform(0.05, 30, 2)
        X = rng.uniform(-4, 4)
        e1, e2 = O.cosx_ecoeffs_1d(la, lb, a, b, X), O.ref_ecoeffs_1d(la, lb, a, b, X)
        assert np.abs(e1 - e2).max() <= 1e-15 * max(1.0, np.abs(e2).max())
    # tests/ints_tests.cpp:69-101: E_00^0 = exp(-mu X^2), E_10^0 = X_PA E_00^0
    a, b, X = 1.3, 0.7, 0.9
    E = O.cosx_ecoeffs_1d(1, 0, a, b, X)
    e00 = math.exp(-a * b / (a + b) * X * X)
    assert E[0, 0, 0] == pytest.approx(e00, rel=1e-14)
    assert E[1, 0, 0] == pytest.approx(b / (a + b) * X * e00, rel=1e-14)


def test_rints_match_reference_headers(oracle):
    O = _ref_or_skip(oracle)
    rng = np.random.default_rng(2)
    for _ in range(300):
        L = int(rng.integers(0, 9))
        p = rng.uniform(0.05, 40)
        PC = rng.uniform(-3, 3, 3)
        r1, r2 = O.cosx_rints(L, p, PC), O.ref_rints(L, p, PC)
        assert np.abs(r1 - r2).max() <= 5e-13 * np.abs(r2).max()  # (2p)^m amplifies the last-ulp FMA differences of F_m
    # tests/ints_tests.cpp:104-128: R_000 = F_0(p |PC|^2)
    assert O.cosx_rints(0, 1.0, [0, 0, 0])[0] == pytest.approx(1.0, rel=1e-13)


def test_esp_s_gaussian_closed_forms(oracle):
    O = oracle
    for alpha in (0.3, 1.5, 20.0):
        c = O.incorporate_shell_norm(0, [alpha], [1.0])
        pts = np.array([[0.0, 0, 0], [1.0, 0, 0], [0.3, -2.0, 1.1], [9.0, 0, 0]])
        R = np.linalg.norm(pts, axis=1)
        V = O.cosx_esp_pair(0, 0, [alpha], [alpha], c, c, [0, 0, 0], [0, 0, 0], False, pts).ravel()
        p = 2 * alpha
        exact = np.array([2 * math.sqrt(p / math.pi)] + [math.erf(math.sqrt(p) * r) / r for r in R[1:]])
        assert np.abs(V - exact).max() < 5e-14
        # long-range operator erf(omega r)/r  (rints_specialized.h:1462-1500)
        omega = 0.3
        Vlr = O.cosx_esp_pair(0, 0, [alpha], [alpha], c, c, [0, 0, 0], [0, 0, 0], False, pts, omega=omega).ravel()
        pe = p * omega ** 2 / (p + omega ** 2)
        exact = np.array([2 * math.sqrt(pe / math.pi)] + [math.erf(math.sqrt(pe) * r) / r for r in R[1:]])
        assert np.abs(Vlr - exact).max() < 5e-14


def _two_centre_basis(pure, ls=range(4)):
    Z = np.array([6, 8], dtype=np.int32)
    pos = np.array([[0.0, 0.0, 0.0], [1.1, -0.7, 0.9]])
    b = aobasis.AOBasis(Z, pos, [], pure, "test")
    rng = np.random.default_rng(5)
    for a in range(2):
        for l in ls:
            exps = np.array([1.9 - 0.3 * l + 0.2 * a, 0.6 + 0.1 * l])
            coefs = rng.uniform(0.3, 1.0, 2)
            b.shells.append(aobasis.Shell(l, pos[a].copy(), exps, aobasis.incorporate_shell_norm(l, exps, coefs), a))
    return Z, pos, b


@pytest.mark.parametrize("pure", [False, True])
def test_esp_matches_quadrature_of_collocated_aos(oracle, pure):
    """V_ab(C) = int phi_a phi_b / |r - C|: ESP integrals against quadrature of the oracle's collocation (which is
    pinned to the reference's gau2grid) on a fine Becke grid; C outside the charge cloud. Covers every (la, lb) with
    l <= 3, both orderings, one- and two-centre pairs, the c2s convention of gto.cpp:289-400 and common_fac."""
    O = oracle
    Z, pos, b = _two_centre_basis(pure)
    ob = O.Basis(b.as_arrays())
    pts, w, _ = O.molecular_grid(Z, pos, ob, O.default_grid_settings(max_angular_points=590, min_angular_points=590,
                                                                    reduced_first_row_element_grid=0))
    nbf = ob.nbf
    phi = np.zeros((len(w), nbf))
    fb = b.first_bf
    for s, sh in enumerate(b.shells):
        v = O.collocate_shell(sh.l, int(pure), pts, sh.coefficients, sh.exponents, sh.origin, 0)[0]
        n = v.shape[0]
        phi[:, fb[s]:fb[s] + n] = v.T * O.lib().orc_common_fac(sh.l, int(pure))
    Cs = np.array([[7.0, 5.0, -6.0], [-8.0, 6.5, 4.0]])
    for C in Cs:
        rinv = 1.0 / np.linalg.norm(pts - C, axis=1)
        Vq = phi.T @ (phi * (w * rinv)[:, None])
        for p, sp in enumerate(b.shells):
            for q, sq in enumerate(b.shells):
                V = O.cosx_esp_pair(sp.l, sq.l, sp.exponents, sq.exponents, sp.coefficients, sq.coefficients, sp.origin,
                                    sq.origin, pure, C[None, :])[0]
                blk = Vq[fb[p]:fb[p] + V.shape[0], fb[q]:fb[q] + V.shape[1]]
                assert np.abs(V - blk).max() < 2e-8, (p, q, np.abs(V - blk).max())


@pytest.mark.parametrize("pure", [False, True])
def test_esp_g_shells_match_quadrature(oracle, pure):
    """the same check for g shells (COSX handles def2-QZVP): (4, lb) and (la, 4) blocks, one- and two-centre, up to L = 8"""
    O = oracle
    Z, pos, b = _two_centre_basis(pure, ls=(0, 2, 4))
    ob = O.Basis(b.as_arrays())
    pts, w, _ = O.molecular_grid(Z, pos, ob, O.default_grid_settings(max_angular_points=590, min_angular_points=590,
                                                                    reduced_first_row_element_grid=0))
    phi = np.zeros((len(w), ob.nbf))
    fb = b.first_bf
    for s, sh in enumerate(b.shells):
        v = O.collocate_shell(sh.l, int(pure), pts, sh.coefficients, sh.exponents, sh.origin, 0)[0]
        phi[:, fb[s]:fb[s] + v.shape[0]] = v.T * O.lib().orc_common_fac(sh.l, int(pure))
    C = np.array([7.0, 5.0, -6.0])
    Vq = phi.T @ (phi * (w / np.linalg.norm(pts - C, axis=1))[:, None])
    for p, sp in enumerate(b.shells):
        for q, sq in enumerate(b.shells):
            if 4 not in (sp.l, sq.l):
                continue
            V = O.cosx_esp_pair(sp.l, sq.l, sp.exponents, sq.exponents, sp.coefficients, sq.coefficients, sp.origin,
                                sq.origin, pure, C[None, :])[0]
            blk = Vq[fb[p]:fb[p] + V.shape[0], fb[q]:fb[q] + V.shape[1]]
            assert np.abs(V - blk).max() < 5e-8, (p, q, np.abs(V - blk).max())


@pytest.mark.parametrize("pure", [False, True])
def test_analytic_overlap_matches_quadrature(oracle, pure):
    O = oracle
    Z, pos, b = _two_centre_basis(pure)
    ob = O.Basis(b.as_arrays())
    pts, w, info = O.molecular_grid(Z, pos, ob, O.default_grid_settings(max_angular_points=590, min_angular_points=590,
                                                                       reduced_first_row_element_grid=0))
    S = O.cosx_overlap(ob)
    Sn = O.cosx_numerical_overlap(ob, O.cosx_atom_grids(pts, info), pts, w)
    assert np.abs(np.diag(S) - 1.0).max() < 1e-12 or pure is False  # contracted shells are normalised (xx-type only)
    assert np.abs(S - Sn).max() < 1e-7
    assert np.abs(S - S.T).max() == 0.0


def _s_eri(exps, cen, coef):
    """(ab|cd) over normalised s primitives: closed form with F_0."""
    n = len(exps)
    from math import erf, sqrt, pi, exp

    def F0(t):
        return 1.0 if t < 1e-14 else 0.5 * sqrt(pi / t) * erf(sqrt(t))
    eri = np.zeros((n, n, n, n))
    for a in range(n):
        for b in range(n):
            p = exps[a] + exps[b]
            P = (exps[a] * cen[a] + exps[b] * cen[b]) / p
            Kab = exp(-exps[a] * exps[b] / p * np.dot(cen[a] - cen[b], cen[a] - cen[b]))
            for c in range(n):
                for d in range(n):
                    q = exps[c] + exps[d]
                    Q = (exps[c] * cen[c] + exps[d] * cen[d]) / q
                    Kcd = exp(-exps[c] * exps[d] / q * np.dot(cen[c] - cen[d], cen[c] - cen[d]))
                    t = p * q / (p + q) * np.dot(P - Q, P - Q)
                    eri[a, b, c, d] = 2 * pi ** 2.5 / (p * q * sqrt(p + q)) * Kab * Kcd * F0(t) * coef[a] * coef[b] * coef[c] * coef[d]
    return eri


def test_cosx_K_matches_exact_exchange_s_only(oracle):
    """The reference's own check (tests/dft_tests.cpp:322-369): COSX K against the exact K, here with closed-form
    two-electron integrals over an s-only basis on a bent H3 frame. Pins sign, the 0.25 (K + K^T) factor, the
    projector and the batch / pair loops of the assembled sweep."""
    O = oracle
    Z = np.array([1, 1, 1], dtype=np.int32)
    pos = np.array([[0.0, 0.0, 0.0], [1.6, 0.2, 0.0], [0.3, 1.5, 0.4]])
    b = aobasis.AOBasis(Z, pos, [], False, "s-only")
    exps_all, cen_all, coef_all = [], [], []
    for a in range(3):
        for e in (2.5, 0.8, 0.25):
            c = aobasis.incorporate_shell_norm(0, np.array([e]), np.array([1.0]))
            b.shells.append(aobasis.Shell(0, pos[a].copy(), np.array([e]), c, a))
            exps_all.append(e)
            cen_all.append(pos[a])
            coef_all.append(c[0] * 0.282094791773878143)  # common_fac(0)
    ob = O.Basis(b.as_arrays())
    eri = _s_eri(np.array(exps_all), np.array(cen_all), np.array(coef_all))
    rng = np.random.default_rng(3)
    Cocc = rng.normal(size=(ob.nbf, 2)) * 0.4
    D = Cocc @ Cocc.T
    K_exact = np.einsum("ls,mlns->mn", D, eri)
    gs = O.default_grid_settings(max_angular_points=302, min_angular_points=302, radial_precision=1e-9,
                                 reduced_first_row_element_grid=0)
    pts, w, info = O.molecular_grid(Z, pos, ob, gs)
    off = O.cosx_atom_grids(pts, info)
    P, S, Sn = O.cosx_overlap_projector(ob, off, pts, w)
    assert np.abs(S - Sn).max() < 1e-6
    K, stats = O.cosx_compute_K(ob, off, pts, w, P, D)
    assert stats[0]["pair_evals"] > 0
    assert np.abs(K - K_exact).max() < 2e-5 * np.abs(K_exact).max()
    # unrestricted: D stacks D_alpha, D_beta (each stored with the factor 1/2 of the reference)
    Da, Db = 0.5 * D, 0.5 * (rng.normal(size=(ob.nbf, 1)) * 0.3) @ np.ones((1, 1))
    Db = Db @ Db.T * 2.0
    KU, _ = O.cosx_compute_K(ob, off, pts, w, P, np.vstack([Da, Db]), unrestricted=True)
    Ka_exact = np.einsum("ls,mlns->mn", 2.0 * Da, eri)
    Kb_exact = np.einsum("ls,mlns->mn", 2.0 * Db, eri)
    n = ob.nbf
    assert np.abs(KU[:n] - Ka_exact).max() < 2e-5 * np.abs(Ka_exact).max()
    assert np.abs(KU[n:] - Kb_exact).max() < 2e-5 * max(np.abs(Kb_exact).max(), 1e-3)


def test_significant_pairs_and_screening_inputs(oracle, case):
    O = oracle
    cs = case("water", "def2-svp")
    ob = cs["obasis"]
    pq, cen, ext = O.cosx_pairs(ob)
    nsh = ob.nsh
    assert len(pq) <= nsh * (nsh + 1) // 2 and (pq[:, 0] >= pq[:, 1]).all()
    # every same-centre pair is kept (integral_engine.h:262-266); a water molecule keeps all pairs at 1e-12
    assert len(pq) == nsh * (nsh + 1) // 2
    # extent = sqrt(-ln(1e-12) / p_min) of the most diffuse primitive pair (esp.cpp:85-100)
    b = cs["basis"]
    k = 0
    p, q = pq[k]
    pmin = min(b.shells[p].exponents) + min(b.shells[q].exponents)
    assert ext[k] == pytest.approx(math.sqrt(27.631021115928547 / pmin), rel=1e-14)


# ---------------------------------------------------------------------------------------------------------
# the COSX grid producer: GridSettings::for_cosx (src/numint/grid_settings.cpp:10-18,71-103,176-181) and the ORCA branch
# of generate_atom_grid (src/numint/grid_utils.cpp:296-335,559-572,588-611,693-705)
# ---------------------------------------------------------------------------------------------------------
def _treutler_ahlrichs(n):
    """generate_treutler_alrichs_radial_grid (src/numint/grid_utils.cpp:559-572), points ascending"""
    i = np.arange(1, n + 1)
    step = np.pi / (n + 1)
    x = np.cos(i * step)
    t1 = (1 + x) ** 0.6 / np.log(2.0)
    t2 = np.log((1 - x) / 2)
    return (-t1 * t2)[::-1], (step * np.sin(i * step) * t1 * (-0.6 / (1 + x) * t2 + 1 / (1 - x)))[::-1]


def test_treutler_ahlrichs_radial_grid_golden_and_single_atom_shell_weights(oracle):
    """the reference's own known answer for the radial rule of the COSX grids (tests/dft_tests.cpp:279-289, 1e-5), and the
    oracle's single-atom COSX grid: shell k carries 4 pi (w_k r_m) (x_k r_m)^2 in total (grid_utils.cpp:603-609,680)"""
    r, w = _treutler_ahlrichs(3)
    assert np.abs(r - [0.10934791, 1.0, 3.82014324]).max() < 1e-5
    assert np.abs(w - [0.34905607, 1.60432893, 4.51614622]).max() < 1e-5
    Z, pos = np.array([8], dtype=np.int32), np.zeros((1, 3))
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    pts, wts, info = oracle.molecular_grid(Z, pos, oracle.Basis(b.as_arrays()), oracle.cosx_grid_settings(1))
    n = info["atom_nradial"][0]
    assert n == 27
    rm = 0.60 / 0.52917721092
    x, wr = _treutler_ahlrichs(n)
    rad = np.linalg.norm(pts, axis=1)
    for k in range(n):
        sel = np.abs(rad - x[k] * rm) < 1e-9 * max(1.0, x[k] * rm)
        assert sel.any()
        want = 4 * np.pi * wr[k] * rm * (x[k] * rm) ** 2
        assert abs(wts[sel].sum() - want) < 1e-8 * want, k       # the Lebedev weights sum to 1 to ~1e-9


@pytest.mark.parametrize("level,nr_h,nr_o", [(1, 22, 27), (2, 25, 30), (3, 30, 35)])
def test_cosx_grid_levels_follow_the_orca_rules(oracle, level, nr_h, nr_o):
    """radial counts nr = round(15 IntAcc - 40 + 5 row); the 5-region angular orders (H / He one Lebedev level lower);
    Treutler-Ahlrichs nodes scaled by the Bragg radius; the partitioned grid integrates a density"""
    O = oracle
    Z, pos = molecules.MOLECULES["water"]()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    ob = O.Basis(b.as_arrays())
    pts, w, info = O.molecular_grid(Z, pos, ob, O.cosx_grid_settings(level))
    assert info["atom_nradial"] == [nr_o, nr_h, nr_h]
    off = O.cosx_atom_grids(pts, info)
    regions = np.array(O.COSX_GRIDS[level][1])
    below = {14: 6, 26: 14, 50: 38, 110: 86, 194: 170}
    bragg = {1: 0.35, 8: 0.60}                                 # Angstrom (grid_utils.cpp bragg_radii: H 0.35, O 0.60)
    for a in range(3):
        P = pts[off[a]:off[a + 1]] - pos[a]
        r = np.linalg.norm(P, axis=1)
        shells, counts = np.unique(np.round(r, 10), return_counts=True)
        assert len(shells) == info["atom_nradial"][a]
        R = bragg[int(Z[a])] / 0.52917721092
        reg = np.searchsorted([0.25 * R, 0.5 * R, 1.0 * R, 2.5 * R], shells, side="right")
        want = regions[reg] if Z[a] > 2 else np.array([below[int(x)] for x in regions[reg]])
        assert np.array_equal(counts, want), (a, counts, want)
        # Treutler-Ahlrichs nodes (grid_utils.cpp:559-572) times the Bragg radius
        n = len(shells)
        i = np.arange(1, n + 1)
        x = np.cos(i * np.pi / (n + 1))
        nodes = np.sort(-(1 / np.log(2)) * (1 + x) ** 0.6 * np.log((1 - x) / 2)) * R
        assert np.abs(shells - nodes).max() < 1e-9 * nodes.max()
    # the Becke-partitioned COSX grid integrates the synthetic density's electron count to grid accuracy
    D2 = molecules.synthetic_density(b, 1)
    phi = np.zeros((len(w), b.nbf))
    col = 0
    for sh in b.shells:
        v = O.collocate_shell(sh.l, int(b.pure), pts, sh.coefficients, sh.exponents, sh.origin, 0)[0].T
        phi[:, col:col + v.shape[1]] = v * O.lib().orc_common_fac(int(sh.l), int(b.pure))
        col += v.shape[1]
    S = O.cosx_overlap(ob)
    n_exact = float((D2 * S).sum())
    n_grid = float(((phi @ D2) * phi).sum(1) @ w)
    assert abs(n_grid - n_exact) < {1: 3e-2, 2: 2e-2, 3: 5e-3}[level] * abs(n_exact)


def test_for_sgx_uniform_angular_grid(oracle):
    """GridSettings::for_sgx: PruningScheme::None keeps max_angular_points on every radial shell, also at >= 50 points
    where the NWChem scheme switches to its own level table (grid_utils.cpp:240-250,719-722)"""
    O = oracle
    Z, pos = molecules.MOLECULES["water"]()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    ob = O.Basis(b.as_arrays())
    kw = dict(max_angular_points=50, min_angular_points=50, radial_precision=1e-5, reduced_first_row_element_grid=0)
    p0, w0, i0 = O.molecular_grid(Z, pos, ob, O.default_grid_settings(**kw))
    p1, w1, i1 = O.molecular_grid(Z, pos, ob, O.default_grid_settings(pruning_scheme=1, **kw))
    assert i1["atom_nradial"] == i0["atom_nradial"]
    assert i1["atom_npts"] == [50 * n for n in i1["atom_nradial"]]
    assert i0["atom_npts"] != i1["atom_npts"]
