"""XDM dispersion consumer over the GPU density: occ_b200.xdm mirrors src/xdm/xdm.cpp and
src/xdm/becke_hole.cpp.  The pair sum is pinned by the reference's own golden (tests/xdm_tests.cpp:40-92: water, POSTG
polarizabilities, energy and forces to 1e-8); the Becke-Roussel inversion by its defining equation; the moment sums over
the GPU deriv-2 density by a dense numpy evaluation of evaluate_density<2> (src/gto/density.cpp:72-101) over the
oracle's collocation."""
import numpy as np
import pytest

from occ_b200 import aobasis, dft, molecules, xdm


def test_xdm_pair_energy_and_forces_match_the_reference_golden():
    """tests/xdm_tests.cpp:40-92"""
    frepol = {1: 0.6668, 8: 0.8020}
    Z = [8, 1, 1]
    pos = np.array([[-1.3269576, -0.1059386, 0.0187882], [-1.9316642, 1.6001735, -0.0217105],
                    [0.4866441, 0.0795981, 0.0098625]])
    moments = np.array([[4.4746785658e+00, 1.0884283711e+00, 1.0929537397e+00],
                        [2.8898799370e+01, 6.6743870561e+00, 6.6878435445e+00],
                        [1.8355332889e+02, 6.7438103019e+01, 6.7479436992e+01]])
    volume = np.array([1.8941883873e+01, 3.8999421841e+00, 3.9031603917e+00])
    vfree = np.array([2.3504820578e+01, 8.7017290471e+00, 8.7017290471e+00])
    pol = np.array([volume[i] * frepol[Z[i]] / 0.52917720859 ** 3 / vfree[i] for i in range(3)])
    e, f = xdm.xdm_dispersion_energy(pos, pol, moments, 0.7, 1.4)
    assert abs(e - (-3.976524483082e-04)) < 1e-8
    want = np.array([[2.248250121483e-07, 9.975579765949e-07, -1.222382988743e-06],
                     [3.350812651483e-07, -8.612850877111e-07, 5.262038225628e-07],
                     [-8.780610885086e-09, 1.877910301194e-08, -9.998492126856e-09]])      # Mat3N: rows x, y, z
    assert np.allclose(f.T, want, rtol=1e-8, atol=1e-8)
    assert np.abs(f.sum(0)).max() < 1e-18


def test_becke_hole_analytic_inverts_its_defining_equation():
    """x exp(-2x/3) / (x - 2) = (2/3) pi^(2/3) rho^(5/3) / Q (becke_hole.cpp:8-13,70-78); the rational fit is good to a few 1e-5 relative"""
    rng = np.random.default_rng(0)
    rho = 10.0 ** rng.uniform(-6, 1, 400)
    Q = rng.normal(size=400) * rho ** (5.0 / 3.0) * 3.0
    b = xdm.becke_hole_br89_analytic(rho, Q, 1.0)
    # recover x from b: alpha = (8 pi rho e^x)^(1/3), b = x / alpha  ->  solve for x by Newton on the defining equation
    rhs = (2.0 / 3.0) * (np.pi * rho) ** (2.0 / 3.0) * rho / Q
    x = np.where(rhs > 0, 3.0, 1.0)
    for _ in range(200):
        ex = np.exp(-2.0 / 3.0 * x)
        f = x * ex / (x - 2.0) - rhs
        df = 2.0 / 3.0 * (2.0 * x - x * x - 3.0) / (x - 2.0) ** 2 * ex
        x = np.clip(x - f / df, np.where(rhs > 0, 2.0 + 1e-12, 1e-14), np.where(rhs > 0, 1e3, 2.0 - 1e-12))
    want = x / (8.0 * np.pi * rho * np.exp(x)) ** (1.0 / 3.0)
    assert np.all(np.isfinite(b)) and np.all(b > 0)
    assert np.abs(b / want - 1.0).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("nspin", [1, 2])
def test_xdm_moments_over_the_gpu_density(oracle, nspin):
    Z, pos = molecules.MOLECULES["water"]()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    ob = oracle.Basis(b.as_arrays())
    pts, w, info = oracle.molecular_grid(Z, pos, ob)
    off = oracle.cosx_atom_grids(pts, info)
    grids = [(pts[off[a]:off[a + 1]], w[off[a]:off[a + 1]]) for a in range(len(Z))]
    zeta = {1: 1.0, 8: 2.2}
    pro = lambda i, r: (zeta[int(Z[i])] ** 3 / np.pi) * np.exp(-2.0 * zeta[int(Z[i])] * r) * int(Z[i])   # noqa: E731
    fpol = np.array([{1: 4.5, 8: 5.4}[int(z)] for z in Z])
    D = 0.5 * molecules.synthetic_density(b, nspin)                      # mo.D
    kind = dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED
    x = xdm.XDM(b, grids, pro, fpol)
    e = x.energy(D, kind)
    # dense numpy evaluate_density<2> over the oracle's collocation
    P, W = x.pts, x.w
    phi = [np.zeros((len(P), b.nbf)) for _ in range(10)]
    col = 0
    for sh in b.shells:
        v = oracle.collocate_shell_deriv2(sh.l, int(b.pure), P, sh.coefficients, sh.exponents, sh.origin)
        fac = oracle.lib().orc_common_fac(int(sh.l), int(b.pure))
        for k in range(10):
            phi[k][:, col:col + v.shape[1]] = v[k].T * fac
        col += v.shape[1]
    rho = np.zeros((len(P), 6))
    for s in range(nspin):
        Ds = D[s * b.nbf:(s + 1) * b.nbf]
        Dr = np.triu(Ds) + np.triu(Ds, 1).T if nspin == 1 else Ds
        X = phi[0] @ Dr
        rho[:, 0] += (phi[0] * X).sum(1)
        for d in range(3):
            rho[:, 1 + d] += 2 * (phi[1 + d] * X).sum(1)
        lap = 2 * ((phi[4] + phi[7] + phi[9]) * X).sum(1)
        tau = sum((phi[1 + d] * (phi[1 + d] @ Ds)).sum(1) for d in range(3))
        rho[:, 4] += lap + 2 * tau
        rho[:, 5] += 0.5 * tau
    r = np.linalg.norm(P[:, None, :] - np.asarray(pos).reshape(-1, 3)[None], axis=2)
    rp = np.stack([pro(i, r[:, i]) for i in range(len(Z))], 1)
    hw = rp / rp.sum(1)[:, None]
    sig = (rho[:, 1:4] ** 2).sum(1)
    q = (rho[:, 4] - 2 * (2 * rho[:, 5] - 0.25 * sig / np.maximum(rho[:, 0], 1e-30))) / 6.0
    bh = xdm.becke_hole_br89_analytic(rho[:, 0], q, 1.0)
    rsb = np.maximum(r - bh[:, None], 0.0)
    wt = (2 * rho[:, 0] * W)[:, None] * hw
    want_m = np.stack([(wt * (r ** k - rsb ** k) ** 2).sum(0) for k in (1, 2, 3)])
    assert np.abs(x.moments / want_m - 1.0).max() < 1e-8
    assert np.abs(x.volume - (wt * r ** 3).sum(0)).max() < 1e-8
    assert abs(x.num_electrons - (2 * rho[:, 0] * W).sum()) < 1e-9
    pol = x.volume * fpol / x.volume_free
    ew, fw = xdm.xdm_dispersion_energy(pos, pol, want_m)
    assert abs(e - ew) < 1e-10 and e < -1e-5
    assert np.abs(x.forces(D, kind) - fw).max() < 1e-10
