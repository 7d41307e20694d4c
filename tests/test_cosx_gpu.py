"""GPU parity of the COSX build (occ_cosx_*, include/occ_cosx_b200.h) against the CPU oracle (oracle/src/cosx.cpp) on
the same basis, atom grids, projector and density. Tolerance: 1e-9 max-abs on K (north-star Vxc tolerance; the two
sides differ by summation order, the Boys downward recursion and libm, all at the 1e-14 level)."""
import numpy as np
import pytest

from occ_b200 import aobasis, molecules

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _setup(oracle, mol, basis_name, pure, angular=38, radial_precision=1e-5, seed=7, nspin=1):
    from occ_b200 import cosx
    O = oracle
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, basis_name, pure=pure)
    ob = O.Basis(b.as_arrays())
    gs = O.default_grid_settings(max_angular_points=angular, min_angular_points=angular,
                                 radial_precision=radial_precision, reduced_first_row_element_grid=0)
    pts, w, info = O.molecular_grid(Z, pos, ob, gs)
    off = O.cosx_atom_grids(pts, info)
    P, S, Sn = O.cosx_overlap_projector(ob, off, pts, w)
    D = molecules.synthetic_density(b, nspin=nspin, seed=seed) * 0.5   # mo.D carries the factor 1/2
    sgx = cosx.SemiNumericalExchange(b, atom_grids=(off, pts, w), projector=P, collect_stats=True)
    return O, b, ob, off, pts, w, P, S, Sn, D, sgx


@pytest.mark.parametrize("mol,basis_name,pure", [("water", "3-21g", False), ("water", "def2-svp", False),
                                                 ("water", "def2-svp", True), ("water", "def2-tzvp", False),
                                                 ("water", "def2-tzvp", True), ("socl2_na", "def2-svp", True)])
def test_cosx_K_matches_oracle(oracle, mol, basis_name, pure):
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, mol, basis_name, pure)
    K = sgx.compute_K(D)
    Kref, stats = O.cosx_compute_K(ob, off, pts, w, P, D)
    assert np.abs(Kref).max() > 1e-3
    assert np.abs(K - Kref).max() < TOL, np.abs(K - Kref).max()
    info = sgx.info()
    assert info["pair_evals"] == stats[0]["pair_evals"]
    assert info["skipped_batches"] == stats[0]["nskipped_batches"]
    assert info["dense_flops"] == pytest.approx(stats[0]["dense_flops"], rel=1e-12)
    assert info["npairs"] == stats[0]["npairs"]
    assert info["launches"] > 0
    sgx.close()


def test_cosx_two_centre_f_pairs(oracle):
    """benzene def2-TZVP: f shells on six carbons, so every (la, lb) class up to (3, 3) has two-centre pairs"""
    for pure in (True, False):
        O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "benzene", "def2-tzvp", pure, angular=14,
                                                         radial_precision=1e-3)
        K = sgx.compute_K(D)
        Kref, stats = O.cosx_compute_K(ob, off, pts, w, P, D)
        assert np.abs(K - Kref).max() < TOL, np.abs(K - Kref).max()
        assert sgx.info()["pair_evals"] == stats[0]["pair_evals"]
        sgx.close()


def test_cosx_raw_K_and_overlaps_match_oracle(oracle):
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", False)
    assert np.abs(sgx.analytic_overlap() - S).max() < 1e-14
    assert np.abs(sgx.compute_overlap_matrix() - Sn).max() < 1e-12
    D2q = P @ (2.0 * D)
    Kraw = sgx.compute_K_for_density(D2q)
    Kref, _ = O.cosx_K_for_density(ob, off, pts, w, D2q)
    assert np.abs(Kraw - Kref).max() < TOL
    sgx.close()


def test_cosx_projector_built_by_the_library(oracle):
    from occ_b200 import cosx
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", True)
    sgx.close()
    sgx2 = cosx.SemiNumericalExchange(b, atom_grids=(off, pts, w))          # projector = solve(Sn, S) from the GPU overlaps
    assert np.abs(sgx2.overlap_projector - P).max() < 1e-9
    K = sgx2.compute_K(D)
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, P, D)
    assert np.abs(K - Kref).max() < 1e-8
    sgx2.close()


def test_cosx_unrestricted_matches_oracle(oracle):
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", False, nspin=2)
    assert D.shape == (2 * b.nbf, b.nbf)
    K = sgx.compute_K(D, unrestricted=True)
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, P, D, unrestricted=True)
    assert np.abs(K - Kref).max() < TOL
    sgx.close()


@pytest.mark.parametrize("mol,basis_name,pure", [("water", "def2-svp", False), ("water", "def2-tzvp", True)])
def test_cosx_mo_form_restricted(oracle, mol, basis_name, pure):
    """occ_cosx_compute_K_mo: F = (Phi.(P C)).C^T instead of Phi.(P 2D) — same K as the D form and as the oracle"""
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, mol, basis_name, pure)
    C = molecules.synthetic_orbitals(b, 1, seed=7)[0]
    assert np.abs(C @ C.T - 2.0 * D).max() < 1e-14
    K = sgx.compute_K(D)
    Kmo = sgx.compute_K_mo(D, C)
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, P, D)
    assert np.abs(Kmo - K).max() < 1e-12, np.abs(Kmo - K).max()
    assert np.abs(Kmo - Kref).max() < TOL
    assert sgx.info()["gemm_flops"] < 0.5 * 4.0 * len(w) * b.nbf ** 2 + 4.0 * len(w) * b.nbf * C.shape[1] + 1
    # long-range and fused range-separated through the same factorisation
    omega, alpha, beta = 0.3, 1.0, -0.842294
    Kf = sgx.compute_K_range_separated_mo(D, C, alpha, beta, omega)
    assert np.abs(Kf - sgx.compute_K_range_separated(D, alpha, beta, omega)).max() < 1e-12
    sgx.set_range_separated_omega(omega)
    assert np.abs(sgx.compute_K_mo(D, C) - sgx.compute_K(D)).max() < 1e-12
    sgx.close()


def test_cosx_mo_form_unrestricted_and_empty_spin(oracle):
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", False, nspin=2)
    Ca, Cb = molecules.synthetic_orbitals(b, 2, seed=7)
    n = b.nbf
    assert np.abs(Ca @ Ca.T - 2.0 * D[:n]).max() < 1e-14 and np.abs(Cb @ Cb.T - 2.0 * D[n:]).max() < 1e-14
    K = sgx.compute_K(D, unrestricted=True)
    Kmo = sgx.compute_K_mo(D, (Ca, Cb), unrestricted=True)
    assert np.abs(Kmo - K).max() < 1e-12
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, P, D, unrestricted=True)
    assert np.abs(Kmo - Kref).max() < TOL
    # different occupied counts per spin, and a spin without electrons (hydrogen-atom-like): K_beta = 0
    D1 = D.copy()
    D1[:n] = 0.5 * Ca[:, :3] @ Ca[:, :3].T
    D1[n:] = 0.0
    K1 = sgx.compute_K_mo(D1, (Ca[:, :3], np.zeros((n, 0))), unrestricted=True)
    assert np.abs(K1 - sgx.compute_K(D1, unrestricted=True)).max() < 1e-12
    assert np.abs(K1[n:]).max() == 0.0 and np.abs(K1[:n]).max() > 1e-3
    sgx.close()


def test_cosx_general_spinorbitals_match_oracle(oracle):
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", True)
    n = b.nbf
    rng = np.random.default_rng(5)
    C = rng.normal(size=(2 * n, 5)) * 0.3
    Dg = 0.5 * (C @ C.T)                     # symmetric general density with non-zero ab / ba blocks
    K = sgx.compute_K(Dg, general=True)
    Kref = O.cosx_compute_K_general(ob, off, pts, w, P, Dg)
    assert np.abs(Kref[n:, :n]).max() > 1e-4 and np.abs(Kref[:n, n:] - Kref[n:, :n].T).max() < 1e-12
    assert np.abs(K - Kref).max() < TOL
    sgx.close()


def test_cosx_long_range_and_fused_range_separated(oracle):
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", True)
    omega, alpha, beta = 0.3, 1.0, -0.842294      # tests/dft_tests.cpp:434-436 (wB97X-style)
    Kf = sgx.compute_K(D)
    sgx.set_range_separated_omega(omega)
    Klr = sgx.compute_K(D)
    sgx.set_range_separated_omega(0.0)
    Klr_ref, _ = O.cosx_compute_K(ob, off, pts, w, P, D, omega=omega)
    assert np.abs(Klr - Klr_ref).max() < TOL
    assert np.abs(Klr).max() < np.abs(Kf).max()       # erf attenuates (tests/dft_tests.cpp:392-394)
    Kfused = sgx.compute_K_range_separated(D, alpha, beta, omega)
    assert np.abs(Kfused - ((alpha + beta) * Kf - beta * Klr)).max() < 1e-10   # tests/dft_tests.cpp:452-456
    # unrestricted range-separated build (seminumerical_exchange.cpp:881-893): two fused sweeps, one per spin block
    Du = np.asfortranarray(np.vstack([0.6 * D, 0.4 * D.T @ np.eye(b.nbf)]))
    Ku_full = sgx.compute_K(Du, unrestricted=True)
    sgx.set_range_separated_omega(omega)
    Ku_lr = sgx.compute_K(Du, unrestricted=True)
    sgx.set_range_separated_omega(0.0)
    Ku_fused = sgx.compute_K_range_separated(Du, alpha, beta, omega, unrestricted=True)
    assert Ku_fused.shape == (2 * b.nbf, b.nbf)
    assert np.abs(Ku_fused - ((alpha + beta) * Ku_full - beta * Ku_lr)).max() < 1e-10
    sgx.close()


def test_cosx_screens_follow_the_reference(oracle):
    """A spread-out system so that the geometric pair screen and the density screen both fire; an empty density gives
    K = 0 with every pair skipped."""
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "h2o_32", "3-21g", False, angular=14, radial_precision=1e-3)
    nb = b.nbf
    Dloc = D.copy()
    Dloc[np.abs(Dloc) < 2e-3] = 0.0          # sparse density: many pair blocks fall below 1e-6
    K = sgx.compute_K(Dloc)
    Kref, stats = O.cosx_compute_K(ob, off, pts, w, P, Dloc)
    info = sgx.info()
    assert info["pair_evals"] == stats[0]["pair_evals"]
    assert info["pair_evals"] < info["npairs"] * info["nbatches"]
    assert np.abs(K - Kref).max() < TOL
    K0 = sgx.compute_K(np.zeros((nb, nb)))
    assert np.abs(K0).max() == 0.0 and sgx.info()["pair_evals"] == 0
    sgx.close()


def test_cosx_sharded_partials_sum_to_K(oracle):
    from occ_b200 import cosx
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-svp", False)
    K = sgx.compute_K(D)
    sgx.close()
    parts = []
    for r in range(3):
        s = cosx.SemiNumericalExchange(b, atom_grids=(off, pts, w), projector=P, rank=r, world_size=3)
        assert s.num_batches() == len(cosx.batch_shard(off, r, 3))      # the rule the CPU gloo test shards by
        parts.append(s.compute_K(D))
        s.close()
    assert np.abs(sum(parts) - K).max() < 1e-12


@pytest.mark.parametrize("mol,pure", [("water", True), ("water", False), ("benzene", True)])
def test_cosx_g_shells(oracle, mol, pure):
    """def2-QZVP: g functions on the heavy atoms, i.e. the (4, lb) classes up to (4, 4) (L = 8); benzene adds two-centre
    g-g pairs"""
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, mol, "def2-qzvp", pure, angular=14, radial_precision=1e-3)
    assert max(sh.l for sh in b.shells) == 4
    K = sgx.compute_K(D)
    Kref, stats = O.cosx_compute_K(ob, off, pts, w, P, D)
    assert np.abs(Kref).max() > 1e-3
    assert np.abs(K - Kref).max() < TOL, np.abs(K - Kref).max()
    info = sgx.info()
    assert info["pair_evals"] == stats[0]["pair_evals"]
    assert info["dense_flops"] == pytest.approx(stats[0]["dense_flops"], rel=1e-12)
    assert np.abs(sgx.analytic_overlap() - S).max() < 1e-12
    sgx.close()


def test_cosx_rejects_h_shells_loudly():
    from occ_b200 import cosx
    Z, pos = molecules.MOLECULES["water"]()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    sh = b.shells[0]
    exps = np.array([1.3])
    b.shells.append(aobasis.Shell(5, sh.origin.copy(), exps, aobasis.incorporate_shell_norm(5, exps, np.array([1.0])), 0))
    offs = np.array([0, 1, 2, 3], dtype=np.uint64)
    with pytest.raises(RuntimeError, match="l = 4"):
        cosx.SemiNumericalExchange(b, atom_grids=(offs, np.zeros((3, 3)), np.ones(3)), projector=None)


def test_cosx_ragged_atom_grids(oracle):
    """atom grids of 1, 129 and 300 points: a one-point batch, a batch boundary at 128 + 1, and the numerical overlap's
    dropped last point per atom (seminumerical_exchange.cpp:349-352) on grids where that point matters"""
    from occ_b200 import cosx
    O = oracle
    Z, pos = molecules.MOLECULES["water"]()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp", pure=True)
    ob = O.Basis(b.as_arrays())
    rng = np.random.default_rng(11)
    counts = [1, 129, 300]
    pts = np.concatenate([pos[a] + rng.normal(scale=1.2, size=(n, 3)) for a, n in enumerate(counts)])
    w = rng.uniform(0.01, 0.2, size=len(pts))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    D = molecules.synthetic_density(b, 1, seed=3) * 0.5
    sgx = cosx.SemiNumericalExchange(b, atom_grids=(off, pts, w), projector=None)
    assert sgx.num_batches() == 1 + 2 + 3
    assert np.abs(sgx.compute_overlap_matrix() - O.cosx_numerical_overlap(ob, off, pts, w)).max() < 1e-12
    K = sgx.compute_K(D)
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, np.eye(b.nbf), D)
    assert np.abs(K - Kref).max() < TOL
    sgx.close()


def test_cosx_on_the_for_cosx_grid(oracle):
    """the engine as HartreeFock::set_cosx_exchange builds it (src/qm/hf.cpp:62-68): GridSettings::for_cosx(Grid1) from the
    product's own grid producer, projector solved by the mirror, K against the oracle on the oracle's copy of that grid"""
    from occ_b200 import cosx
    O = oracle
    Z, pos = molecules.MOLECULES["water"]()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp", pure=True)
    ob = O.Basis(b.as_arrays())
    sgx = cosx.SemiNumericalExchange(b, settings=cosx.cosx_grid_settings(1))
    pts, w, info = O.molecular_grid(Z, pos, ob, O.cosx_grid_settings(1))
    off = O.cosx_atom_grids(pts, info)
    assert sgx.num_grid_points() == len(w) and np.array_equal(sgx.atom_offsets, off)
    P, S, Sn = O.cosx_overlap_projector(ob, off, pts, w)
    D = molecules.synthetic_density(b, 1, seed=7) * 0.5
    K = sgx.compute_K(D)
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, P, D)
    assert np.abs(K - Kref).max() < 1e-8 * max(1.0, np.abs(Kref).max())   # the projector is solved on either side
    sgx.close()


@pytest.mark.parametrize("devices", [(0, 0), (0, 0, 0), (0, 1), (0, 1, 2, 3)])
def test_cosx_single_process_multi_gpu(oracle, devices):
    """occ_cosx_multi_*: shards of grid batches driven from one process; K (D form, factorised form, unrestricted,
    fused range-separated) equals the one-handle build to 1e-12 and the oracle to 1e-9; the projector is solved from the
    summed numerical overlap of the shards"""
    import torch
    from occ_b200 import cosx
    if max(devices) >= torch.cuda.device_count():
        pytest.skip(f"needs {max(devices) + 1} GPUs")
    O, b, ob, off, pts, w, P, S, Sn, D, sgx = _setup(oracle, "water", "def2-tzvp", True)
    K1 = sgx.compute_K(D)
    m = cosx.MultiGPUSemiNumericalExchange(b, atom_grids=(off, pts, w), devices=devices)
    assert sum(m.info(r)["nbatches"] for r in range(len(devices))) == sgx.info()["nbatches"]
    assert np.abs(m.compute_overlap_matrix() - Sn).max() < 1e-12
    assert np.abs(m.overlap_projector - P).max() < 1e-8
    m.set_projector(P)
    Kref, _ = O.cosx_compute_K(ob, off, pts, w, P, D)
    Cocc = molecules.synthetic_orbitals(b, 1, seed=7)[0]
    for K in (m.compute_K(D), m.compute_K(D, Cocc)):
        assert np.abs(K - K1).max() < 1e-12 and np.abs(K - Kref).max() < TOL
    Du = np.asfortranarray(np.vstack([0.6 * D, 0.4 * D]))
    assert np.abs(m.compute_K(Du, unrestricted=True) - sgx.compute_K(Du, unrestricted=True)).max() < 1e-12
    Kf = m.compute_K_range_separated(D, 1.0, -0.842294, 0.3)
    assert np.abs(Kf - sgx.compute_K_range_separated(D, 1.0, -0.842294, 0.3)).max() < 1e-12
    m.close()
    sgx.close()
