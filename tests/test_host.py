"""Host-side logic that needs no device: method table, basis loading, extents, the leaf partition of the multi-GPU build."""
import numpy as np
import pytest

from occ_b200 import aobasis, dft, molecules


def test_method_table_mirrors_the_reference():
    """src/dft/dft_method.cpp:14-37"""
    assert dft.get_dft_method("b3lyp") == [402]
    assert dft.get_dft_method("PBE") == [101, 130]
    assert dft.get_dft_method("blyp") == [106, 131]
    assert dft.get_dft_method("lda") == dft.get_dft_method("svwn") == [1, 30]
    assert dft.get_dft_method("svwn5") == [1, 7]
    assert dft.get_dft_method("r2scan") == [497, 498]
    assert dft.get_dft_method("gga_x_b88 gga_c_lyp") == [106, 131]
    with pytest.raises(RuntimeError):
        dft.get_dft_method("not_a_functional")


def test_decay_cutoff_matches_oracle(oracle):
    """gto::evaluate_decay_cutoff (src/gto/gto.cpp:18-67): host code of the library vs the oracle — the same rung of
    the 0.1-Bohr ladder for every shell (i*0.1+0.5 may differ in the last ulp: the oracle is compiled with FMA)"""
    for mol, bas, pure in (("water", "def2-svp", False), ("benzene", "def2-tzvp", False), ("benzene", "def2-tzvp", True),
                           ("water", "def2-qzvp", True)):
        Z, pos = molecules.MOLECULES[mol]()
        b = aobasis.AOBasis.load(Z, pos, bas, pure=pure)
        ours = dft.decay_cutoff(b, 1e-10)
        ref = oracle.decay_cutoff(oracle.Basis(b.as_arrays()), 1e-10)
        assert np.array_equal(np.rint((ours - 0.5) / 0.1), np.rint((ref - 0.5) / 0.1))
        assert np.abs(ours - ref).max() < 1e-14
        assert ours.min() >= 0.5 and ours.max() <= 30.0


def test_leaf_partition_is_contiguous_balanced_and_complete(oracle, case):
    c = case("benzene", "def2-tzvp")
    for world in (1, 2, 4, 8):
        rank, nb = dft.partition_leaves(c["basis"], c["leaves"], world)
        assert (rank >= 0).all() and rank.max() == world - 1
        assert (np.diff(rank) >= 0).all()                      # Morton-contiguous runs
        cost = c["leaves"]["count"].astype(float) * nb.astype(float) ** 2
        per = np.array([cost[rank == r].sum() for r in range(world)])
        assert per.sum() == cost.sum()
        assert per.max() <= cost.sum() / world + cost.max() + 1e-9
    # nb agrees with the oracle's shell lists
    ext = oracle.decay_cutoff(c["obasis"])
    import ctypes
    buf = np.zeros(c["basis"].nsh, dtype=np.int32)
    sizes = np.array([aobasis.shell_size(s.l, False) for s in c["basis"].shells])
    for li in range(0, len(c["leaves"]), 97):
        lf = oracle.LeafT(int(c["leaves"][li]["offset"]), int(c["leaves"][li]["count"]),
                          (ctypes.c_double * 3)(*c["leaves"][li]["center"]), float(c["leaves"][li]["radius"]))
        n = oracle.lib().orc_leaf_shell_list(c["obasis"].ref, ext.ctypes.data_as(ctypes.c_void_p), ctypes.byref(lf),
                                             buf.ctypes.data_as(ctypes.c_void_p))
        assert sizes[buf[:n]].sum() == nb[li]


def test_synthetic_density_is_symmetric_psd():
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    D = molecules.synthetic_density(b, 1)
    assert np.allclose(D, D.T) and np.linalg.eigvalsh(D).min() > -1e-12
    D2 = molecules.synthetic_density(b, 2)
    assert D2.shape == (2 * b.nbf, b.nbf)
