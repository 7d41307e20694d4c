"""Host-side logic that needs no device: method table, basis loading, extents, the leaf partition of the multi-GPU build."""
import numpy as np
import pytest

from occ_b200 import aobasis, dft, molecules


def test_general_spinorbitals_are_rejected_like_the_reference():
    """compute_vxc: "Not implemented: DFT for General spinorbitals" (include/occ/dft/dft.h:426-428) — raised before any
    device call"""
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    with pytest.raises(RuntimeError, match="General spinorbitals"):
        dft.DFT("pbe", b, dft.GridHierarchy(np.zeros((0, 3)), np.zeros(0), np.zeros(0, dtype=dft.LEAF_DTYPE)), "general")


def test_method_table_mirrors_the_reference():
    """src/dft/dft_method.cpp:14-37"""
    assert dft.get_dft_method("b3lyp") == [402]
    assert dft.get_dft_method("PBE") == [101, 130]
    assert dft.get_dft_method("blyp") == [106, 131]
    assert dft.get_dft_method("lda") == dft.get_dft_method("svwn") == [1, 30]
    assert dft.get_dft_method("svwn5") == [1, 7]
    assert dft.get_dft_method("r2scan") == [497, 498]
    assert dft.get_dft_method("gga_x_b88 gga_c_lyp") == [106, 131]
    assert dft.get_dft_method("bp86") == [106, 132] and dft.get_dft_method("b97") == [407]
    assert dft.get_dft_method("b3pw91") == [401] and dft.get_dft_method("wb97x") == [464]
    assert dft.get_dft_method("ub3lyp") == [402] and dft.get_dft_method("upbe") == [101, 130]   # leading 'u' stripped (:175-178)
    # entries of share/methods/dft_methods.json beyond the compiled-in table
    assert dft.get_dft_method("wb97x-v") == [466] and dft.get_dft_method("b3p86") == [403]
    assert dft.get_dft_method("b86bpbe") == [105, 130] and dft.get_dft_method("b86bpbeh") == [105, 130]
    for missing in ("m062x", "wb97m", "wb97m-v"):            # libxc components this library does not implement
        with pytest.raises(RuntimeError):
            dft.get_dft_method(missing)
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


# ------------------------------------------------------------------------------------------------------------------
# The product's DEVICE functionals (occ_b200/csrc/xc_functionals.cuh, __host__ __device__) compiled for the host by a
# test harness (tests/cpp/functional_table.cu): the CPU suite can then check them against the oracle and the
# reference's known answers before any GPU time is spent.
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FT_EXE = os.path.join(ROOT, "tests", "cpp", "functional_table")


def _functional_table_exe():
    src = os.path.join(ROOT, "tests", "cpp", "functional_table.cu")
    inc = os.path.join(ROOT, "occ_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("xc_functionals.cuh", "xc_dual.cuh")]
    if not os.path.exists(FT_EXE) or os.path.getmtime(FT_EXE) < max(os.path.getmtime(d) for d in deps):
        nvcc = "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            pytest.skip("nvcc not available and tests/cpp/functional_table not prebuilt")
        subprocess.check_call([nvcc, "-std=c++17", "-O1", "-w", "-ccbin", "/usr/bin/g++", "-I" + inc, "-o", FT_EXE, src])
    return FT_EXE


def _device_functional_on_host(fid, polarized, fam, rho, sigma, tau):
    n = len(rho)
    s, g = (2, 3) if polarized else (1, 1)
    buf = b"".join(np.ascontiguousarray(a, dtype=np.float64).tobytes() for a in (rho, sigma, tau))
    out = subprocess.run([_functional_table_exe(), str(fid), str(int(polarized)), str(fam), str(n)], input=buf,
                         capture_output=True, check=True).stdout
    r = np.frombuffer(out, dtype=np.float64).reshape(n, 1 + s + g + s)
    return {"exc": r[:, 0], "vrho": r[:, 1:1 + s], "vsigma": r[:, 1 + s:1 + s + g], "vtau": r[:, 1 + s + g:]}


def test_device_functionals_on_host_match_oracle_and_reference_kat(oracle):
    gk = np.load(os.path.join(ROOT, "tests", "golden", "xc_golden.npz"))
    rows = gk["tpss_kat_rows"]
    rho, sigma, tau = rows[:, 0], (rows[:, 1:4] ** 2).sum(1), rows[:, 5]
    TPSS_KAT = {"exc_x": gk["tpss_kat_x_exc"], "exc_c": gk["tpss_kat_c"][0], "vrho_c": gk["tpss_kat_c"][1],
                "vsigma_c": gk["tpss_kat_c"][2], "vtau_c": gk["tpss_kat_c"][3]}
    x = _device_functional_on_host(202, False, 2, rho, sigma, tau)
    c = _device_functional_on_host(231, False, 2, rho, sigma, tau)
    assert np.abs(x["exc"] - TPSS_KAT["exc_x"]).max() < 1e-6          # tests/dft_tests.cpp:138-191
    assert np.abs(c["exc"] - TPSS_KAT["exc_c"]).max() < 1e-6
    assert np.abs(c["vrho"][:, 0] - TPSS_KAT["vrho_c"]).max() < 1e-6
    assert np.abs(c["vsigma"][:, 0] - TPSS_KAT["vsigma_c"]).max() < 1e-6
    assert np.abs(c["vtau"][:, 0] - TPSS_KAT["vtau_c"]).max() < 1e-6
    g = np.load(os.path.join(ROOT, "tests", "golden", "xc_golden.npz"))
    rho, sig, tau = g["fx_rho"], g["fx_sigma"], g["fx_tau"]
    for name, fid in dft.FUNCTIONAL_IDS.items():
        fam = oracle.lib().orc_xc_family(fid)
        u = _device_functional_on_host(fid, False, fam, rho.sum(1), sig[:, 0] + 2 * sig[:, 1] + sig[:, 2], tau.sum(1))
        p = _device_functional_on_host(fid, True, fam, rho, sig, tau)
        for got, ref in ((u, g[f"fx_{name}_u"]), (p, g[f"fx_{name}_p"])):
            s = got["vrho"].shape[1]
            nsig = got["vsigma"].shape[1]
            want = {"exc": ref[:, 0], "vrho": ref[:, 1:1 + s], "vsigma": ref[:, 1 + s:1 + s + nsig], "vtau": ref[:, 1 + s + nsig:]}
            keys = ["exc", "vrho"] + (["vsigma"] if fam >= 1 else []) + (["vtau"] if fam >= 2 else [])
            for k in keys:
                assert np.allclose(got[k], want[k], rtol=1e-10, atol=1e-13), (name, k)


def test_device_functionals_on_host_hydrogen_atom_constraints():
    """the same published hydrogen-atom values / exact constraints as tests/test_oracle_pins.py, evaluated with the
    PRODUCT's device functionals (polarized branch) compiled for the host"""
    from test_oracle_pins import HYDROGEN, hydrogen_atom_quadrature
    rho, sigma, tau, qw = hydrogen_atom_quadrature()
    keep = rho[:, 0] > 1e-14                      # below the functionals' density threshold nothing is evaluated
    for func, (want, tol) in HYDROGEN.items():
        fid = dft.FUNCTIONAL_IDS[func]
        fam = 0 if func.startswith("lda") else (1 if func.startswith("gga") else 2)
        o = _device_functional_on_host(fid, True, fam, rho[keep], sigma[keep], tau[keep])
        assert abs(float((o["exc"] * qw[keep]).sum()) - want) < max(tol, 1e-9), func


def test_device_functionals_on_host_lithium_atom():
    """partial spin polarization (0 < zeta < 1): the product's device functionals on the UHF lithium density against
    the published energies of tests/test_oracle_pins.py (PBE Table I, Becke 88, LYP)"""
    from test_oracle_pins import ATOMS, atom_quadrature
    res, rho, sigma, tau, vol = atom_quadrature("Li")
    keep = rho.sum(1) > 1e-14
    for func, want in ATOMS["Li"][3].items():
        fam = 0 if func.startswith("lda") else (1 if func.startswith("gga") else 2)
        o = _device_functional_on_host(dft.FUNCTIONAL_IDS[func], True, fam, rho[keep], sigma[keep], tau[keep])
        got = float((o["exc"] * rho.sum(1)[keep] * vol[keep]).sum())
        assert abs(got - want) < 1.5e-4, (func, got, want)


def test_dft_methods_json_overrides_the_builtin_table(tmp_path, monkeypatch):
    """get_method_definitions (src/dft/dft_method.cpp:96-132): share/methods/dft_methods.json under OCC_DATA_PATH
    replaces the builtin table; components this library does not implement fail with the reference's message"""
    import json
    from occ_b200 import dft
    (tmp_path / "methods").mkdir()
    (tmp_path / "methods" / "dft_methods.json").write_text(json.dumps({"methods": {
        "mypbe": {"components": [{"id": "gga_x_b88"}, {"id": "gga_c_pbe", "factor": 0.5}]},
        "m062x": {"components": [{"id": "hyb_mgga_x_m06_2x"}, {"id": "mgga_c_m06_2x"}]}}}))
    assert dft.get_dft_method("pbe") == [101, 130]                      # builtin, no OCC_DATA_PATH file
    monkeypatch.setenv("OCC_DATA_PATH", str(tmp_path))
    assert dft.get_dft_method("mypbe") == [106, 130]
    assert dft.get_dft_method("gga_x_pbe") == [101]                     # plain libxc names still resolve
    with pytest.raises(RuntimeError):
        dft.get_dft_method("pbe")                                       # the file REPLACES the builtin table
    with pytest.raises(RuntimeError, match="hyb_mgga_x_m06_2x not found"):
        dft.get_dft_method("m062x")
    # the reference's own share/methods file parses (when the reference tree is present in this container)
    ref = "/root/reference/share"
    if os.path.exists(os.path.join(ref, "methods", "dft_methods.json")):
        monkeypatch.setenv("OCC_DATA_PATH", ref)
        assert dft.get_dft_method("b3lyp") == [402] and dft.get_dft_method("wb97x") == [464]
