"""The C++ host side (adapter/occ_dft_b200.h) over the C ABI: compiled with g++ against libocc_xc_b200.so.
CPU: it builds, links and turns a library failure (no device) into std::runtime_error. GPU: K/Exc parity with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_driver")


def _compile():
    src = os.path.join(ROOT, "tests", "cpp", "adapter_driver.cpp")
    lib = os.path.join(ROOT, "occ_b200")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(ROOT, "adapter", "occ_dft_b200.h"))):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-o", EXE, src, "-L" + lib, "-locc_xc_b200",
                               "-Wl,-rpath," + lib])
    return EXE


def _write_case(path, basis, nspin, func_ids, spts, sw, leaves, D, Cocc=None, perm=None, atom_offsets=None):
    a = basis.as_arrays()

    def vec(f, arr, fmt):
        arr = np.ascontiguousarray(arr, dtype=fmt)
        f.write(struct.pack("q", arr.size))
        f.write(arr.tobytes())

    with open(path, "wb") as f:
        vec(f, [a["pure"], nspin] + list(func_ids), np.int32)
        vec(f, a["l"], np.int32)
        vec(f, a["nprim"], np.int32)
        vec(f, a["shell_atom"], np.int32)
        vec(f, a["origin"], np.float64)
        vec(f, a["exps"], np.float64)
        vec(f, a["coefs"], np.float64)
        vec(f, spts, np.float64)
        vec(f, sw, np.float64)
        raw = np.stack([leaves["offset"].astype(float), leaves["count"].astype(float), leaves["center"][:, 0],
                        leaves["center"][:, 1], leaves["center"][:, 2], leaves["radius"]], 1)
        vec(f, raw, np.float64)
        vec(f, np.asfortranarray(D).ravel(order="F"), np.float64)
        # mo.Cocc blocks (column-major, alpha then beta) and the occupied counts; empty = D-only handle
        if Cocc is None:
            vec(f, [], np.float64)
            vec(f, [], np.int32)
        else:
            vec(f, np.concatenate([np.asfortranarray(c).ravel(order="F") for c in Cocc]), np.float64)
            vec(f, [Cocc[0].shape[1], Cocc[-1].shape[1]], np.int32)
        if perm is not None:          # NLC cases: sort permutation and atom-grid prefix sums
            vec(f, perm, np.int64)
            vec(f, atom_offsets, np.int64)


def test_adapter_compiles_and_reports_errors_as_exceptions(case, tmp_path):
    import torch
    exe = _compile()
    c = case("water", "def2-svp")
    b = c["basis"]
    D = 0.5 * __import__("occ_b200").molecules.synthetic_density(b, 1)
    p = str(tmp_path / "case.bin")
    if torch.cuda.is_available():
        # force a failure: a leaf that runs past the point array
        leaves = c["leaves"].copy()
        leaves["count"][-1] += 10_000
        _write_case(p, b, 1, [101, 130], c["spts"], c["sw"], leaves, D)
    else:
        _write_case(p, b, 1, [101, 130], c["spts"], c["sw"], c["leaves"], D)   # fails: no CUDA device, no fallback
    r = subprocess.run([exe, p, str(tmp_path / "out.bin"), "--expect-error"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "runtime_error" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0", "0,1", "0,1,2,3,4,5,6,7"])
@pytest.mark.parametrize("method,nspin", [("b3lyp", 1), ("pbe", 2)])
def test_adapter_build_matches_oracle(case, oracle, tmp_path, method, nspin, devices):
    """XCBuild::compute_K_dft / compute_K_dft_mo with the reference's own conventions (mo.D undoubled, mo.Cocc
    unscaled: the factors are applied on the device); `devices`: the single-process multi-GPU handle (ngpu = 2..8)."""
    import torch
    from occ_b200 import dft, molecules
    if devices is not None and max(int(x) for x in devices.split(",")) >= torch.cuda.device_count():
        pytest.skip("not enough GPUs")
    exe = _compile()
    c = case("water", "def2-svp")
    b = c["basis"]
    Cs = molecules.synthetic_orbitals(b, nspin)                 # D2_s = C_s C_s^T
    D2 = np.asfortranarray(np.vstack([x @ x.T for x in Cs]))
    Cocc = [x / np.sqrt(2.0) for x in Cs] if nspin == 1 else Cs  # restricted: mo.D = Cocc Cocc^T = D2 / 2
    p, o = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    _write_case(p, b, nspin, dft.get_dft_method(method), c["spts"], c["sw"], c["leaves"], 0.5 * D2, Cocc)  # mo.D = D2 / 2
    r = subprocess.run([exe, p, o] + (["--devices", devices] if devices else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(o)
    K = raw[3:].reshape(b.nbf, nspin * b.nbf).T
    names = [n for i in dft.get_dft_method(method) for n, v in oracle.FUNC_IDS.items() if v == i]
    ref = oracle.xc_build(c["obasis"], c["spts"], c["sw"], c["leaves"], names, nspin, D2)
    assert np.abs(K - ref["K"]).max() < 1e-9 and abs(raw[0] - ref["exc"]) < 1e-8
    assert abs(raw[1] - ref["n_alpha"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("method,nspin", [("pbe", 1), ("tpss", 2)])
def test_adapter_xc_gradient_matches_oracle(case, oracle, tmp_path, method, nspin):
    """XCBuild::compute_xc_gradient / _mo with mo.D and mo.Cocc as the reference stores them (the factor 2 of
    dft.h:606 and the sqrt(2) of the restricted orbitals are applied on the device)"""
    from occ_b200 import dft, molecules
    exe = _compile()
    c = case("water", "def2-svp")
    b = c["basis"]
    Cs = molecules.synthetic_orbitals(b, nspin)
    D2 = np.asfortranarray(np.vstack([x @ x.T for x in Cs]))
    Cocc = [x / np.sqrt(2.0) for x in Cs] if nspin == 1 else Cs
    p, o = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    _write_case(p, b, nspin, dft.get_dft_method(method), c["spts"], c["sw"], c["leaves"], 0.5 * D2, Cocc)
    r = subprocess.run([exe, p, o, "--gradient"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(o)
    g = raw[3 + nspin * b.nbf * b.nbf:].reshape(-1, 3)
    names = [n for i in dft.get_dft_method(method) for n, v in oracle.FUNC_IDS.items() if v == i]
    gref = oracle.xc_gradient(c["obasis"], c["spts"], c["sw"], c["leaves"], names, nspin, D2, len(g))
    assert np.abs(g - gref).max() < 1e-8 and np.abs(gref).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0", "0,1"])
def test_adapter_nonlocal_correlation_matches_oracle(oracle, tmp_path, devices):
    """adapter NonLocalCorrelation (vv10 + compute_gradient) over the NLC grid, mo.D undoubled; one device or shards"""
    import torch
    from occ_b200 import aobasis, grid, molecules
    if devices is not None and max(int(x) for x in devices.split(",")) >= torch.cuda.device_count():
        pytest.skip("not enough GPUs")
    exe = _compile()
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    ob = oracle.Basis(b.as_arrays())
    mg = grid.MolecularGrid(b, grid.GridSettings(110, 50, 1e-7, True, True))
    h, perm = mg.hierarchy(with_permutation=True)
    pts, w = mg.points()
    off = np.concatenate([[0], np.cumsum([mg.atom_info(a)[0] for a in range(mg.natoms)])])
    mg.close()
    D2 = molecules.synthetic_density(b, 1)
    p, o = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    _write_case(p, b, 1, [106], h.sorted_points, h.sorted_weights, h.leaves, 0.5 * D2, None, perm, off)
    r = subprocess.run([exe, p, o, "--nlc"] + (["--devices", devices] if devices else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(o)
    n3 = 3 * len(Z)
    e, eg, g, K = raw[0], raw[1], raw[2:2 + n3].reshape(-1, 3), raw[2 + n3:].reshape(b.nbf, b.nbf).T
    ref = oracle.vv10_gradient(ob, pts, w, off.astype(np.uint64), D2)
    rk = oracle.vv10(ob, pts, w, D2)
    assert abs(e - rk["energy"]) < 1e-8 and abs(eg - e) < 1e-12
    assert np.abs(K - rk["K"]).max() < 1e-9
    assert np.abs(g - ref["gradient"]).max() < 1e-9


def test_cosx_adapter_compiles_and_fails_loudly_without_device():
    """adapter/occ_cosx_b200.h (SemiNumericalExchange mirror): g++ compile + link; no device -> std::runtime_error."""
    src = os.path.join(ROOT, "tests", "cpp", "adapter_cosx_driver.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "adapter_cosx_driver")
    lib = os.path.join(ROOT, "occ_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-o", exe, src, "-L" + lib, "-locc_xc_b200",
                           "-Wl,-rpath," + lib])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    import torch
    if torch.cuda.is_available():
        assert out.stdout.startswith("ok K00=")
    else:
        assert "runtime_error" in out.stdout and "CUDA" in out.stdout
