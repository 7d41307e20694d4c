"""The C-ABI library loads and exports every symbol include/occ_xc_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "occ_xc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(occ_xc_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = header_symbols()
    for s in ("occ_xc_create", "occ_xc_build", "occ_xc_build_device", "occ_xc_destroy", "occ_xc_last_error"):
        assert s in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    from occ_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == header_symbols()
    L.occ_xc_abi_version.restype = ctypes.c_int
    assert L.occ_xc_abi_version() == 2


def test_pure_host_entry_points():
    """family / exact-exchange tables and the error channel need no device"""
    from occ_b200 import _lib
    L = _lib.lib()
    assert [L.occ_xc_family(i) for i in (1, 7, 8, 30, 101, 106, 130, 131, 402, 497, 498, 999)] == \
        [0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, -1]
    assert L.occ_xc_exact_exchange(402) == 0.20 and L.occ_xc_exact_exchange(101) == 0.0
    assert L.occ_xc_decay_cutoff(None, ctypes.c_double(1e-10), None) != 0
    assert b"null" in L.occ_xc_last_error()


def test_no_cpu_fallback_without_device():
    """create() must fail loudly when there is no CUDA device (this container) instead of computing on the CPU"""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        return
    from occ_b200 import aobasis, dft, molecules
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    g = dft.GridHierarchy(np.zeros((1, 3)), np.ones(1), np.zeros(1, dtype=dft.LEAF_DTYPE))
    try:
        dft.DFT("pbe", b, g)
    except RuntimeError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("DFT() succeeded without a GPU")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "occ_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), os.path.join(dirpath, f)


def test_header_is_plain_c_and_links_from_c99():
    """include/occ_xc_b200.h compiled by gcc -std=c99 -pedantic and linked against the library: the boundary is a C ABI"""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "abi_c99.c")
    exe = os.path.join(ROOT, "tests", "cpp", "abi_c99")
    lib = os.path.join(ROOT, "occ_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-o", exe, src, "-L" + lib,
                           "-locc_xc_b200", "-Wl,-rpath," + lib])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def cosx_header_symbols():
    text = open(os.path.join(ROOT, "include", "occ_cosx_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(occ_cosx_[a-zA-Z_0-9]+)\s*\(", text)))


def test_cosx_library_exports_every_declared_symbol():
    from occ_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = cosx_header_symbols()
    assert "occ_cosx_create" in syms and "occ_cosx_compute_K" in syms
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(_lib.COSX_SYMBOLS) == syms


def test_cosx_header_is_plain_c():
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "abi_cosx_c99.c")
    exe = os.path.join(ROOT, "tests", "cpp", "abi_cosx_c99")
    lib = os.path.join(ROOT, "occ_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-o", exe, src, "-L" + lib,
                           "-locc_xc_b200", "-Wl,-rpath," + lib])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def test_cosx_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    import numpy as np
    from occ_b200 import aobasis, cosx, molecules
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    offs = np.array([0, 1, 2, 3], dtype=np.uint64)
    try:
        cosx.SemiNumericalExchange(b, atom_grids=(offs, np.zeros((3, 3)), np.ones(3)), projector=None)
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("SemiNumericalExchange() succeeded without a GPU")
