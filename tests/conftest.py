import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def build_case(mol, basis_name, pure=False):
    """(Z, pos, AOBasis, oracle Basis, oracle grid hierarchy) for a named molecule — oracle-side inputs."""
    from occ_b200 import aobasis, molecules
    from oracle import oracle as O
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, basis_name, pure=pure)
    ob = O.Basis(b.as_arrays())
    pts, w, info = O.molecular_grid(Z, pos, ob)
    spts, sw, perm, leaves = O.hierarchy(pts, w)
    return {"Z": Z, "pos": pos, "basis": b, "obasis": ob, "pts": pts, "w": w, "spts": spts, "sw": sw,
            "perm": perm, "leaves": leaves, "info": info}


_CASES = {}


@pytest.fixture(scope="session")
def case():
    def get(mol, basis_name, pure=False):
        key = (mol, basis_name, pure)
        if key not in _CASES:
            _CASES[key] = build_case(mol, basis_name, pure)
        return _CASES[key]
    return get
