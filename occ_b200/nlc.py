"""Host-side mirror of occ::dft::NonLocalCorrelationFunctional (include/occ/dft/nonlocal_correlation.h,
src/dft/nonlocal_correlation.cpp): the post-SCF VV10 correction, SURVEY.md 8(f) rank 4.

    nlc = NonLocalCorrelationFunctional()
    nlc.set_parameters(vv10_b=6.0, vv10_C=0.01)
    nlc.set_integration_grid(basis)              # default NLC grid {110, 50, 50, 1e-7, no first-row reduction}
    energy, Vxc = nlc(D)                         # D = mo.D (restricted); operator()(basis, mo) of the reference

All arithmetic runs in the CUDA library (occ_xc_vv10, include/occ_xc_b200.h); no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, ptr
from .aobasis import AOBasis
from .dft import DFT
from .grid import GridSettings, MolecularGrid


class NonLocalCorrelationFunctional:
    def __init__(self):
        self.vv10_b, self.vv10_C = 6.0, 0.01      # Parameters defaults, nonlocal_correlation.h:14-18
        self._dft = None
        self._orig = None

    def set_parameters(self, vv10_b: float = 6.0, vv10_C: float = 0.01):
        self.vv10_b, self.vv10_C = float(vv10_b), float(vv10_C)

    def set_integration_grid(self, basis: AOBasis, settings: GridSettings | None = None, device: int = -1,
                             screening_threshold: float = 0.0):
        """set_integration_grid (nonlocal_correlation.cpp:17-31).  screening_threshold 0 keeps every shell at every
        point like the reference's dense evaluate_basis; 1e-10 screens per Morton leaf like the XC build."""
        self.close()
        s = settings or GridSettings(110, 50, 1e-7, False, True)
        mg = MolecularGrid(basis, s)
        h, perm = mg.hierarchy(with_permutation=True)
        mg.close()
        self.basis = basis
        self._orig = np.ascontiguousarray(perm, dtype=np.uint64)      # sorted index -> atom-grid order
        self.num_grid_points = len(h.sorted_weights)
        self._dft = DFT("gga_x_b88", basis, h, device=device, xc_screening_threshold=screening_threshold)
        return self

    def __call__(self, D: np.ndarray):
        """Result operator()(basis, mo): D = mo.D of a restricted wavefunction; returns (energy, Vxc)."""
        return self.vv10(np.asfortranarray(2.0 * np.asarray(D, dtype=np.float64)))

    def vv10(self, D2: np.ndarray):
        if self._dft is None:
            raise RuntimeError("NonLocalCorrelationFunctional: set_integration_grid has not been called")
        nbf = self.basis.nbf
        D2 = np.asfortranarray(D2, dtype=np.float64)
        if D2.shape != (nbf, nbf):
            raise RuntimeError("Only restricted NLC functional implemented right now")   # nonlocal_correlation.cpp:370
        K = np.zeros((nbf, nbf), order="F")
        e = C.c_double()
        check(lib().occ_xc_vv10(self._dft._h, ptr(D2), ptr(self._orig), C.c_double(self.vv10_b), C.c_double(self.vv10_C),
                                ptr(K), C.byref(e)))
        return e.value, K

    def close(self):
        if getattr(self, "_dft", None) is not None:
            self._dft.close()
            self._dft = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
