"""Host-side mirror of occ::dft::NonLocalCorrelationFunctional (include/occ/dft/nonlocal_correlation.h,
src/dft/nonlocal_correlation.cpp): the post-SCF VV10 correction, SURVEY.md 8(f) rank 4.

    nlc = NonLocalCorrelationFunctional()
    nlc.set_parameters(vv10_b=6.0, vv10_C=0.01)
    nlc.set_integration_grid(basis)              # default NLC grid: 110 / 50 angular, LMG 1e-7, H / He one Lebedev level lower
    energy, Vxc = nlc(D)                         # D = mo.D (restricted); operator()(basis, mo) of the reference

    g = nlc.compute_gradient(D)                  # GradientResult::gradient (natoms, 3): vv10_gradient of the reference

`devices=[0, 1, ...]` in set_integration_grid runs the same calls on several GPUs of this process
(occ_xc_multi_vv10 / occ_xc_multi_vv10_gradient).  All arithmetic runs in the CUDA library
(include/occ_xc_b200.h); no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, ptr
from .aobasis import AOBasis
from .dft import DFT, MultiGPUDFT
from .grid import GridSettings, MolecularGrid


class NonLocalCorrelationFunctional:
    def __init__(self):
        self.vv10_b, self.vv10_C = 6.0, 0.01      # Parameters defaults, nonlocal_correlation.h:14-18
        self._dft = None
        self._orig = None

    def set_parameters(self, vv10_b: float = 6.0, vv10_C: float = 0.01):
        self.vv10_b, self.vv10_C = float(vv10_b), float(vv10_C)

    def set_integration_grid(self, basis: AOBasis, settings: GridSettings | None = None, device: int = -1,
                             screening_threshold: float = 0.0, gradient: bool = False, devices=None):
        """set_integration_grid (nonlocal_correlation.cpp:17-31).  screening_threshold 0 keeps every shell at every
        point like the reference's dense evaluate_basis; 1e-10 screens per Morton leaf like the XC build."""
        self.close()
        # nonlocal_correlation.h:37-40 passes {110, 50, 50, 1e-7, false}; the struct's fifth member is int_acc
        # (grid_settings.h:81-90), so reduced_first_row_element_grid keeps its default TRUE in the reference — H and He get 86
        # angular points.  Pinned by the reference's H2 golden (tests/dft_tests.cpp:608-683), met to 2e-11 only this way.
        s = settings or GridSettings(110, 50, 1e-7, True, True)
        mg = MolecularGrid(basis, s)
        h, perm = mg.hierarchy(with_permutation=True)
        counts = [mg.atom_info(a)[0] for a in range(mg.natoms)]
        self._atom_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)   # m_nlc_atom_grids ranges
        mg.close()
        self.basis = basis
        self._orig = np.ascontiguousarray(perm, dtype=np.uint64)      # sorted index -> atom-grid order
        self.num_grid_points = len(h.sorted_weights)
        self._multi = devices is not None
        if self._multi:
            self._dft = MultiGPUDFT("gga_x_b88", basis, h, devices=list(devices), gradient=gradient,
                                    xc_screening_threshold=screening_threshold)
        else:
            self._dft = DFT("gga_x_b88", basis, h, device=device, xc_screening_threshold=screening_threshold,
                            gradient=gradient)
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
        if self._multi:
            check(lib().occ_xc_multi_vv10(self._dft._m, ptr(D2), ptr(self._orig), C.c_double(self.vv10_b),
                                          C.c_double(self.vv10_C), ptr(K), C.byref(e)))
        else:
            check(lib().occ_xc_vv10(self._dft._h, ptr(D2), ptr(self._orig), C.c_double(self.vv10_b),
                                    C.c_double(self.vv10_C), ptr(K), C.byref(e)))
        return e.value, K

    def compute_gradient(self, D: np.ndarray):
        """GradientResult compute_gradient(basis, mo) (nonlocal_correlation.cpp:530-534): D = mo.D; returns (natoms, 3)"""
        return self.vv10_gradient(np.asfortranarray(2.0 * np.asarray(D, dtype=np.float64)))[1]

    def vv10_gradient(self, D2: np.ndarray):
        """vv10_gradient (nonlocal_correlation.cpp:392-522) for D2 = 2 mo.D; returns (energy, gradient (natoms, 3));
        needs set_integration_grid(..., gradient=True)"""
        if self._dft is None:
            raise RuntimeError("NonLocalCorrelationFunctional: set_integration_grid has not been called")
        nbf = self.basis.nbf
        D2 = np.asfortranarray(D2, dtype=np.float64)
        if D2.shape != (nbf, nbf):
            raise RuntimeError("Only restricted NLC functional gradients implemented right now")  # :397-399
        g = np.zeros((len(self._atom_offsets) - 1, 3))
        e = C.c_double()
        if self._multi:
            check(lib().occ_xc_multi_vv10_gradient(self._dft._m, ptr(D2), ptr(self._orig), ptr(self._atom_offsets),
                                                   C.c_double(self.vv10_b), C.c_double(self.vv10_C), ptr(g), C.byref(e)))
        else:
            check(lib().occ_xc_vv10_gradient(self._dft._h, ptr(D2), ptr(self._orig), ptr(self._atom_offsets),
                                             C.c_double(self.vv10_b), C.c_double(self.vv10_C), ptr(g), C.byref(e)))
        return e.value, g

    def close(self):
        if getattr(self, "_dft", None) is not None:
            self._dft.close()
            self._dft = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
