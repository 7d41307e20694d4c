"""Host-side mirror of occ::numint::MolecularGrid / SpatialGridHierarchy for the XC build's inputs
(reference: include/occ/numint/molecular_grid.h, grid_settings.h:81-95, spatial_grid_hierarchy.h:33-78).

The work is done by the native library (occ_b200/csrc/grid.cu: host C++ for the radial/angular construction and
the Morton sort, a CUDA kernel for the Becke partition); this module only carries arrays across the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import LEAF_DTYPE, BasisArrays, GridSettingsT, check, lib, ptr
from .aobasis import AOBasis
from .dft import GridHierarchy


@dataclass
class GridSettings:
    """include/occ/numint/grid_settings.h:81-95 defaults"""
    max_angular_points: int = 302
    min_angular_points: int = 110
    radial_precision: float = 1e-12
    reduced_first_row_element_grid: bool = True
    treutler_alrichs_adjustment: bool = True
    pruning_scheme: int = 0            # 0 NWChem, 1 None (uniform angular order: GridSettings::for_sgx)
    int_acc: float = 0.0               # ORCA IntAcc (GridSettings::for_cosx)
    angular_regions: tuple = (0, 0, 0, 0, 0)   # ORCA 5-region angular orders; [0] > 0 selects the COSX grid producer

    @staticmethod
    def for_sgx(angular_points: int = 50) -> "GridSettings":
        """GridSettings::for_sgx (src/numint/grid_settings.cpp:59-69)"""
        return GridSettings(angular_points, angular_points, 1e-5, False, True, 1)

    @staticmethod
    def for_cosx(level: int = 1) -> "GridSettings":
        """GridSettings::for_cosx(COSXGridLevel) (src/numint/grid_settings.cpp:10-18,71-103): ORCA GRIDX 1 / 2 / 3 —
        IntAcc radial count, Treutler-Ahlrichs radial nodes, 5-region angular pruning, H / He one Lebedev level lower"""
        table = {1: (3.816, (14, 26, 50, 50, 26), 50, 14), 2: (4.020, (14, 26, 50, 110, 50), 110, 14),
                 3: (4.338, (26, 50, 110, 194, 110), 194, 26)}
        ia, regions, mx, mn = table[int(level)]
        return GridSettings(mx, mn, 1e-12, True, True, 0, ia, regions)

    @staticmethod
    def from_grid_quality(quality: str) -> "GridSettings":
        """GridSettings::from_grid_quality (src/numint/grid_settings.cpp:32-58); radial_points is ignored by LMG and
        radial_precision keeps its default"""
        table = {"coarse": (170, 110), "standard": (194, 110), "fine": (302, 194), "veryfine": (590, 302)}
        mx, mn = table[quality.lower()]
        return GridSettings(mx, mn)


class MolecularGrid:
    def __init__(self, basis: AOBasis, settings: GridSettings | None = None):
        self.settings = settings or GridSettings()
        s = self.settings
        st = GridSettingsT(s.max_angular_points, s.min_angular_points, s.radial_precision,
                           int(s.reduced_first_row_element_grid), int(s.treutler_alrichs_adjustment),
                           int(s.pruning_scheme), float(s.int_acc), (C.c_int * 5)(*[int(x) for x in s.angular_regions]))
        ba = BasisArrays(basis.as_arrays())
        Z = np.ascontiguousarray(basis.atomic_numbers, dtype=np.int32)
        pos = np.ascontiguousarray(basis.positions, dtype=np.float64)
        self._g = C.c_void_p()
        check(lib().occ_xc_grid_create(C.c_int(len(Z)), ptr(Z), ptr(pos), ba.ref, C.byref(st), C.byref(self._g)))
        self.natoms = len(Z)

    @staticmethod
    def from_points(points: np.ndarray, weights: np.ndarray) -> "MolecularGrid":
        self = MolecularGrid.__new__(MolecularGrid)
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self._g = C.c_void_p()
        self.natoms = 0
        check(lib().occ_xc_grid_from_points(C.c_size_t(len(w)), ptr(pts), ptr(w), C.byref(self._g)))
        return self

    def num_points(self) -> int:
        return lib().occ_xc_grid_npts(self._g)

    def atom_info(self, atom: int):
        n, nr = C.c_size_t(), C.c_int()
        check(lib().occ_xc_grid_atom_info(self._g, C.c_int(atom), C.byref(n), C.byref(nr)))
        return n.value, nr.value

    def points(self):
        """MolecularGridPoints::points()/weights() (atom order)"""
        n = self.num_points()
        pts, w = np.zeros((n, 3)), np.zeros(n)
        check(lib().occ_xc_grid_copy(self._g, None, None, None, None, ptr(pts), ptr(w)))
        return pts, w

    def hierarchy(self, with_permutation: bool = False):
        """MolecularGridPoints::get_hierarchy()"""
        n, nl = self.num_points(), lib().occ_xc_grid_nleaves(self._g)
        spts, sw = np.zeros((n, 3)), np.zeros(n)
        leaves = np.zeros(nl, dtype=LEAF_DTYPE)
        perm = np.zeros(n, dtype=np.uint64)
        check(lib().occ_xc_grid_copy(self._g, ptr(spts), ptr(sw), ptr(leaves), ptr(perm), None, None))
        h = GridHierarchy(spts, sw, leaves)
        return (h, perm) if with_permutation else h

    def close(self):
        if getattr(self, "_g", None) is not None and self._g.value:
            lib().occ_xc_grid_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
