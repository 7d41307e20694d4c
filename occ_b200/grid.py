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
                           int(s.reduced_first_row_element_grid), int(s.treutler_alrichs_adjustment))
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
