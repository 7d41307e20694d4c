"""ctypes binding of the C ABI (include/occ_xc_b200.h) exported by occ_b200/libocc_xc_b200.so.

The product path has no CPU fallback: if the CUDA library is missing or cannot be loaded this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCC_XC_B200_LIB", os.path.join(HERE, "libocc_xc_b200.so"))  # override: kernel-variant experiments

# every symbol include/occ_xc_b200.h declares
SYMBOLS = ["occ_xc_create", "occ_xc_build", "occ_xc_build_device", "occ_xc_build_mo", "occ_xc_build_mo_device", "occ_xc_gradient", "occ_xc_gradient_mo", "occ_xc_vv10", "occ_xc_density_on_grid", "occ_xc_pinned_alloc",
           "occ_xc_pinned_free", "occ_xc_host_register", "occ_xc_host_unregister", "occ_xc_get_info", "occ_xc_set_profiling",
           "occ_xc_get_phase_ms", "occ_xc_decay_cutoff", "occ_xc_eval_functional", "occ_xc_leaf_collocation",
           "occ_xc_family", "occ_xc_exact_exchange", "occ_xc_range_separation", "occ_xc_options_init", "occ_xc_destroy", "occ_xc_last_error", "occ_xc_abi_version",
           "occ_xc_partition_leaves", "occ_xc_register_lebedev", "occ_xc_grid_create", "occ_xc_grid_from_points", "occ_xc_grid_npts",
           "occ_xc_grid_nleaves", "occ_xc_grid_atom_info", "occ_xc_grid_copy", "occ_xc_grid_destroy",
           "occ_xc_measure_fp64_tensor_peak", "occ_xc_multi_create", "occ_xc_multi_build", "occ_xc_multi_ndev", "occ_xc_multi_get_info", "occ_xc_multi_destroy",
           "occ_xc_vv10_gradient", "occ_xc_multi_vv10", "occ_xc_multi_vv10_gradient", "occ_xc_multi_gradient"]
# every symbol include/occ_cosx_b200.h declares
COSX_SYMBOLS = ["occ_cosx_create", "occ_cosx_compute_K", "occ_cosx_compute_K_range_separated", "occ_cosx_compute_K_mo",
                "occ_cosx_compute_K_range_separated_mo", "occ_cosx_compute_K_range_separated_kind",
                "occ_cosx_compute_K_range_separated_mo_kind", "occ_cosx_K_for_density",
                "occ_cosx_K_for_density_device", "occ_cosx_numerical_overlap", "occ_cosx_analytic_overlap",
                "occ_cosx_set_projector", "occ_cosx_get_info", "occ_cosx_destroy",
                "occ_cosx_multi_create", "occ_cosx_multi_compute_K", "occ_cosx_multi_compute_K_range_separated",
                "occ_cosx_multi_numerical_overlap", "occ_cosx_multi_analytic_overlap", "occ_cosx_multi_set_projector",
                "occ_cosx_multi_ndev", "occ_cosx_multi_get_info", "occ_cosx_multi_destroy"]


class BasisT(C.Structure):
    _fields_ = [("nsh", C.c_int), ("pure", C.c_int), ("l", C.c_void_p), ("origin", C.c_void_p), ("nprim", C.c_void_p),
                ("prim_off", C.c_void_p), ("exps", C.c_void_p), ("coefs", C.c_void_p), ("shell_atom", C.c_void_p)]


class LeafT(C.Structure):
    _fields_ = [("offset", C.c_size_t), ("count", C.c_size_t), ("center", C.c_double * 3), ("radius", C.c_double)]


class OptionsT(C.Structure):
    """occ_xc_options_t (ABI version 2)"""
    _fields_ = [("device", C.c_int), ("rank", C.c_int), ("world_size", C.c_int), ("batch_bytes", C.c_size_t),
                ("mo_form", C.c_int), ("nocc_alpha", C.c_int), ("nocc_beta", C.c_int), ("gradient", C.c_int),
                ("deterministic", C.c_int), ("nvtx", C.c_int), ("lanes", C.c_int), ("debug", C.c_uint),
                ("density_scale", C.c_double), ("orbital_scale", C.c_double), ("reserved", C.c_int * 6)]


class InfoT(C.Structure):
    _fields_ = [("nbf", C.c_int), ("nspin", C.c_int), ("deriv", C.c_int), ("nsh", C.c_int), ("npts", C.c_size_t),
                ("nleaves", C.c_size_t), ("ntiles", C.c_size_t), ("nbatches", C.c_size_t), ("sum_n_nb2", C.c_double),
                ("flops_per_build", C.c_double), ("nb_min", C.c_int), ("nb_max", C.c_int),
                ("workspace_bytes", C.c_size_t), ("launches_per_build", C.c_size_t), ("mo_density_flops", C.c_double),
                ("launches_per_build_mo", C.c_size_t), ("vxc_flops", C.c_double)]


class GridSettingsT(C.Structure):
    _fields_ = [("max_angular_points", C.c_int), ("min_angular_points", C.c_int), ("radial_precision", C.c_double),
                ("reduced_first_row_element_grid", C.c_int), ("treutler_alrichs_adjustment", C.c_int),
                ("pruning_scheme", C.c_int), ("int_acc", C.c_double), ("angular_regions", C.c_int * 5)]


LEAF_DTYPE = np.dtype([("offset", np.uint64), ("count", np.uint64), ("center", np.float64, 3), ("radius", np.float64)])
NPHASES = 7
PHASE_NAMES = ["prepare_D", "collocation", "density", "functional", "gbuild", "vxc", "finalise"]

_lib = None


def lib():
    """Load libocc_xc_b200.so; raise loudly if it is not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (make -C occ_b200/csrc). occ_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.occ_xc_last_error.restype = C.c_char_p
        L.occ_xc_exact_exchange.restype = C.c_double
        L.occ_xc_destroy.restype = None
        L.occ_xc_destroy.argtypes = [C.c_void_p]
        L.occ_xc_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.occ_xc_build_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.occ_xc_get_info.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_xc_gradient.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.occ_xc_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.occ_xc_get_phase_ms.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_xc_grid_npts.restype = C.c_size_t
        L.occ_xc_grid_nleaves.restype = C.c_size_t
        L.occ_xc_grid_npts.argtypes = [C.c_void_p]
        L.occ_xc_grid_nleaves.argtypes = [C.c_void_p]
        L.occ_xc_grid_destroy.restype = None
        L.occ_xc_grid_destroy.argtypes = [C.c_void_p]
        L.occ_xc_grid_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.occ_xc_grid_atom_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.occ_xc_options_init.restype = None
        L.occ_xc_multi_destroy.restype = None
        L.occ_xc_multi_destroy.argtypes = [C.c_void_p]
        L.occ_xc_multi_build.argtypes = [C.c_void_p] * 7
        L.occ_xc_multi_get_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        tables = np.load(os.path.join(HERE, "data", "lebedev_tables.npz"))
        for k in tables.files:
            a = np.ascontiguousarray(tables[k], dtype=np.float64)
            check_rc(L, L.occ_xc_register_lebedev(C.c_int(a.shape[0]), a.ctypes.data_as(C.c_void_p)))
        _lib = L
    return _lib


def check_rc(L, rc: int):
    if rc != 0:
        raise RuntimeError(L.occ_xc_last_error().decode())


def check(rc: int):
    if rc != 0:
        raise RuntimeError(lib().occ_xc_last_error().decode())


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class BasisArrays:
    """Keeps the numpy arrays behind an occ_xc_basis_t alive."""

    def __init__(self, arrays: dict):
        self.a = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in arrays.items()}
        a = self.a
        self.t = BasisT(a["nsh"], a["pure"], ptr(a["l"]).value, ptr(a["origin"]).value, ptr(a["nprim"]).value,
                        ptr(a["prim_off"]).value, ptr(a["exps"]).value, ptr(a["coefs"]).value,
                        ptr(a["shell_atom"]).value)

    @property
    def ref(self):
        return C.byref(self.t)
