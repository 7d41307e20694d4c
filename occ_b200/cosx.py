"""Host-side mirror of occ::qm::cosx::SemiNumericalExchange (include/occ/qm/seminumerical_exchange.h:104-205,
src/qm/seminumerical_exchange.cpp), ESP mode: the seminumerical (COSX) exchange build, SURVEY.md 8(f) rank 2.

    sgx = SemiNumericalExchange(basis, GridSettings.for_sgx(50))     # or atom_grids=(offsets, points, weights)
    K = sgx.compute_K(D)                                             # D = mo.D (restricted: nbf x nbf)
    K = sgx.compute_K(D, unrestricted=True)                          # D: (2 nbf x nbf), alpha block on top
    sgx.set_range_separated_omega(0.3); K_lr = sgx.compute_K(D)      # erf(omega r)/r
    K = sgx.compute_K_range_separated(D, alpha, beta, omega)         # fused two-operator sweep

Everything numerical runs in the CUDA library (include/occ_cosx_b200.h); the overlap projector is formed here from
the library's numerical and analytic overlaps (seminumerical_exchange.cpp:258-262). No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import BasisArrays, check, lib, ptr
from .aobasis import AOBasis
from .grid import GridSettings, MolecularGrid


class CosxOptionsT(C.Structure):
    _fields_ = [("margin", C.c_double), ("density_threshold", C.c_double), ("pair_threshold", C.c_double),
                ("batch_bytes", C.c_size_t), ("device", C.c_int), ("collect_stats", C.c_int), ("rank", C.c_int),
                ("world_size", C.c_int), ("reserved", C.c_int * 4)]


class CosxInfoT(C.Structure):
    _fields_ = [("nbf", C.c_int), ("nsh", C.c_int), ("natoms", C.c_int), ("npts", C.c_size_t), ("nbatches", C.c_size_t),
                ("npairs", C.c_size_t), ("nprimitive_pairs", C.c_size_t), ("device_bytes", C.c_size_t),
                ("pair_evals", C.c_size_t), ("skipped_batches", C.c_size_t), ("dense_flops", C.c_double),
                ("gemm_flops", C.c_double), ("ms_collocate", C.c_double), ("ms_fgemm", C.c_double),
                ("ms_esp", C.c_double), ("ms_kgemm", C.c_double), ("launches", C.c_size_t), ("gemm_calls", C.c_size_t), ("kgemm_blocks", C.c_size_t)]


def sgx_grid_settings(angular_points: int = 50, radial_precision: float = 1e-5) -> GridSettings:
    """GridSettings::for_sgx (src/numint/grid_settings.cpp:59-69): one angular order for every shell, loose radial
    precision, no first-row reduction. The native producer prunes with the NWChem scheme, which is uniform below 50
    points (grid_utils.cpp:240-243); at 50 and above the innermost shells are thinned."""
    return GridSettings(angular_points, angular_points, radial_precision, False, True)


def cosx_grid_settings(level: int = 1) -> GridSettings:
    """GridSettings::for_cosx(level) — the grid HartreeFock::set_cosx_exchange builds the engine on (src/qm/hf.cpp:62-68)"""
    return GridSettings.for_cosx(level)


def batch_shard(atom_offsets, rank: int, world_size: int):
    """(start, count) of the grid batches rank `rank` owns: blocks of 128 consecutive points of each atom grid
    (seminumerical_exchange.cpp:676-699), numbered over the whole molecule and dealt round-robin — the rule
    occ_cosx_create applies for options.rank / world_size."""
    out, g = [], 0
    for a in range(len(atom_offsets) - 1):
        n = int(atom_offsets[a + 1]) - int(atom_offsets[a])
        for s in range(0, n, 128):
            if g % world_size == rank:
                out.append((int(atom_offsets[a]) + s, min(128, n - s)))
            g += 1
    return out


class SemiNumericalExchange:
    def __init__(self, basis: AOBasis, settings: GridSettings | None = None, atom_grids=None, device: int = -1,
                 margin: float = 1.0, collect_stats: bool = False, batch_bytes: int = 0, rank: int = 0,
                 world_size: int = 1, projector: np.ndarray | None | str = "auto", kgemm_library: bool = False):
        """SemiNumericalExchange(basis, grid settings = {}) (seminumerical_exchange.cpp:232-275): like the reference's
        constructor, no settings means the default-constructed GridSettings (the 302 / 110 DFT grid); the SCF front end
        passes GridSettings::for_cosx(level) (src/qm/hf.cpp:63-64) = cosx_grid_settings(level) here.  atom_grids, when
        given, is (offsets[natoms+1], points (N,3), weights (N,)) — the reference's m_atom_grids."""
        L = lib()
        self.basis = basis
        self.omega = 0.0
        if atom_grids is None:
            mg = MolecularGrid(basis, settings or GridSettings())
            pts, w = mg.points()
            counts = [mg.atom_info(a)[0] for a in range(mg.natoms)]
            mg.close()
            offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        else:
            offs, pts, w = atom_grids
        self.atom_offsets = np.ascontiguousarray(offs, dtype=np.uint64)
        self.points = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.weights = np.ascontiguousarray(w, dtype=np.float64)
        self._ba = BasisArrays(basis.as_arrays())
        opt = CosxOptionsT(margin, 0.0, 0.0, batch_bytes, device, int(collect_stats), rank, world_size)
        if kgemm_library:
            raise RuntimeError("the cuBLAS A/B path of K' += Phi^T G was removed in round 2 (no library GEMM is left)")
        self._h = C.c_void_p()
        L.occ_cosx_destroy.restype = None
        L.occ_cosx_destroy.argtypes = [C.c_void_p]
        for f in ("occ_cosx_compute_K", "occ_cosx_K_for_density", "occ_cosx_compute_K_range_separated",
                  "occ_cosx_numerical_overlap", "occ_cosx_analytic_overlap", "occ_cosx_set_projector",
                  "occ_cosx_get_info"):
            getattr(L, f).restype = C.c_int
        L.occ_cosx_compute_K.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.occ_cosx_K_for_density.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.occ_cosx_compute_K_range_separated.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                                         C.c_double, C.c_void_p]
        L.occ_cosx_numerical_overlap.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_cosx_analytic_overlap.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_cosx_set_projector.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_cosx_get_info.argtypes = [C.c_void_p, C.c_void_p]
        P0 = None if isinstance(projector, str) or projector is None else np.asfortranarray(projector, dtype=np.float64)
        check(L.occ_cosx_create(self._ba.ref, C.c_int(len(self.atom_offsets) - 1), ptr(self.atom_offsets),
                                ptr(self.points), ptr(self.weights), ptr(P0) if P0 is not None else None,
                                C.byref(opt), C.byref(self._h)))
        self.overlap_projector = P0
        if isinstance(projector, str):
            # m_overlap_projector = m_numerical_overlap.ldlt().solve(m_overlap)  (seminumerical_exchange.cpp:258-262);
            # with sharded batches the numerical overlap has to be summed over ranks by the caller first
            if world_size > 1:
                raise RuntimeError("SemiNumericalExchange: pass the overlap projector explicitly when sharding")
            S, Sn = self.analytic_overlap(), self.compute_overlap_matrix()
            self.set_projector(np.linalg.solve(Sn, S))

    # -- reference API -------------------------------------------------------------------------------------
    def set_range_separated_omega(self, omega: float):
        self.omega = float(omega)

    def range_separated_omega(self) -> float:
        return self.omega

    def compute_K(self, D: np.ndarray, precision: float = np.finfo(np.float64).eps, unrestricted: bool = False,
                  general: bool = False):
        """compute_K(mo): D = mo.D — (nbf, nbf) restricted, (2 nbf, nbf) unrestricted, (2 nbf, 2 nbf) general"""
        nbf = self.basis.nbf
        kind = 2 if general else (1 if unrestricted else 0)
        shape = ((nbf, nbf), (2 * nbf, nbf), (2 * nbf, 2 * nbf))[kind]
        D = np.asfortranarray(D, dtype=np.float64)
        if D.shape != shape:
            raise RuntimeError(f"compute_K: density matrix of shape {D.shape}, expected {shape}")
        K = np.zeros(shape, order="F")
        check(lib().occ_cosx_compute_K(self._h, kind, ptr(D), precision, self.omega, ptr(K)))
        return K

    def compute_K_mo(self, D: np.ndarray, Cocc, precision: float = np.finfo(np.float64).eps, unrestricted: bool = False):
        """compute_K(mo) with the occupied orbitals as well (mo.D and mo.Cocc both live in `mo`): Cocc is the restricted
        (nbf, nocc) block with 2 D = C C^T, or the pair (C_alpha, C_beta) with 2 D_s = C_s C_s^T. F = Phi.D2q then runs in
        the factorised form (occ_cosx_compute_K_mo)."""
        nbf = self.basis.nbf
        kind = 1 if unrestricted else 0
        shape = ((nbf, nbf), (2 * nbf, nbf))[kind]
        D = np.asfortranarray(D, dtype=np.float64)
        if D.shape != shape:
            raise RuntimeError(f"compute_K_mo: density matrix of shape {D.shape}, expected {shape}")
        blocks = [np.asarray(c, dtype=np.float64).reshape(nbf, -1) for c in (Cocc if unrestricted else (Cocc,))]
        nocc = [b.shape[1] for b in blocks] + [0]
        Cp = np.asfortranarray(np.concatenate(blocks, axis=1)) if sum(nocc) else np.zeros((nbf, 1), order="F")
        K = np.zeros(shape, order="F")
        L = lib()
        L.occ_cosx_compute_K_mo.restype = C.c_int
        L.occ_cosx_compute_K_mo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_double, C.c_double, C.c_void_p]
        check(L.occ_cosx_compute_K_mo(self._h, kind, ptr(D), ptr(Cp), nocc[0], nocc[1], precision, self.omega, ptr(K)))
        return K

    def compute_K_range_separated_mo(self, D, Cocc, alpha: float, beta: float, omega: float,
                                     precision: float = np.finfo(np.float64).eps):
        nbf = self.basis.nbf
        D = np.asfortranarray(D, dtype=np.float64)
        Cp = np.asfortranarray(np.asarray(Cocc, dtype=np.float64).reshape(nbf, -1))
        K = np.zeros((nbf, nbf), order="F")
        L = lib()
        L.occ_cosx_compute_K_range_separated_mo.restype = C.c_int
        L.occ_cosx_compute_K_range_separated_mo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                                            C.c_double, C.c_double, C.c_double, C.c_void_p]
        check(L.occ_cosx_compute_K_range_separated_mo(self._h, ptr(D), ptr(Cp), Cp.shape[1], alpha, beta, omega,
                                                      precision, ptr(K)))
        return K

    def compute_K_range_separated(self, D, alpha: float, beta: float, omega: float,
                                  precision: float = np.finfo(np.float64).eps, unrestricted: bool = False):
        """compute_K_range_separated (seminumerical_exchange.cpp:830-893): restricted, or unrestricted with the
        stacked (2 nbf x nbf) D; general spinorbitals are refused like the reference"""
        nbf = self.basis.nbf
        D = np.asfortranarray(D, dtype=np.float64)
        rows = 2 * nbf if unrestricted else nbf
        if D.shape != (rows, nbf):
            raise RuntimeError(f"density matrix has shape {D.shape}, expected {(rows, nbf)}")
        K = np.zeros((rows, nbf), order="F")
        L = lib()
        L.occ_cosx_compute_K_range_separated_kind.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                                              C.c_double, C.c_double, C.c_void_p]
        check(L.occ_cosx_compute_K_range_separated_kind(self._h, 1 if unrestricted else 0, ptr(D), alpha, beta, omega,
                                                        precision, ptr(K)))
        return K

    def compute_K_for_density(self, D2q, precision: float = np.finfo(np.float64).eps):
        nbf = self.basis.nbf
        D2q = np.asfortranarray(D2q, dtype=np.float64)
        K = np.zeros((nbf, nbf), order="F")
        check(lib().occ_cosx_K_for_density(self._h, ptr(D2q), precision, self.omega, ptr(K)))
        return K

    def compute_K_for_density_device(self, D2q_dev, K_dev, precision: float = np.finfo(np.float64).eps,
                                     w_full: float = 1.0, w_lr: float = 0.0, omega: float = 0.0):
        """Device-resident compute_K_for_density: D2q_dev / K_dev are torch CUDA float64 tensors (nbf x nbf, column-major
        = the transpose of a contiguous torch tensor) on this handle's device. With rank / world_size sharding K_dev is
        this rank's partial sum; the caller adds the partials (one NCCL all-reduce, scripts/gpu_cosx_scaling.py)."""
        L = lib()
        L.occ_cosx_K_for_density_device.restype = C.c_int
        L.occ_cosx_K_for_density_device.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                                    C.c_double, C.c_void_p]
        check(L.occ_cosx_K_for_density_device(self._h, C.c_void_p(D2q_dev.data_ptr()), precision, w_full, w_lr, omega,
                                              C.c_void_p(K_dev.data_ptr())))
        return K_dev

    def compute_overlap_matrix(self):
        S = np.zeros((self.basis.nbf, self.basis.nbf), order="F")
        check(lib().occ_cosx_numerical_overlap(self._h, ptr(S)))
        return S

    def analytic_overlap(self):
        S = np.zeros((self.basis.nbf, self.basis.nbf), order="F")
        check(lib().occ_cosx_analytic_overlap(self._h, ptr(S)))
        return S

    def set_projector(self, P):
        self.overlap_projector = None if P is None else np.asfortranarray(P, dtype=np.float64)
        check(lib().occ_cosx_set_projector(self._h, ptr(self.overlap_projector) if P is not None else None))

    def num_grid_points(self) -> int:
        return len(self.weights)

    def num_batches(self) -> int:
        return int(self.info()["nbatches"])

    def num_atoms(self) -> int:
        return len(self.atom_offsets) - 1

    def info(self) -> dict:
        t = CosxInfoT()
        check(lib().occ_cosx_get_info(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in CosxInfoT._fields_}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().occ_cosx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGPUSemiNumericalExchange:
    """occ_cosx_multi_*: the same engine on several GPUs driven from THIS process (one shard of grid batches per device,
    inputs over every GPU's own PCIe link, the partial K matrices summed through peer memory).  Same constructor arguments
    as SemiNumericalExchange plus `devices` (list of CUDA ordinals; an ordinal may repeat)."""

    def __init__(self, basis: AOBasis, settings: GridSettings | None = None, atom_grids=None, devices=(0,),
                 margin: float = 1.0, batch_bytes: int = 0, projector: np.ndarray | None | str = "auto"):
        L = lib()
        self.basis = basis
        self.omega = 0.0
        if atom_grids is None:
            mg = MolecularGrid(basis, settings or GridSettings())
            pts, w = mg.points()
            counts = [mg.atom_info(a)[0] for a in range(mg.natoms)]
            mg.close()
            offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        else:
            offs, pts, w = atom_grids
        self.atom_offsets = np.ascontiguousarray(offs, dtype=np.uint64)
        self.points = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.weights = np.ascontiguousarray(w, dtype=np.float64)
        self._ba = BasisArrays(basis.as_arrays())
        opt = CosxOptionsT(margin, 0.0, 0.0, batch_bytes, -1, 0, 0, 1)
        dev = np.array(list(devices), dtype=np.int32)
        self.devices = [int(x) for x in dev]
        P0 = None if isinstance(projector, str) or projector is None else np.asfortranarray(projector, dtype=np.float64)
        self._m = C.c_void_p()
        L.occ_cosx_multi_destroy.restype = None
        L.occ_cosx_multi_destroy.argtypes = [C.c_void_p]
        L.occ_cosx_multi_compute_K.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                               C.c_double, C.c_void_p]
        L.occ_cosx_multi_compute_K_range_separated.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                               C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                                               C.c_void_p]
        L.occ_cosx_multi_numerical_overlap.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_cosx_multi_analytic_overlap.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_cosx_multi_set_projector.argtypes = [C.c_void_p, C.c_void_p]
        L.occ_cosx_multi_get_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        check(L.occ_cosx_multi_create(self._ba.ref, C.c_int(len(self.atom_offsets) - 1), ptr(self.atom_offsets),
                                      ptr(self.points), ptr(self.weights), ptr(P0) if P0 is not None else None,
                                      C.byref(opt), C.c_int(len(dev)), ptr(dev), C.byref(self._m)))
        self.overlap_projector = P0
        if isinstance(projector, str):      # Sn.ldlt().solve(S) over the SUMMED numerical overlap of the shards
            S, Sn = self.analytic_overlap(), self.compute_overlap_matrix()
            self.set_projector(np.linalg.solve(Sn, S))

    def set_range_separated_omega(self, omega: float):
        self.omega = float(omega)

    @staticmethod
    def _pack(nbf, Cocc, unrestricted):
        if Cocc is None:
            return None, 0, 0
        blocks = [np.asarray(c, dtype=np.float64).reshape(nbf, -1) for c in (Cocc if unrestricted else (Cocc,))]
        nocc = [b.shape[1] for b in blocks] + [0]
        Cp = np.asfortranarray(np.concatenate(blocks, axis=1)) if sum(nocc) else np.zeros((nbf, 1), order="F")
        return Cp, nocc[0], nocc[1]

    def compute_K(self, D, Cocc=None, precision: float = np.finfo(np.float64).eps, unrestricted: bool = False,
                  general: bool = False):
        """compute_K(mo) / compute_K_mo over all devices; Cocc as in SemiNumericalExchange.compute_K_mo or None"""
        nbf = self.basis.nbf
        kind = 2 if general else (1 if unrestricted else 0)
        shape = ((nbf, nbf), (2 * nbf, nbf), (2 * nbf, 2 * nbf))[kind]
        D = np.asfortranarray(D, dtype=np.float64)
        if D.shape != shape:
            raise RuntimeError(f"compute_K: density matrix of shape {D.shape}, expected {shape}")
        Cp, na, nb = self._pack(nbf, Cocc, unrestricted)
        K = np.zeros(shape, order="F")
        check(lib().occ_cosx_multi_compute_K(self._m, kind, ptr(D), ptr(Cp) if Cp is not None else None, na, nb,
                                             precision, self.omega, ptr(K)))
        return K

    def compute_K_range_separated(self, D, alpha: float, beta: float, omega: float, Cocc=None,
                                  precision: float = np.finfo(np.float64).eps, unrestricted: bool = False):
        nbf = self.basis.nbf
        rows = 2 * nbf if unrestricted else nbf
        D = np.asfortranarray(D, dtype=np.float64)
        if D.shape != (rows, nbf):
            raise RuntimeError(f"density matrix has shape {D.shape}, expected {(rows, nbf)}")
        Cp, na, nb = self._pack(nbf, Cocc, unrestricted)
        K = np.zeros((rows, nbf), order="F")
        check(lib().occ_cosx_multi_compute_K_range_separated(self._m, 1 if unrestricted else 0, ptr(D),
                                                             ptr(Cp) if Cp is not None else None, na, nb, alpha, beta,
                                                             omega, precision, ptr(K)))
        return K

    def compute_overlap_matrix(self):
        S = np.zeros((self.basis.nbf, self.basis.nbf), order="F")
        check(lib().occ_cosx_multi_numerical_overlap(self._m, ptr(S)))
        return S

    def analytic_overlap(self):
        S = np.zeros((self.basis.nbf, self.basis.nbf), order="F")
        check(lib().occ_cosx_multi_analytic_overlap(self._m, ptr(S)))
        return S

    def set_projector(self, P):
        self.overlap_projector = None if P is None else np.asfortranarray(P, dtype=np.float64)
        check(lib().occ_cosx_multi_set_projector(self._m, ptr(self.overlap_projector) if P is not None else None))

    def info(self, rank: int) -> dict:
        t = CosxInfoT()
        check(lib().occ_cosx_multi_get_info(self._m, C.c_int(rank), C.byref(t)))
        return {f: getattr(t, f) for f, _ in CosxInfoT._fields_}

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            lib().occ_cosx_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
