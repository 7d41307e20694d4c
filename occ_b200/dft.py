"""Host-side mirror of occ::dft::DFT for the exchange-correlation build (reference: include/occ/dft/dft.h:36-763).

Same names and argument meaning as the reference's interface for this path:
  DFT(method, basis, grid, spinorbital_kind)   ~ DFT::DFT (src/dft/dft.cpp:19-35) + set_integration_grid + set_method
  DFT.compute_K_dft(D)                         ~ DFT::compute_K_dft<deriv, kind>(mo, Schwarz) (dft.h:190-388); takes mo.D
  DFT.compute_vxc(D)                           ~ DFT::compute_vxc (dft.h:395-429)
  DFT.exchange_correlation_energy()            ~ m_exc_dft
  DFT.density_derivative(), exact_exchange_factor(), set_density_threshold ... as in the reference
All arithmetic runs in the CUDA library behind the C ABI (include/occ_xc_b200.h); PyTorch is used only for device
buffers / streams in the device-resident entry (build_device) and torch.distributed for the multi-GPU sum.
"""
from __future__ import annotations

import ctypes as C
import ctypes as C_
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import LEAF_DTYPE, BasisArrays, InfoT, OptionsT, check, lib, ptr
from .aobasis import AOBasis

FUNCTIONAL_IDS = {"lda_x": 1, "lda_c_vwn": 7, "lda_c_vwn_rpa": 8, "lda_c_vwn_3": 30, "gga_x_pbe": 101,
                  "gga_x_b88": 106, "gga_c_pbe": 130, "gga_c_lyp": 131, "hyb_gga_xc_b3lyp": 402,
                  "mgga_x_r2scan": 497, "mgga_c_r2scan": 498, "lda_c_pw": 12, "lda_c_pw_mod": 13, "mgga_x_tpss": 202,
                  "mgga_c_tpss": 231, "hyb_gga_xc_pbeh": 406, "hyb_gga_xc_bhandh": 435, "hyb_gga_xc_bhandhlyp": 436,
                  "hyb_gga_xc_b3lyp5": 475, "hyb_gga_xc_wb97x": 464, "hyb_gga_xc_b97": 407,
                  "hyb_gga_xc_b3pw91": 401, "gga_c_pw91": 134, "gga_c_p86": 132, "lda_c_pz": 9,
                  "hyb_gga_xc_wb97x_v": 466, "hyb_gga_xc_b3p86": 403, "gga_x_b86_mgc": 105}
# DensityFunctional::needs_nlc_correction (src/dft/functional.cpp:66-74): hyb_mgga_xc_wb97m_v (not implemented), hyb_gga_xc_wb97x_v
NLC_FUNCTIONAL_IDS = (466,)

# src/dft/dft_method.cpp:14-37 (builtin_definitions), restricted to components this library implements
METHOD_DEFINITIONS = {
    "b3lyp": ["hyb_gga_xc_b3lyp"],
    "pbe": ["gga_x_pbe", "gga_c_pbe"],
    "pbepbe": ["gga_x_pbe", "gga_c_pbe"],
    "svwn": ["lda_x", "lda_c_vwn_3"],
    "lda": ["lda_x", "lda_c_vwn_3"],
    "lsda": ["lda_x", "lda_c_vwn_3"],
    "svwn5": ["lda_x", "lda_c_vwn"],
    "blyp": ["gga_x_b88", "gga_c_lyp"],
    "bpbe": ["gga_x_b88", "gga_c_pbe"],
    "r2scan": ["mgga_x_r2scan", "mgga_c_r2scan"],
    "tpss": ["mgga_x_tpss", "mgga_c_tpss"],
    "b3lyp5": ["hyb_gga_xc_b3lyp5"],
    "pbeh": ["hyb_gga_xc_pbeh"],
    "wb97x": ["hyb_gga_xc_wb97x"],
    "b97": ["hyb_gga_xc_b97"],
    "b3pw91": ["hyb_gga_xc_b3pw91"],
    "bp86": ["gga_x_b88", "gga_c_p86"],
    # share/methods/dft_methods.json entries beyond the compiled-in table
    "wb97x-v": ["hyb_gga_xc_wb97x_v"],        # "dispersion": "vv10" -> DFT.post_scf_nlc_correction
    "b3p86": ["hyb_gga_xc_b3p86"],
    "b86bpbe": ["gga_x_b86_mgc", "gga_c_pbe"],
    "b86bpbeh": ["gga_x_b86_mgc", "gga_c_pbe"],   # factor 0.75 / hfx 0.25 carried, never applied (as pbe0 below)
    # the 0.75 component factor of pbe0 / pbe1pbe is stored but never applied by the reference
    # (include/occ/dft/functional.h:506-508), so its XC build evaluates full PBE exchange; same here
    "pbe0": ["gga_x_pbe", "gga_c_pbe"],
    "pbe1pbe": ["gga_x_pbe", "gga_c_pbe"],
}
# builtin entries whose libxc components are NOT implemented here (get_dft_method raises for them): m062x, wb97m(-v)

RESTRICTED, UNRESTRICTED = "restricted", "unrestricted"


def method_definitions() -> dict:
    """get_method_definitions (src/dft/dft_method.cpp:96-132): `$OCC_DATA_PATH/methods/dft_methods.json`, else
    `./dft_methods.json`, REPLACES the builtin table when it exists and holds a "methods" object; otherwise the
    builtin definitions.  Each entry lists components {"id": libxc name, "factor", "hfx"}; like the reference's XC build
    the component factor is carried but never applied (include/occ/dft/functional.h:506-508)."""
    import json
    import os
    base = os.environ.get("OCC_DATA_PATH", ".")
    for path in (os.path.join(base, "methods", "dft_methods.json"), "dft_methods.json"):
        if os.path.exists(path):
            try:
                with open(path) as f:
                    j = json.load(f)
                if isinstance(j.get("methods"), dict):
                    return {name: [c["id"] for c in d.get("components", [])] for name, d in j["methods"].items()}
            except Exception:
                pass
            break
    return METHOD_DEFINITIONS


def get_dft_method(method_string: str) -> list[int]:
    """get_dft_method (src/dft/dft_method.cpp:161-227): space-separated tokens, each a method name or a libxc
    functional name; a leading 'u' marks the polarized variant (the spin kind is chosen by the caller here)."""
    ids = []
    defs = method_definitions()
    for token in method_string.strip().lower().split():
        m = token[1:] if token.startswith("u") else token      # dft_method.cpp:175-178 strips a leading 'u' unconditionally
        if m in defs:
            missing = [c for c in defs[m] if c not in FUNCTIONAL_IDS]
            if missing:
                raise RuntimeError(f"functional {missing[0]} not found")   # functional.cpp:32 analogue
            ids += [FUNCTIONAL_IDS[c] for c in defs[m]]
        elif m in FUNCTIONAL_IDS:
            ids.append(FUNCTIONAL_IDS[m])
        else:
            raise RuntimeError(f"functional {m} not found")  # functional.cpp:32 analogue
    return ids


@dataclass
class GridHierarchy:
    """SpatialGridHierarchy (include/occ/numint/spatial_grid_hierarchy.h:33-78): Morton-sorted points/weights + leaves."""
    sorted_points: np.ndarray    # (npts, 3) C-order == Mat3N column-major
    sorted_weights: np.ndarray   # (npts,)
    leaves: np.ndarray           # LEAF_DTYPE

    @property
    def num_leaves(self):
        return len(self.leaves)


class _NonLocalCorrelationMixin:
    """VV10 hooks of occ::dft::DFT (include/occ/dft/dft.h:44,467-495): post-SCF energy and gradient of the NLC part of
    wB97X-V.  The host class provides func_ids, nspin, basis and _nlc_devices() (None: one handle on its own device)."""
    _nlc = None
    nlc_energy = 0.0

    def have_nonlocal_correlation(self) -> bool:
        return any(int(i) in NLC_FUNCTIONAL_IDS for i in self.func_ids)

    def set_nlc_grid(self, settings=None, gradient: bool = True, devices=None):
        """DFT::set_nlc_grid (src/dft/dft.cpp:119-121); default as NonLocalCorrelationFunctional.set_integration_grid (110 / 50 angular, 1e-7, H / He reduced)"""
        from .nlc import NonLocalCorrelationFunctional
        if self.nspin != 1:
            raise RuntimeError("Only restricted NLC functional implemented right now")   # nonlocal_correlation.cpp:370
        if self._nlc is not None:
            self._nlc.close()
        self._nlc = NonLocalCorrelationFunctional()
        devices = devices if devices is not None else self._nlc_devices()
        self._nlc.set_integration_grid(self.basis, settings, device=getattr(self, "_device", -1), gradient=gradient,
                                       devices=devices)
        return self._nlc

    def post_scf_nlc_correction(self, D: np.ndarray) -> float:
        """DFT::post_scf_nlc_correction(mo): the VV10 energy of D = mo.D, 0 for a functional without an NLC part"""
        if self.have_nonlocal_correlation():
            nlc = self._nlc or self.set_nlc_grid()
            self.nlc_energy = nlc(D)[0]
        return self.nlc_energy

    def compute_nlc_gradient(self, D: np.ndarray) -> np.ndarray:
        """DFT::compute_nlc_gradient(mo): zeros without an NLC part, else NonLocalCorrelationFunctional::compute_gradient"""
        if not self.have_nonlocal_correlation():
            return np.zeros((len(self.basis.atomic_numbers), 3))
        nlc = self._nlc or self.set_nlc_grid()
        return nlc.compute_gradient(D)

    def _close_nlc(self):
        if self._nlc is not None:
            self._nlc.close()
            self._nlc = None


class DFT(_NonLocalCorrelationMixin):
    def __init__(self, method: str, basis: AOBasis, grid: GridHierarchy, spinorbital_kind: str = RESTRICTED,
                 device: int = -1, rank: int = 0, world_size: int = 1, batch_bytes: int = 0,
                 density_threshold: float = 1e-10, xc_screening_threshold: float = 1e-10, shell_extents=None,
                 debug: int = 0, lanes: int = 1, gradient: bool = False, nocc=None, deterministic: bool = False,
                 nvtx: bool = False, density_scale: float = 0.0, orbital_scale: float = 0.0):
        self.basis = basis
        self.method = method
        self._device = device
        self.func_ids = np.array(get_dft_method(method), dtype=np.int32)
        if spinorbital_kind not in (RESTRICTED, UNRESTRICTED):
            raise RuntimeError("Not implemented: DFT for General spinorbitals")      # dft.h:426-428
        self.nspin = 2 if spinorbital_kind == UNRESTRICTED else 1
        self.nbf = basis.nbf
        self._basis_arrays = BasisArrays(basis.as_arrays())
        pts = np.ascontiguousarray(grid.sorted_points, dtype=np.float64).reshape(-1, 3)
        w = np.ascontiguousarray(grid.sorted_weights, dtype=np.float64)
        leaves = np.ascontiguousarray(grid.leaves, dtype=LEAF_DTYPE)
        ext = None if shell_extents is None else np.ascontiguousarray(shell_extents, dtype=np.float64)
        opt = OptionsT()
        lib().occ_xc_options_init(C.byref(opt))
        opt.device, opt.rank, opt.world_size, opt.batch_bytes = device, rank, world_size, batch_bytes
        opt.debug, opt.lanes, opt.gradient = debug, lanes, int(gradient)
        opt.deterministic, opt.nvtx = int(deterministic), int(nvtx)
        opt.density_scale, opt.orbital_scale = density_scale, orbital_scale
        self.nocc = None
        if nocc is not None:        # enables build_mo: occupied-orbital counts (alpha[, beta]); beta may be 0
            na, nb_ = (nocc, nocc) if np.isscalar(nocc) else (nocc[0], nocc[-1])
            opt.mo_form, opt.nocc_alpha, opt.nocc_beta = 1, int(na), int(nb_)
            self.nocc = (int(na), int(nb_) if self.nspin == 2 else 0)
        self._h = C.c_void_p()
        check(lib().occ_xc_create(self._basis_arrays.ref, C.c_size_t(len(w)), ptr(pts), ptr(w), C.c_size_t(len(leaves)),
                                  ptr(leaves), ptr(ext) if ext is not None else None, C.c_int(len(self.func_ids)),
                                  ptr(self.func_ids), C.c_int(self.nspin), C.c_double(density_threshold),
                                  C.c_double(xc_screening_threshold), C.byref(opt), C.byref(self._h)))
        self.rank, self.world_size = rank, world_size
        self.m_exc_dft = 0.0
        self.total_density = (0.0, 0.0)

    # ---- reference-named accessors ----
    def density_derivative(self) -> int:
        return max(lib().occ_xc_family(int(f)) for f in self.func_ids)

    def exact_exchange_factor(self) -> float:
        """DFTMethod::exchange_factor (src/dft/dft_method.cpp:39-47): the first component with a positive factor.  The "hfx"
        of a method definition (pbe0: 0.25) is set on the POLARIZED functional objects only (:150-152) while this walk
        uses the unpolarized list, so it never shows up here either — pbe0 reports 0 like in the reference."""
        for f in self.func_ids:
            x = float(lib().occ_xc_exact_exchange(int(f)))
            if x > 0.0:
                return x
        return 0.0

    def range_separated_parameters(self) -> dict:
        """DFT::range_separated_parameters (src/dft/dft.cpp:128-130): the component with omega != 0, if any"""
        out = {"omega": 0.0, "alpha": 0.0, "beta": 0.0}
        for f in self.func_ids:
            o, a, b = C.c_double(), C.c_double(), C.c_double()
            check(lib().occ_xc_range_separation(int(f), C.byref(o), C.byref(a), C.byref(b)))
            if o.value != 0.0:
                out = {"omega": o.value, "alpha": a.value, "beta": b.value}
        return out

    def exchange_correlation_energy(self) -> float:
        return self.m_exc_dft

    def info(self) -> dict:
        i = InfoT()
        check(lib().occ_xc_get_info(self._h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in InfoT._fields_}

    # ---- the hot path ----
    def compute_K_dft(self, D: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """D = mo.D, (nspin*nbf, nbf); returns K like the reference (the caller negates: jk.K = -compute_vxc)."""
        D2 = np.asfortranarray(2.0 * np.asarray(D, dtype=np.float64))  # dft.h:205
        return self.build(D2, out)

    compute_vxc = compute_K_dft

    def build(self, D2: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """occ_xc_build: D2 already doubled, column-major host buffers."""
        n = self.nspin * self.nbf
        D2 = np.asfortranarray(D2, dtype=np.float64)
        if D2.shape != (n, self.nbf):
            raise RuntimeError(f"density matrix has shape {D2.shape}, expected {(n, self.nbf)}")
        K = out if out is not None else np.zeros((n, self.nbf), order="F")
        exc, na, nb = C.c_double(), C.c_double(), C.c_double()
        check(lib().occ_xc_build(self._h, ptr(D2), ptr(K), C.byref(exc), C.byref(na), C.byref(nb)))
        self.m_exc_dft = exc.value
        self.total_density = (na.value, nb.value)
        return K

    def build_mo(self, D2: np.ndarray, C, out: np.ndarray | None = None) -> np.ndarray:
        """occ_xc_build_mo: D2 as in build(); C = list of nspin matrices (nbf x nocc_s) with D2_s = C_s C_s^T
        (restricted: sqrt(2) * mo.Cocc).  Same results as build() up to the rounding of C C^T."""
        if self.nocc is None:
            raise RuntimeError("DFT was created without nocc: build_mo is not available")
        n = self.nspin * self.nbf
        if D2 is not None:            # None: D2 = C C^T is formed on the device
            D2 = np.asfortranarray(D2, dtype=np.float64)
            if D2.shape != (n, self.nbf):
                raise RuntimeError(f"density matrix has shape {D2.shape}, expected {(n, self.nbf)}")
        Cs = [C] if isinstance(C, np.ndarray) else list(C)
        if len(Cs) != self.nspin or any(c.shape != (self.nbf, self.nocc[s]) for s, c in enumerate(Cs)):
            raise RuntimeError("orbital coefficients do not match (nbf, nocc) of the handle")
        packed = np.concatenate([np.asfortranarray(c, dtype=np.float64).ravel(order="F") for c in Cs])
        K = out if out is not None else np.zeros((n, self.nbf), order="F")
        exc, na, nb = C_.c_double(), C_.c_double(), C_.c_double()
        check(lib().occ_xc_build_mo(self._h, ptr(D2) if D2 is not None else None, ptr(packed), ptr(K), C_.byref(exc),
                                    C_.byref(na), C_.byref(nb)))
        self.m_exc_dft = exc.value
        self.total_density = (na.value, nb.value)
        return K

    def build_mo_device(self, d2_dev_ptr: int, c_dev_ptr: int, k_dev_ptr: int, scalars_dev_ptr: int, stream: int = 0):
        check(lib().occ_xc_build_mo_device(self._h, C_.c_void_p(d2_dev_ptr or None), C_.c_void_p(c_dev_ptr),
                                           C_.c_void_p(k_dev_ptr), C_.c_void_p(scalars_dev_ptr), C_.c_void_p(stream)))

    def build_mo_raw(self, d2_ptr: int, c_ptr: int, k_ptr: int):
        """occ_xc_build_mo on raw host pointers (d2_ptr = 0: D2 formed on the device); returns (exc, n_alpha, n_beta)"""
        exc, na, nb = C_.c_double(), C_.c_double(), C_.c_double()
        check(lib().occ_xc_build_mo(self._h, C_.c_void_p(d2_ptr or None), C_.c_void_p(c_ptr), C_.c_void_p(k_ptr),
                                    C_.byref(exc), C_.byref(na), C_.byref(nb)))
        return exc.value, na.value, nb.value

    def build_raw(self, d2_ptr: int, k_ptr: int):
        """occ_xc_build on raw host pointers (pinned buffers owned by the caller); returns (exc, n_alpha, n_beta)."""
        exc, na, nb = C.c_double(), C.c_double(), C.c_double()
        check(lib().occ_xc_build(self._h, C.c_void_p(d2_ptr), C.c_void_p(k_ptr), C.byref(exc), C.byref(na),
                                 C.byref(nb)))
        return exc.value, na.value, nb.value

    def build_device(self, d2_dev_ptr: int, k_dev_ptr: int, scalars_dev_ptr: int, stream: int = 0):
        """occ_xc_build_device: everything resident in HBM, asynchronous on `stream` (a cudaStream_t value)."""
        check(lib().occ_xc_build_device(self._h, C.c_void_p(d2_dev_ptr), C.c_void_p(k_dev_ptr),
                                        C.c_void_p(scalars_dev_ptr), C.c_void_p(stream)))

    def compute_xc_gradient(self, D: np.ndarray) -> np.ndarray:
        """DFT::compute_xc_gradient(mo) (src/dft/dft_gradients.cpp:68-109); D = mo.D.  For a functional with a VV10 part
        (wB97X-V) the NLC gradient is added like the reference does (:104).
        Returns (natoms, 3) — the bytes of the reference's 3 x natoms column-major Mat3N."""
        g = self.gradient(np.asfortranarray(2.0 * np.asarray(D, dtype=np.float64)))
        if self.have_nonlocal_correlation():
            g = g + self.compute_nlc_gradient(D)
        return g

    def _nlc_devices(self):
        return None

    def gradient(self, D2: np.ndarray) -> np.ndarray:
        n = self.nspin * self.nbf
        D2 = np.asfortranarray(D2, dtype=np.float64)
        if D2.shape != (n, self.nbf):
            raise RuntimeError(f"density matrix has shape {D2.shape}, expected {(n, self.nbf)}")
        g = np.zeros((len(self.basis.atomic_numbers), 3))
        check(lib().occ_xc_gradient(self._h, ptr(D2), ptr(g)))
        return g

    def gradient_mo(self, D2: np.ndarray, C) -> np.ndarray:
        """occ_xc_gradient_mo: the gradient with the factorised density / A.D^T products (C as in build_mo)"""
        if self.nocc is None:
            raise RuntimeError("DFT was created without nocc: gradient_mo is not available")
        D2 = np.asfortranarray(D2, dtype=np.float64)
        Cs = [C] if isinstance(C, np.ndarray) else list(C)
        packed = np.concatenate([np.asfortranarray(c, dtype=np.float64).ravel(order="F") for c in Cs])
        g = np.zeros((len(self.basis.atomic_numbers), 3))
        check(lib().occ_xc_gradient_mo(self._h, ptr(D2), ptr(packed), ptr(g)))
        return g

    def set_profiling(self, on: bool):
        check(lib().occ_xc_set_profiling(self._h, C.c_int(int(on))))

    def phase_ms(self) -> dict:
        ms = (C.c_double * _lib.NPHASES)()
        check(lib().occ_xc_get_phase_ms(self._h, ms))
        return dict(zip(_lib.PHASE_NAMES, list(ms)))

    def leaf_collocation(self, leaf: int, count: int, deriv: int = 1):
        cap = count * self.nbf
        outs = [np.zeros(cap) for _ in range(4 if deriv else 1)]
        nb = C.c_int()
        check(lib().occ_xc_leaf_collocation(self._h, C.c_size_t(leaf), C.c_int(deriv), C.byref(nb), ptr(outs[0]),
                                            ptr(outs[1]) if deriv else None, ptr(outs[2]) if deriv else None,
                                            ptr(outs[3]) if deriv else None, C.c_size_t(cap)))
        return [o[:count * nb.value].reshape(nb.value, count).T for o in outs]

    def close(self):
        self._close_nlc()
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().occ_xc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGPUDFT(_NonLocalCorrelationMixin):
    """occ_xc_multi_*: the same build on several GPUs driven from THIS process (what the reference's single-process
    SCF needs, include/occ/qm/scf_impl.h:539): one leaf shard per device, inputs sliced over the PCIe links and
    gathered over NVLink, K reduce-scattered through peer memory.  Same constructor arguments as DFT plus `devices`
    (list of CUDA ordinals; an ordinal may repeat)."""

    def __init__(self, method: str, basis: AOBasis, grid: GridHierarchy, spinorbital_kind: str = RESTRICTED,
                 devices=(0,), batch_bytes: int = 0, density_threshold: float = 1e-10,
                 xc_screening_threshold: float = 1e-10, nocc=None, deterministic: bool = False,
                 density_scale: float = 0.0, orbital_scale: float = 0.0, gradient: bool = False):
        self.basis, self.method = basis, method
        self.func_ids = np.array(get_dft_method(method), dtype=np.int32)
        self.nspin = 2 if spinorbital_kind == UNRESTRICTED else 1
        self.nbf = basis.nbf
        self._basis_arrays = BasisArrays(basis.as_arrays())
        pts = np.ascontiguousarray(grid.sorted_points, dtype=np.float64).reshape(-1, 3)
        w = np.ascontiguousarray(grid.sorted_weights, dtype=np.float64)
        leaves = np.ascontiguousarray(grid.leaves, dtype=LEAF_DTYPE)
        opt = OptionsT()
        lib().occ_xc_options_init(C.byref(opt))
        opt.batch_bytes, opt.deterministic = batch_bytes, int(deterministic)
        opt.density_scale, opt.orbital_scale = density_scale, orbital_scale
        opt.gradient = int(gradient)
        self.nocc = None
        if nocc is not None:
            na, nb_ = (nocc, nocc) if np.isscalar(nocc) else (nocc[0], nocc[-1])
            opt.mo_form, opt.nocc_alpha, opt.nocc_beta = 1, int(na), int(nb_)
            self.nocc = (int(na), int(nb_) if self.nspin == 2 else 0)
        dev = np.array(list(devices), dtype=np.int32)
        self.devices = [int(x) for x in dev]
        self._m = C.c_void_p()
        check(lib().occ_xc_multi_create(self._basis_arrays.ref, C.c_size_t(len(w)), ptr(pts), ptr(w),
                                        C.c_size_t(len(leaves)), ptr(leaves), None, C.c_int(len(self.func_ids)),
                                        ptr(self.func_ids), C.c_int(self.nspin), C.c_double(density_threshold),
                                        C.c_double(xc_screening_threshold), C.byref(opt), C.c_int(len(dev)), ptr(dev),
                                        C.byref(self._m)))
        self.m_exc_dft = 0.0
        self.total_density = (0.0, 0.0)

    def info(self, rank: int) -> dict:
        i = InfoT()
        check(lib().occ_xc_multi_get_info(self._m, C.c_int(rank), C.byref(i)))
        return {f: getattr(i, f) for f, _ in InfoT._fields_}

    def exchange_correlation_energy(self) -> float:
        return self.m_exc_dft

    def build(self, D2, Cs=None, out: np.ndarray | None = None) -> np.ndarray:
        """occ_xc_multi_build: D2 (or None when Cs is given) and optionally the orbitals as in DFT.build_mo"""
        n = self.nspin * self.nbf
        if D2 is not None:
            D2 = np.asfortranarray(D2, dtype=np.float64)
            if D2.shape != (n, self.nbf):
                raise RuntimeError(f"density matrix has shape {D2.shape}, expected {(n, self.nbf)}")
        packed = None
        if Cs is not None:
            Cs = [Cs] if isinstance(Cs, np.ndarray) else list(Cs)
            if self.nocc is None or any(c.shape != (self.nbf, self.nocc[s]) for s, c in enumerate(Cs)):
                raise RuntimeError("orbital coefficients do not match (nbf, nocc) of the handles")
            packed = np.concatenate([np.asfortranarray(c, dtype=np.float64).ravel(order="F") for c in Cs])
        K = out if out is not None else np.zeros((n, self.nbf), order="F")
        exc, na, nb = C.c_double(), C.c_double(), C.c_double()
        check(lib().occ_xc_multi_build(self._m, ptr(D2) if D2 is not None else None,
                                       ptr(packed) if packed is not None else None, ptr(K), C.byref(exc), C.byref(na),
                                       C.byref(nb)))
        self.m_exc_dft = exc.value
        self.total_density = (na.value, nb.value)
        return K

    def gradient(self, D2: np.ndarray, Cs=None) -> np.ndarray:
        """occ_xc_multi_gradient: the XC nuclear gradient (natoms, 3) summed over the shards; needs gradient=True"""
        n = self.nspin * self.nbf
        D2 = np.asfortranarray(D2, dtype=np.float64)
        if D2.shape != (n, self.nbf):
            raise RuntimeError(f"density matrix has shape {D2.shape}, expected {(n, self.nbf)}")
        packed = None
        if Cs is not None:
            Cs = [Cs] if isinstance(Cs, np.ndarray) else list(Cs)
            packed = np.concatenate([np.asfortranarray(c, dtype=np.float64).ravel(order="F") for c in Cs])
        g = np.zeros((len(self.basis.atomic_numbers), 3))
        check(lib().occ_xc_multi_gradient(self._m, ptr(D2), ptr(packed) if packed is not None else None, ptr(g)))
        return g

    def compute_xc_gradient(self, D: np.ndarray) -> np.ndarray:
        """DFT::compute_xc_gradient(mo) on the shards (src/dft/dft_gradients.cpp:68-109), D = mo.D; the VV10 gradient of a
        functional with an NLC part is added (:104), computed on the same devices"""
        g = self.gradient(np.asfortranarray(2.0 * np.asarray(D, dtype=np.float64)))
        if self.have_nonlocal_correlation():
            g = g + self.compute_nlc_gradient(D)
        return g

    def _nlc_devices(self):
        return self.devices

    def build_raw(self, d2_ptr: int, c_ptr: int, k_ptr: int):
        """occ_xc_multi_build on raw host pointers (pinned buffers owned by the caller; 0 = NULL)"""
        exc, na, nb = C.c_double(), C.c_double(), C.c_double()
        check(lib().occ_xc_multi_build(self._m, C.c_void_p(d2_ptr or None), C.c_void_p(c_ptr or None), C.c_void_p(k_ptr),
                                       C.byref(exc), C.byref(na), C.byref(nb)))
        return exc.value, na.value, nb.value

    def close(self):
        self._close_nlc()
        if getattr(self, "_m", None) is not None and self._m.value:
            lib().occ_xc_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decay_cutoff(basis: AOBasis, threshold: float = 1e-10) -> np.ndarray:
    """gto::evaluate_decay_cutoff (src/gto/gto.cpp:18-67)"""
    ba = BasisArrays(basis.as_arrays())
    ext = np.zeros(basis.nsh)
    check(lib().occ_xc_decay_cutoff(ba.ref, C.c_double(threshold), ptr(ext)))
    return ext


def eval_functional(func, polarized: bool, rho, sigma=None, tau=None) -> dict:
    """DensityFunctional::evaluate (src/dft/functional.cpp:18-65) on libxc-layout arrays, on the GPU."""
    fid = FUNCTIONAL_IDS[func] if isinstance(func, str) else int(func)
    s = 2 if polarized else 1
    rho = np.ascontiguousarray(rho, dtype=np.float64).ravel()
    n = rho.size // s
    ns = 3 if polarized else 1
    sigma = np.ascontiguousarray(sigma if sigma is not None else np.zeros(n * ns), dtype=np.float64).ravel()
    tau = np.ascontiguousarray(tau if tau is not None else np.zeros(n * s), dtype=np.float64).ravel()
    exc, vrho, vsigma, vtau = np.zeros(n), np.zeros(n * s), np.zeros(n * ns), np.zeros(n * s)
    check(lib().occ_xc_eval_functional(C.c_int(fid), C.c_int(int(polarized)), C.c_size_t(n), ptr(rho), ptr(sigma),
                                       ptr(tau), ptr(exc), ptr(vrho), ptr(vsigma), ptr(vtau)))
    return {"exc": exc, "vrho": vrho.reshape(n, s), "vsigma": vsigma.reshape(n, ns), "vtau": vtau.reshape(n, s)}


def partition_leaves(basis: AOBasis, leaves: np.ndarray, world_size: int, screening_threshold: float = 1e-10,
                     shell_extents=None):
    """Leaf -> rank map of the multi-GPU build (occ_xc_partition_leaves; host only). Returns (leaf_rank, leaf_nb)."""
    ba = BasisArrays(basis.as_arrays())
    leaves = np.ascontiguousarray(leaves, dtype=LEAF_DTYPE)
    rank = np.zeros(len(leaves), dtype=np.int32)
    nb = np.zeros(len(leaves), dtype=np.int32)
    ext = None if shell_extents is None else np.ascontiguousarray(shell_extents, dtype=np.float64)
    check(lib().occ_xc_partition_leaves(ba.ref, C.c_size_t(len(leaves)), ptr(leaves),
                                        ptr(ext) if ext is not None else None, C.c_double(screening_threshold),
                                        C.c_int(world_size), ptr(rank), ptr(nb)))
    return rank, nb


def evaluate_density_on_grid(basis: AOBasis, D: np.ndarray, grid_pts: np.ndarray, max_derivative: int = 0,
                             spinorbital_kind: str = RESTRICTED, device: int = -1,
                             screening_threshold: float = 0.0) -> np.ndarray:
    """occ::density::evaluate_density_on_grid<max_derivative, kind>(basis, D, grid_pts) (include/occ/gto/density.h:67-107).
    D is the matrix evaluate_density receives (callers pass 2 * mo.D, density.h:84-87); grid_pts (npts, 3) in Bohr.
    Returns (nspin * npts, {1, 4, 6}) in the CALLER's point order: rho, grad rho, laplacian, tau.
    screening_threshold 0 = every shell at every point, like the reference's evaluate_basis."""
    from .grid import MolecularGrid          # local import: grid.py imports this module
    nspin = 2 if spinorbital_kind == UNRESTRICTED else 1
    method = {0: "lda_x", 1: "gga_x_b88", 2: "mgga_x_tpss"}[int(max_derivative)]
    pts = np.ascontiguousarray(grid_pts, dtype=np.float64).reshape(-1, 3)
    mg = MolecularGrid.from_points(pts, np.ones(len(pts)))
    h, perm = mg.hierarchy(with_permutation=True)
    mg.close()
    d = DFT(method, basis, h, spinorbital_kind, device=device, xc_screening_threshold=screening_threshold)
    try:
        ncol = {0: 1, 1: 4, 2: 6}[int(max_derivative)]
        n = len(pts)
        D2 = np.asfortranarray(D, dtype=np.float64)
        if D2.shape != (nspin * basis.nbf, basis.nbf):
            raise RuntimeError(f"density matrix has shape {D2.shape}, expected {(nspin * basis.nbf, basis.nbf)}")
        out = np.zeros((nspin * n, ncol), order="F")
        check(lib().occ_xc_density_on_grid(d._h, ptr(D2), ptr(out)))
    finally:
        d.close()
    res = np.zeros_like(out)
    idx = perm.astype(np.int64)
    for s in range(nspin):
        res[s * n + idx] = out[s * n:(s + 1) * n]
    return res
