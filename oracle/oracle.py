"""ctypes front end of the CPU ORACLE (oracle/src/*.cpp -> oracle/liboccxc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under occ_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboccxc_oracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libocc_ref.so")

FUNC_IDS = {"lda_x": 1, "lda_c_vwn": 7, "lda_c_vwn_rpa": 8, "lda_c_vwn_3": 30, "gga_x_pbe": 101, "gga_x_b88": 106,
            "gga_c_pbe": 130, "gga_c_lyp": 131, "hyb_gga_xc_b3lyp": 402, "mgga_x_r2scan": 497, "mgga_c_r2scan": 498,
            "lda_c_pw": 12, "lda_c_pw_mod": 13, "mgga_x_tpss": 202, "mgga_c_tpss": 231, "hyb_gga_xc_pbeh": 406,
            "hyb_gga_xc_bhandh": 435, "hyb_gga_xc_bhandhlyp": 436, "hyb_gga_xc_b3lyp5": 475, "hyb_gga_xc_wb97x": 464, "hyb_gga_xc_b97": 407,
            "hyb_gga_xc_b3pw91": 401, "gga_c_pw91": 134, "gga_c_p86": 132, "lda_c_pz": 9,
            "hyb_gga_xc_wb97x_v": 466, "hyb_gga_xc_b3p86": 403, "gga_x_b86_mgc": 105}

# src/dft/dft_method.cpp:14-37 (same entries in share/methods/dft_methods.json)
METHODS = {
    "b3lyp": ["hyb_gga_xc_b3lyp"],
    "pbe": ["gga_x_pbe", "gga_c_pbe"],
    "blyp": ["gga_x_b88", "gga_c_lyp"],
    "lda": ["lda_x", "lda_c_vwn_3"],
    "svwn": ["lda_x", "lda_c_vwn_3"],
    "svwn5": ["lda_x", "lda_c_vwn"],
    "r2scan": ["mgga_x_r2scan", "mgga_c_r2scan"],
    "tpss": ["mgga_x_tpss", "mgga_c_tpss"],
    "b3lyp5": ["hyb_gga_xc_b3lyp5"],
    "wb97x": ["hyb_gga_xc_wb97x"],
    "b97": ["hyb_gga_xc_b97"],
    "b3pw91": ["hyb_gga_xc_b3pw91"],
    "bp86": ["gga_x_b88", "gga_c_p86"],
    "wb97x-v": ["hyb_gga_xc_wb97x_v"],
    "b3p86": ["hyb_gga_xc_b3p86"],
    "b86bpbe": ["gga_x_b86_mgc", "gga_c_pbe"],
    "pbeh": ["hyb_gga_xc_pbeh"],
    "bpbe": ["gga_x_b88", "gga_c_pbe"],
    # pbe0 / pbe1pbe carry a 0.75 component factor that the reference stores but never applies
    # (include/occ/dft/functional.h:506-508): they evaluate full PBE exchange
    "pbe0": ["gga_x_pbe", "gga_c_pbe"],
    "pbe1pbe": ["gga_x_pbe", "gga_c_pbe"],
}


class BasisT(C.Structure):
    _fields_ = [("nsh", C.c_int), ("pure", C.c_int), ("l", C.c_void_p), ("origin", C.c_void_p), ("nprim", C.c_void_p),
                ("prim_off", C.c_void_p), ("exps", C.c_void_p), ("coefs", C.c_void_p), ("shell_atom", C.c_void_p)]


class GridSettingsT(C.Structure):
    _fields_ = [("max_angular_points", C.c_int), ("min_angular_points", C.c_int), ("radial_precision", C.c_double),
                ("reduced_first_row_element_grid", C.c_int), ("treutler_alrichs_adjustment", C.c_int),
                ("pruning_scheme", C.c_int), ("int_acc", C.c_double), ("angular_regions", C.c_int * 5)]


class LeafT(C.Structure):
    _fields_ = [("offset", C.c_size_t), ("count", C.c_size_t), ("center", C.c_double * 3), ("radius", C.c_double)]


class StatsT(C.Structure):
    _fields_ = [("sum_n_nb2", C.c_double), ("sum_nb2", C.c_double), ("nleaves", C.c_size_t), ("nskipped", C.c_size_t),
                ("nb_min", C.c_int), ("nb_max", C.c_int), ("t_colloc", C.c_double), ("t_density", C.c_double),
                ("t_func", C.c_double), ("t_vxc", C.c_double), ("t_total", C.c_double), ("t_leaf_loop", C.c_double),
                ("t_fixed", C.c_double)]


LEAF_DTYPE = np.dtype([("offset", np.uint64), ("count", np.uint64), ("center", np.float64, 3), ("radius", np.float64)])

_lib = None


def build(force: bool = False):
    """Compile the oracle (and oracle/_ref when the reference tree is present). Building the checker is not using it."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in glob.glob(os.path.join(HERE, "src", "*"))):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    if os.path.isdir("/root/reference/src/3rdparty/gau2grid") and not os.path.exists(REF_LIB_PATH):
        subprocess.check_call(["make", "-C", HERE, "-s", "-j8", "ref"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_common_fac.restype = C.c_double
        L.orc_xc_exx.restype = C.c_double
        L.orc_molecular_grid_create.restype = C.c_void_p
        L.orc_grid_npts.restype = C.c_size_t
        L.orc_grid_atom_npts.restype = C.c_size_t
        L.orc_hierarchy.restype = C.c_size_t
        L.orc_grid_npts.argtypes = [C.c_void_p]
        L.orc_grid_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_grid_atom_npts.argtypes = [C.c_void_p, C.c_int]
        L.orc_grid_atom_nradial.argtypes = [C.c_void_p, C.c_int]
        L.orc_grid_destroy.argtypes = [C.c_void_p]
        _lib = L
        tables = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "lebedev_tables.npz"))
        for k in tables.files:
            a = np.ascontiguousarray(tables[k])
            L.orc_register_lebedev(C.c_int(a.shape[0]), a.ctypes.data_as(C.c_void_p))
        try:
            import scipy
            cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
            if cands:
                L.orc_set_blas(cands[0].encode())
        except Exception:
            pass
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Basis:
    """Keeps the numpy arrays alive behind an orc_basis_t."""

    def __init__(self, arrays: dict):
        self.a = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in arrays.items()}
        a = self.a
        self.t = BasisT(a["nsh"], a["pure"], _p(a["l"]).value, _p(a["origin"]).value, _p(a["nprim"]).value,
                        _p(a["prim_off"]).value, _p(a["exps"]).value, _p(a["coefs"]).value, _p(a["shell_atom"]).value)
        self.nsh = a["nsh"]
        self.pure = a["pure"]
        self.nbf = int(sum(lib().orc_nbf_shell(int(l), int(a["pure"])) for l in a["l"]))

    @property
    def ref(self):
        return C.byref(self.t)


def incorporate_shell_norm(l, exps, coefs):
    e = np.ascontiguousarray(exps, dtype=np.float64)
    c = np.array(coefs, dtype=np.float64)
    lib().orc_incorporate_shell_norm(C.c_int(l), C.c_int(len(e)), _p(e), _p(c))
    return c


def collocate_shell(l, pure, pts, coefs, exps, center, deriv):
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    nout = lib().orc_nbf_shell(C.c_int(l), C.c_int(pure))
    out = [np.zeros((nout, n)) for _ in range(4 if deriv else 1)]
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    exps = np.ascontiguousarray(exps, dtype=np.float64)
    center = np.ascontiguousarray(center, dtype=np.float64)
    lib().orc_collocate_shell(C.c_int(l), C.c_int(pure), C.c_size_t(n), _p(pts), C.c_int(len(exps)), _p(coefs), _p(exps),
                              _p(center), C.c_int(deriv), _p(out[0]), _p(out[1]) if deriv else None,
                              _p(out[2]) if deriv else None, _p(out[3]) if deriv else None)
    return out


def decay_cutoff(basis: Basis, threshold=1e-10):
    ext = np.zeros(basis.nsh)
    lib().orc_decay_cutoff(basis.ref, C.c_double(threshold), _p(ext))
    return ext


def default_grid_settings(**kw):
    s = GridSettingsT(302, 110, 1e-12, 1, 1)
    for k, v in kw.items():
        if k == "angular_regions":
            v = (C.c_int * 5)(*[int(x) for x in v])
        setattr(s, k, v)
    return s


# GridSettings::for_cosx (src/numint/grid_settings.cpp:10-18,71-103): IntAcc, 5 angular regions, max / min angular points
COSX_GRIDS = {1: (3.816, (14, 26, 50, 50, 26), 50, 14), 2: (4.020, (14, 26, 50, 110, 50), 110, 14),
              3: (4.338, (26, 50, 110, 194, 110), 194, 26)}


def cosx_grid_settings(level: int = 1):
    ia, regions, mx, mn = COSX_GRIDS[int(level)]
    return default_grid_settings(max_angular_points=mx, min_angular_points=mn, int_acc=ia, angular_regions=regions,
                                 reduced_first_row_element_grid=1, treutler_alrichs_adjustment=1)


def molecular_grid(atomic_numbers, positions, basis: Basis, settings=None, nthreads=None):
    Z = np.ascontiguousarray(atomic_numbers, dtype=np.int32)
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    s = settings or default_grid_settings()
    nthreads = nthreads or os.cpu_count() or 1
    g = lib().orc_molecular_grid_create(C.c_int(len(Z)), _p(Z), _p(pos), basis.ref, C.byref(s), C.c_int(nthreads))
    try:
        n = lib().orc_grid_npts(g)
        pts = np.zeros((n, 3))
        w = np.zeros(n)
        lib().orc_grid_copy(g, _p(pts), _p(w))
        info = {"atom_npts": [lib().orc_grid_atom_npts(g, a) for a in range(len(Z))],
                "atom_nradial": [lib().orc_grid_atom_nradial(g, a) for a in range(len(Z))]}
    finally:
        lib().orc_grid_destroy(g)
    return pts, w, info


def hierarchy(pts, w):
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64)
    n = len(w)
    spts = np.zeros_like(pts)
    sw = np.zeros_like(w)
    perm = np.zeros(n, dtype=np.uint64)
    cap = n // 128 + 2
    leaves = np.zeros(cap, dtype=LEAF_DTYPE)
    nl = lib().orc_hierarchy(C.c_size_t(n), _p(pts), _p(w), _p(spts), _p(sw), _p(perm), _p(leaves), C.c_size_t(cap))
    return spts, sw, perm, leaves[:nl].copy()


def xc_eval(func, polarized, rho, sigma=None, tau=None):
    fid = FUNC_IDS[func] if isinstance(func, str) else int(func)
    fam = lib().orc_xc_family(C.c_int(fid))
    s = 2 if polarized else 1
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    n = rho.size // s
    sigma = np.ascontiguousarray(sigma if sigma is not None else np.zeros(n * (3 if polarized else 1)), dtype=np.float64)
    tau = np.ascontiguousarray(tau if tau is not None else np.zeros(n * s), dtype=np.float64)
    exc = np.zeros(n)
    vrho = np.zeros(n * s)
    vsigma = np.zeros(n * (3 if polarized else 1))
    vtau = np.zeros(n * s)
    rc = lib().orc_xc_eval(C.c_int(fid), C.c_int(int(polarized)), C.c_size_t(n), _p(rho), _p(sigma), _p(tau), _p(exc),
                           _p(vrho), _p(vsigma), _p(vtau))
    if rc:
        raise RuntimeError(f"functional with id = {fid} not found")
    return {"exc": exc, "vrho": vrho.reshape(n, s), "vsigma": vsigma.reshape(n, -1), "vtau": vtau.reshape(n, s),
            "family": fam}


def method_deriv(func_names):
    return max(lib().orc_xc_family(C.c_int(FUNC_IDS[f])) for f in func_names)


def xc_build(basis: Basis, spts, sw, leaves, func_names, nspin, D2, density_threshold=1e-10,
             screening_threshold=1e-10, nthreads=None, deriv=None):
    """compute_K_dft restatement. D2: (nspin*nbf, nbf) column-major, already doubled. Returns dict."""
    ids = np.array([FUNC_IDS[f] for f in func_names], dtype=np.int32)
    if deriv is None:
        deriv = method_deriv(func_names)
    nbf = basis.nbf
    D2 = np.asfortranarray(D2, dtype=np.float64)
    assert D2.shape == (nspin * nbf, nbf), (D2.shape, nspin, nbf)
    K = np.zeros((nspin * nbf, nbf), order="F")
    exc, na, nb = C.c_double(0), C.c_double(0), C.c_double(0)
    st = StatsT()
    spts = np.ascontiguousarray(spts, dtype=np.float64)
    sw = np.ascontiguousarray(sw, dtype=np.float64)
    leaves = np.ascontiguousarray(leaves)
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().orc_xc_build(basis.ref, C.c_size_t(len(sw)), _p(spts), _p(sw), C.c_size_t(len(leaves)), _p(leaves),
                            C.c_int(len(ids)), _p(ids), C.c_int(nspin), C.c_int(deriv), C.c_double(density_threshold),
                            C.c_double(screening_threshold), _p(D2), _p(K), C.byref(exc), C.byref(na), C.byref(nb),
                            C.c_int(nthreads), C.byref(st))
    if rc:
        raise RuntimeError("orc_xc_build failed")
    stats = {f: getattr(st, f) for f, _ in StatsT._fields_}
    return {"K": K, "exc": exc.value, "n_alpha": na.value, "n_beta": nb.value, "stats": stats}


def collocate_shell_deriv2(l, pure, pts, coefs, exps, center):
    """value, x, y, z, xx, xy, xz, yy, yz, zz of one shell: array (10, nout, npts), no common_fac"""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    nout = lib().orc_nbf_shell(C.c_int(l), C.c_int(pure))
    out = np.zeros((10, nout, n))
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    exps = np.ascontiguousarray(exps, dtype=np.float64)
    center = np.ascontiguousarray(center, dtype=np.float64)
    lib().orc_collocate_shell_deriv2(C.c_int(l), C.c_int(pure), C.c_size_t(n), _p(pts), C.c_int(len(exps)), _p(coefs),
                                     _p(exps), _p(center), _p(out))
    return out


def xc_gradient(basis: Basis, spts, sw, leaves, func_names, nspin, D2, natoms, density_threshold=1e-10,
                screening_threshold=1e-10, nthreads=None):
    """compute_xc_gradient_impl restatement: returns (natoms, 3)"""
    ids = np.array([FUNC_IDS[f] for f in func_names], dtype=np.int32)
    deriv = method_deriv(func_names)
    D2 = np.asfortranarray(D2, dtype=np.float64)
    spts = np.ascontiguousarray(spts, dtype=np.float64)
    sw = np.ascontiguousarray(sw, dtype=np.float64)
    leaves = np.ascontiguousarray(leaves)
    grad = np.zeros((natoms, 3))
    rc = lib().orc_xc_gradient(basis.ref, C.c_size_t(len(sw)), _p(spts), _p(sw), C.c_size_t(len(leaves)), _p(leaves),
                               C.c_int(len(ids)), _p(ids), C.c_int(nspin), C.c_int(deriv), C.c_double(density_threshold),
                               C.c_double(screening_threshold), _p(D2), C.c_int(natoms), _p(grad),
                               C.c_int(nthreads or os.cpu_count() or 1))
    if rc:
        raise RuntimeError("orc_xc_gradient failed")
    return grad


def ref_collocate_shell_deriv2(l, pure, pts, coefs, exps, center):
    """gg_collocation_deriv2 from the compiled reference gau2grid: (10, nout, npts)"""
    R = ref_lib()
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    nout = (2 * l + 1) if pure else (l + 1) * (l + 2) // 2
    order = 300 if pure else 400
    out = np.zeros((10, nout, n))
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    exps = np.ascontiguousarray(exps, dtype=np.float64)
    center = np.ascontiguousarray(center, dtype=np.float64)
    R.gg_collocation_deriv2(C.c_int(l), C.c_ulong(n), _p(pts), C.c_ulong(3), C.c_int(len(exps)), _p(coefs), _p(exps),
                            _p(center), C.c_int(order), *[_p(out[k]) for k in range(10)])
    return out


# ---- the real reference (gau2grid) for cross-checks ----
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB_PATH):
            build()
        if not os.path.exists(REF_LIB_PATH):
            return None
        _ref = C.CDLL(REF_LIB_PATH)
    return _ref


def ref_collocate_shell(l, pure, pts, coefs, exps, center, deriv):
    """gg_collocation / gg_collocation_deriv1 from the compiled reference gau2grid."""
    R = ref_lib()
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    nout = (2 * l + 1) if pure else (l + 1) * (l + 2) // 2
    order = 300 if pure else 400  # GG_SPHERICAL_CCA / GG_CARTESIAN_CCA
    out = [np.zeros((nout, n)) for _ in range(4 if deriv else 1)]
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    exps = np.ascontiguousarray(exps, dtype=np.float64)
    center = np.ascontiguousarray(center, dtype=np.float64)
    args = [C.c_int(l), C.c_ulong(n), _p(pts), C.c_ulong(3), C.c_int(len(exps)), _p(coefs), _p(exps), _p(center),
            C.c_int(order)]
    if deriv:
        R.gg_collocation_deriv1(*args, _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]))
    else:
        R.gg_collocation(*args, _p(out[0]))
    return out


def vv10(basis: Basis, pts, w, D2, b=6.0, Cpar=0.01, nthreads=None):
    """NonLocalCorrelationFunctional::vv10 restatement (src/dft/nonlocal_correlation.cpp:38-118, 367-420).
    pts/w in the reference's NLC atom-grid order; D2 (nbf, nbf) doubled.  Returns energy, Vxc, rho (n x 4), F/U/W."""
    nbf = basis.nbf
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64)
    n = len(w)
    D2 = np.asfortranarray(D2, dtype=np.float64)
    assert D2.shape == (nbf, nbf)
    K = np.zeros((nbf, nbf), order="F")
    e = C.c_double(0)
    rho4 = np.zeros((n, 4))
    fuw = np.zeros((3, n))
    lib().orc_vv10.restype = C.c_int
    rc = lib().orc_vv10(basis.ref, C.c_size_t(n), _p(pts), _p(w), _p(D2), C.c_double(b), C.c_double(Cpar), _p(K),
                        C.byref(e), _p(rho4), _p(fuw), C.c_int(nthreads or os.cpu_count() or 1))
    if rc:
        raise RuntimeError("orc_vv10 failed")
    return {"energy": e.value, "K": K, "rho": rho4, "F": fuw[0], "U": fuw[1], "W": fuw[2]}


def vv10_gradient(basis: Basis, pts, w, atom_offsets, D2, b=6.0, Cpar=0.01, nthreads=None):
    """NonLocalCorrelationFunctional::vv10_gradient restatement (src/dft/nonlocal_correlation.cpp:108-328, 392-522).
    atom_offsets (natoms + 1): point ranges of the concatenated NLC atom grids.  Returns the gradient (natoms, 3), its
    potential and grid-displacement parts, and the energy."""
    nbf = basis.nbf
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64)
    off = np.ascontiguousarray(atom_offsets, dtype=np.uint64)
    natoms = len(off) - 1
    D2 = np.asfortranarray(D2, dtype=np.float64)
    assert D2.shape == (nbf, nbf) and int(off[-1]) == len(w)
    g = np.zeros((natoms, 3))
    parts = np.zeros((2, natoms, 3))
    e = C.c_double(0)
    lib().orc_vv10_gradient.restype = C.c_int
    rc = lib().orc_vv10_gradient(basis.ref, C.c_size_t(len(w)), _p(pts), _p(w), C.c_int(natoms), _p(off), _p(D2),
                                 C.c_double(b), C.c_double(Cpar), _p(g), C.byref(e), _p(parts),
                                 C.c_int(nthreads or os.cpu_count() or 1))
    if rc:
        raise RuntimeError("orc_vv10_gradient failed")
    return {"gradient": g, "potential": parts[0], "grid": parts[1], "energy": e.value}


# ---------------------------------------------------------------------------------------------------------
# COSX seminumerical exchange (oracle/src/cosx.cpp): SURVEY 8(f) rank 2
# ---------------------------------------------------------------------------------------------------------
class CosxStatsT(C.Structure):
    _fields_ = [("nbatches", C.c_size_t), ("nskipped_batches", C.c_size_t), ("pair_evals", C.c_size_t),
                ("dense_flops", C.c_double), ("npairs", C.c_size_t)]


def cosx_boys(x, m):
    lib().orc_cosx_boys.restype = C.c_double
    return lib().orc_cosx_boys(C.c_double(x), C.c_int(m))


def cosx_boys_reference(x, m):
    lib().orc_cosx_boys_reference.restype = C.c_double
    return lib().orc_cosx_boys_reference(C.c_double(x), C.c_int(m))


def cosx_ecoeffs_1d(la, lb, a, b, XAB):
    out = np.zeros((la + 1, lb + 1, la + lb + 1))
    lib().orc_cosx_ecoeffs_1d(C.c_int(la), C.c_int(lb), C.c_double(a), C.c_double(b), C.c_double(XAB), _p(out))
    return out


def cosx_rints(L, p, PC):
    out = np.zeros((L + 1) * (L + 2) * (L + 3) // 6)
    lib().orc_cosx_rints(C.c_int(L), C.c_double(p), C.c_double(PC[0]), C.c_double(PC[1]), C.c_double(PC[2]), _p(out))
    return out


def cosx_esp_pair(la, lb, ea, eb, ca, cb, A, B, spherical, pts, omega=0.0):
    """ESPEvaluator::evaluate for one shell pair: V (npts, na, nb)"""
    ea, eb, ca, cb, A, B = (np.ascontiguousarray(x, dtype=np.float64) for x in (ea, eb, ca, cb, A, B))
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    na = lib().orc_nbf_shell(C.c_int(la), C.c_int(int(spherical)))
    nb = lib().orc_nbf_shell(C.c_int(lb), C.c_int(int(spherical)))
    V = np.zeros((len(pts), na, nb))
    lib().orc_cosx_esp_pair(C.c_int(la), C.c_int(lb), C.c_int(len(ea)), C.c_int(len(eb)), _p(ea), _p(eb), _p(ca), _p(cb),
                            _p(A), _p(B), C.c_int(int(spherical)), C.c_size_t(len(pts)), _p(pts), C.c_double(omega),
                            _p(V))
    return V


def cosx_pairs(basis: Basis):
    """significant shell pairs (p >= q) in the reference's registration order, with centre and extent"""
    lib().orc_cosx_pairs.restype = C.c_size_t
    n = lib().orc_cosx_pairs(basis.ref, None, None, None)
    pq = np.zeros((n, 2), dtype=np.int32)
    cen = np.zeros((n, 3))
    ext = np.zeros(n)
    lib().orc_cosx_pairs(basis.ref, _p(pq), _p(cen), _p(ext))
    return pq, cen, ext


def cosx_overlap(basis: Basis):
    S = np.zeros((basis.nbf, basis.nbf), order="F")
    lib().orc_cosx_overlap(basis.ref, _p(S))
    return S


def cosx_atom_grids(pts, info):
    """atom offsets (natoms + 1) of a grid in atom-grid order (the layout of SemiNumericalExchange::m_atom_grids)"""
    return np.concatenate([[0], np.cumsum(info["atom_npts"])]).astype(np.uint64)


def cosx_numerical_overlap(basis: Basis, atom_off, pts, w):
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64)
    atom_off = np.ascontiguousarray(atom_off, dtype=np.uint64)
    Sn = np.zeros((basis.nbf, basis.nbf), order="F")
    lib().orc_cosx_numerical_overlap(basis.ref, C.c_int(len(atom_off) - 1), _p(atom_off), _p(pts), _p(w), _p(Sn))
    return Sn


def cosx_overlap_projector(basis: Basis, atom_off, pts, w):
    """m_overlap_projector = Sn.ldlt().solve(S)  (seminumerical_exchange.cpp:258-262)"""
    S = cosx_overlap(basis)
    Sn = cosx_numerical_overlap(basis, atom_off, pts, w)
    return np.asfortranarray(np.linalg.solve(Sn, S)), S, Sn


def cosx_K_for_density(basis: Basis, atom_off, pts, w, D2q, precision=np.finfo(float).eps, omega=0.0, margin=1.0):
    """compute_K_for_density (seminumerical_exchange.cpp:817-828): raw K of one projected density"""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64)
    atom_off = np.ascontiguousarray(atom_off, dtype=np.uint64)
    D2q = np.asfortranarray(D2q, dtype=np.float64)
    K = np.zeros((basis.nbf, basis.nbf), order="F")
    st = CosxStatsT()
    rc = lib().orc_cosx_K_for_density(basis.ref, C.c_int(len(atom_off) - 1), _p(atom_off), _p(pts), _p(w), _p(D2q),
                                      C.c_double(precision), C.c_double(omega), C.c_double(margin), _p(K), C.byref(st))
    if rc:
        raise RuntimeError("orc_cosx_K_for_density failed")
    return K, {f: getattr(st, f) for f, _ in CosxStatsT._fields_}


def cosx_compute_K(basis: Basis, atom_off, pts, w, projector, D, unrestricted=False, precision=np.finfo(float).eps,
                   omega=0.0, margin=1.0):
    """SemiNumericalExchange::compute_K, ESP mode (seminumerical_exchange.cpp:517-545 restricted, 617-660
    unrestricted). D is mo.D: (nbf, nbf), or (2 nbf, nbf) with the alpha block on top."""
    nbf = basis.nbf
    stats = []
    if not unrestricted:
        K, st = cosx_K_for_density(basis, atom_off, pts, w, projector @ (2.0 * D), precision, omega, margin)
        return np.asfortranarray(0.25 * (K + K.T)), [st]
    out = np.zeros((2 * nbf, nbf), order="F")
    for s in range(2):
        Ds = 2.0 * D[s * nbf:(s + 1) * nbf]
        K, st = cosx_K_for_density(basis, atom_off, pts, w, projector @ Ds, precision, omega, margin)
        out[s * nbf:(s + 1) * nbf] = 0.5 * (K + K.T)
        stats.append(st)
    return out, stats


def cosx_compute_K_general(basis: Basis, atom_off, pts, w, projector, D, precision=np.finfo(float).eps, omega=0.0,
                           margin=1.0):
    """SemiNumericalExchange::compute_K_general (seminumerical_exchange.cpp:658-712): D, K are (2 nbf, 2 nbf); the blocks
    follow qm::block (include/occ/qm/opmatrix.h:14-28): ab = bottom-left, ba = top-right."""
    n = basis.nbf
    blk = {"aa": (slice(0, n), slice(0, n)), "ab": (slice(n, 2 * n), slice(0, n)), "ba": (slice(0, n), slice(n, 2 * n)),
           "bb": (slice(n, 2 * n), slice(n, 2 * n))}
    Kr = {k: cosx_K_for_density(basis, atom_off, pts, w, projector @ (2.0 * D[v]), precision, omega, margin)[0]
          for k, v in blk.items()}
    K = np.zeros((2 * n, 2 * n), order="F")
    K[blk["aa"]] = 0.5 * (Kr["aa"] + Kr["aa"].T)
    K[blk["bb"]] = 0.5 * (Kr["bb"] + Kr["bb"].T)
    K[blk["ab"]] = 0.5 * (Kr["ab"] + Kr["ba"].T)
    K[blk["ba"]] = 0.5 * (Kr["ba"] + Kr["ab"].T)
    return K


def ref_boys_table():
    R = ref_lib()
    n = R.occ_ref_boys_table(None)
    t = np.zeros(n)
    R.occ_ref_boys_table(_p(t))
    return t


def ref_boys(x, m):
    R = ref_lib()
    R.occ_ref_boys.restype = C.c_double
    return R.occ_ref_boys(C.c_double(x), C.c_int(m))


def ref_ecoeffs_1d(la, lb, a, b, XAB):
    out = np.zeros((la + 1, lb + 1, la + lb + 1))
    ref_lib().occ_ref_ecoeffs_1d(C.c_int(la), C.c_int(lb), C.c_double(a), C.c_double(b), C.c_double(XAB), _p(out))
    return out


def ref_rints(L, p, PC):
    out = np.zeros((L + 1) * (L + 2) * (L + 3) // 6)
    ref_lib().occ_ref_rints(C.c_int(L), C.c_double(p), C.c_double(PC[0]), C.c_double(PC[1]), C.c_double(PC[2]), _p(out))
    return out
