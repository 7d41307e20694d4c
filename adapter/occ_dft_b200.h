// adapter/occ_dft_b200.h — header-only C++17 host side over the C ABI (include/occ_xc_b200.h).
//
// Mirrors the piece of occ::dft::DFT that owns the XC build (reference: include/occ/dft/dft.h:36-763):
//   XCBuild::compute_K_dft(D, K)            <->  Mat DFT::compute_K_dft<deriv, kind>(mo, Schwarz)   dft.h:190-388
//   XCBuild::exchange_correlation_energy()  <->  m_exc_dft (dft.h:202,377)
//   XCBuild::total_density_alpha/beta()     <->  the debug totals of dft.h:384-385
//   errors                                  <->  std::runtime_error, as dft.h:409,422,427 / functional.cpp:32,61
// No Eigen dependency: matrices cross as raw column-major pointers, which is what Eigen's Mat::data() gives
// (core/linear_algebra.h:23).  INTEGRATION.md shows the ~25-line body that replaces dft.h:216-382 with this class.
#pragma once
#include "../include/occ_xc_b200.h"
#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

namespace occ_b200 {

enum class SpinorbitalKind { Restricted = 1, Unrestricted = 2 }; // occ::qm::SpinorbitalKind (qm/spinorbital.h)

// Flat copy of occ::gto::AOBasis: fill from basis.shells() (Shell::l, origin, exponents,
// contraction_coefficients.col(0)) and basis.shell_to_atom(); kind() == Spherical -> pure = true.
struct BasisArrays {
  bool pure = false;
  std::vector<int> l, nprim, prim_off, shell_atom;
  std::vector<double> origin, exps, coefs;

  void add_shell(int L, const double *xyz, int np, const double *alpha, const double *c, int atom) {
    l.push_back(L);
    nprim.push_back(np);
    prim_off.push_back(static_cast<int>(exps.size()));
    shell_atom.push_back(atom);
    origin.insert(origin.end(), xyz, xyz + 3);
    exps.insert(exps.end(), alpha, alpha + np);
    coefs.insert(coefs.end(), c, c + np);
  }
  occ_xc_basis_t view() const {
    occ_xc_basis_t b;
    b.nsh = static_cast<int>(l.size());
    b.pure = pure ? 1 : 0;
    b.l = l.data();
    b.origin = origin.data();
    b.nprim = nprim.data();
    b.prim_off = prim_off.data();
    b.exps = exps.data();
    b.coefs = coefs.data();
    b.shell_atom = shell_atom.data();
    return b;
  }
  int nbf() const {
    int n = 0;
    for (int L : l) n += pure ? 2 * L + 1 : (L + 1) * (L + 2) / 2;
    return n;
  }
};

inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(occ_xc_last_error());
}

// Owns one occ_xc_handle: everything geometry-only (grid, shell lists, work lists) is uploaded once, when the
// reference would (re)build its grid (DFT::set_integration_grid / set_method, src/dft/dft.cpp:56-117).
class XCBuild {
public:
  // sorted_points (3 x npts, column-major), sorted_weights and leaves come from
  // m_grid.get_molecular_grid_points().get_hierarchy() (dft.h:295-298); func_ids from
  // m_method.functionals[_polarized][i].id() (dft.h:207-209).  nocc_alpha >= 0 enables compute_K_dft_mo (the handle
  // then also holds the factorised-density work lists); nocc_beta = 0 is legal for unrestricted one-electron systems.
  XCBuild(const BasisArrays &basis, size_t npts, const double *sorted_points, const double *sorted_weights,
          const std::vector<occ_xc_leaf_t> &leaves, const std::vector<int> &func_ids, SpinorbitalKind kind,
          double density_threshold = 1e-10, double xc_screening_threshold = 1e-10, int device = -1, int rank = 0,
          int world_size = 1, int nocc_alpha = -1, int nocc_beta = -1, bool deterministic = false,
          const std::vector<int> &devices = {}, bool gradient = false)
      : m_nspin(static_cast<int>(kind)), m_nbf(basis.nbf()) {
    occ_xc_options_t opt;
    occ_xc_options_init(&opt);
    opt.device = device;
    opt.rank = rank;
    opt.world_size = world_size;
    opt.deterministic = deterministic ? 1 : 0;
    opt.gradient = gradient ? 1 : 0; // the handle then also serves compute_xc_gradient (needs basis.shell_atom)
    // The reference hands over mo.D and mo.Cocc; D2 = 2 mo.D (dft.h:205) and C~ = sqrt(2) Cocc (restricted: mo.D =
    // Cocc Cocc^T; unrestricted: mo.D_s = 1/2 C_s C_s^T, include/occ/qm/orb.h:64-82) are formed ON THE DEVICE
    opt.density_scale = 2.0;
    opt.orbital_scale = m_nspin == 1 ? 1.4142135623730951 : 1.0;
    if (nocc_alpha >= 0) {
      opt.mo_form = 1;
      opt.nocc_alpha = nocc_alpha;
      opt.nocc_beta = m_nspin == 2 ? (nocc_beta >= 0 ? nocc_beta : nocc_alpha) : 0;
      m_nocc[0] = opt.nocc_alpha;
      m_nocc[1] = opt.nocc_beta;
    }
    const occ_xc_basis_t b = basis.view();
    // `devices` non-empty: one leaf shard per listed GPU, all driven from this process (occ_xc_multi_*) — the drop-in
    // for the reference's tbb::parallel_for over leaves + serial reduction of thread copies (dft.h:303-371,375-380)
    if (!devices.empty())
      check(occ_xc_multi_create(&b, npts, sorted_points, sorted_weights, leaves.size(), leaves.data(), nullptr,
                                static_cast<int>(func_ids.size()), func_ids.data(), m_nspin, density_threshold,
                                xc_screening_threshold, &opt, static_cast<int>(devices.size()), devices.data(), &m_multi));
    else
      check(occ_xc_create(&b, npts, sorted_points, sorted_weights, leaves.size(), leaves.data(), nullptr,
                          static_cast<int>(func_ids.size()), func_ids.data(), m_nspin, density_threshold,
                          xc_screening_threshold, &opt, &m_handle));
  }
  ~XCBuild() {
    for (void *p : m_registered) occ_xc_host_unregister(p);
    occ_xc_pinned_free(m_K);
    occ_xc_destroy(m_handle);
    occ_xc_multi_destroy(m_multi);
  }
  XCBuild(const XCBuild &) = delete;
  XCBuild &operator=(const XCBuild &) = delete;

  // Page-lock caller storage that persists across SCF iterations (e.g. the Eigen storage of mo.D / mo.Cocc / K): the
  // copies then run at PCIe speed instead of through the driver's bounce buffer.  Unregistered in the destructor.
  void register_host_buffer(const void *p, size_t bytes) {
    check(occ_xc_host_register(const_cast<void *>(p), bytes));
    m_registered.push_back(const_cast<void *>(p));
  }

  // D = mo.D ((nspin*nbf) x nbf column-major), exactly as the reference stores it. Writes K (same shape).
  // No host-side pass over D: the factor 2 of dft.h:205 is applied on the device (options.density_scale).
  void compute_K_dft(const double *D, double *K) {
    if (m_multi) check(occ_xc_multi_build(m_multi, D, nullptr, K, &m_exc_dft, &m_total_density_a, &m_total_density_b));
    else check(occ_xc_build(m_handle, D, K, &m_exc_dft, &m_total_density_a, &m_total_density_b));
  }
  // Same, but K stays in a page-locked buffer owned by this object (valid until the next call): the caller maps it
  // (Eigen::Map<const Mat>(ptr, nspin * nbf, nbf)) — the reference negates / copies K anyway (dft.h:435).
  const double *compute_K_dft(const double *D) {
    ensure_K();
    compute_K_dft(D, m_K);
    return m_K;
  }
  // The same build with the occupied orbitals of `mo` (mo.Cocc; restricted: nbf x nocc, unrestricted: one
  // nbf x nocc_s matrix per spin, the beta block directly behind the alpha block as in mo.Cocc's column blocks),
  // column-major.  D may be nullptr: the density matrix is then formed on the device from the orbitals and nothing
  // but 8 nbf nocc bytes crosses PCIe on the way in.
  void compute_K_dft_mo(const double *D, const double *Cocc, double *K) {
    if (m_nocc[0] < 0) throw std::runtime_error("XCBuild was created without occupied-orbital counts");
    if (m_multi) check(occ_xc_multi_build(m_multi, D, Cocc, K, &m_exc_dft, &m_total_density_a, &m_total_density_b));
    else check(occ_xc_build_mo(m_handle, D, Cocc, K, &m_exc_dft, &m_total_density_a, &m_total_density_b));
  }
  // two separate blocks (Cocc_alpha, Cocc_beta): packed into a buffer of this object
  void compute_K_dft_mo(const double *D, const double *Cocc_alpha, const double *Cocc_beta, double *K) {
    if (m_nocc[0] < 0) throw std::runtime_error("XCBuild was created without occupied-orbital counts");
    const size_t na = static_cast<size_t>(m_nbf) * m_nocc[0], nb = m_nspin == 2 ? static_cast<size_t>(m_nbf) * m_nocc[1] : 0;
    if (nb == 0 || Cocc_beta == Cocc_alpha + na) return compute_K_dft_mo(D, Cocc_alpha, K);
    if (m_C.size() != na + nb) m_C.resize(na + nb);
    std::copy(Cocc_alpha, Cocc_alpha + na, m_C.begin());
    std::copy(Cocc_beta, Cocc_beta + nb, m_C.begin() + na);
    compute_K_dft_mo(D, m_C.data(), K);
  }
  // DFT::compute_xc_gradient_impl (dft.h:598-749) without the VV10 term: D = mo.D, gradient 3 x natoms column-major
  // (Mat3N).  Needs `gradient = true` at construction and a single-device handle.
  void compute_xc_gradient(const double *D, double *gradient) {
    if (m_multi) throw std::runtime_error("compute_xc_gradient: single-device handle required");
    check(occ_xc_gradient(m_handle, D, gradient));
  }
  void compute_xc_gradient_mo(const double *D, const double *Cocc, double *gradient) {
    if (m_multi) throw std::runtime_error("compute_xc_gradient_mo: single-device handle required");
    check(occ_xc_gradient_mo(m_handle, D, Cocc, gradient));
  }
  double exchange_correlation_energy() const { return m_exc_dft; }
  double total_density_alpha() const { return m_total_density_a; }
  double total_density_beta() const { return m_total_density_b; }
  int density_derivative() const { return info().deriv; } // DFT::density_derivative, src/dft/dft.cpp:11-17
  occ_xc_info_t info(int shard = 0) const {
    occ_xc_info_t i;
    if (m_multi) check(occ_xc_multi_get_info(m_multi, shard, &i));
    else check(occ_xc_get_info(m_handle, &i));
    return i;
  }
  int num_devices() const { return m_multi ? occ_xc_multi_ndev(m_multi) : 1; }
  occ_xc_handle *handle() { return m_handle; }

private:
  occ_xc_handle *m_handle = nullptr;
  occ_xc_multi *m_multi = nullptr;
  int m_nspin, m_nbf;
  double m_exc_dft = 0.0, m_total_density_a = 0.0, m_total_density_b = 0.0;
  int m_nocc[2] = {-1, -1};
  double *m_K = nullptr;
  std::vector<double> m_C;
  std::vector<void *> m_registered;
  void ensure_K() {
    if (m_K) return;
    void *p = nullptr;
    check(occ_xc_pinned_alloc(static_cast<size_t>(m_nspin) * m_nbf * m_nbf * sizeof(double), &p));
    m_K = static_cast<double *>(p);
  }
};

// occ::dft::NonLocalCorrelationFunctional (include/occ/dft/nonlocal_correlation.h): the post-SCF VV10 correction and its
// nuclear gradient.  Created over the NLC grid — hierarchy of the concatenated m_nlc_atom_grids
// (set_integration_grid, nonlocal_correlation.cpp:17-31): sorted points / weights / leaves, the sort permutation
// (sorted index -> atom-grid order) and the prefix sums of the atom-grid sizes.
class NonLocalCorrelation {
public:
  struct Parameters {
    double vv10_b{6.0};
    double vv10_C{0.01};
  }; // nonlocal_correlation.h:14-18
  NonLocalCorrelation(const BasisArrays &basis, size_t npts, const double *sorted_points, const double *sorted_weights,
                      const std::vector<occ_xc_leaf_t> &leaves, const std::vector<size_t> &sort_permutation,
                      const std::vector<size_t> &atom_offsets, bool gradient = false, int device = -1,
                      const std::vector<int> &devices = {})
      : m_nbf(basis.nbf()), m_perm(sort_permutation), m_atom_offsets(atom_offsets) {
    occ_xc_options_t opt;
    occ_xc_options_init(&opt);
    opt.device = device;
    opt.gradient = gradient ? 1 : 0;
    opt.density_scale = 2.0; // the caller passes mo.D; D2 = 2 mo.D (nonlocal_correlation.cpp:376) is formed on the device
    const occ_xc_basis_t b = basis.view();
    const int b88 = 106; // any GGA id: only rho and grad rho of the handle are used
    if (!devices.empty())
      check(occ_xc_multi_create(&b, npts, sorted_points, sorted_weights, leaves.size(), leaves.data(), nullptr, 1, &b88, 1,
                                1e-10, 0.0, &opt, static_cast<int>(devices.size()), devices.data(), &m_multi));
    else
      check(occ_xc_create(&b, npts, sorted_points, sorted_weights, leaves.size(), leaves.data(), nullptr, 1, &b88, 1, 1e-10,
                          0.0, &opt, &m_handle));
  }
  ~NonLocalCorrelation() {
    occ_xc_destroy(m_handle);
    occ_xc_multi_destroy(m_multi);
  }
  NonLocalCorrelation(const NonLocalCorrelation &) = delete;
  NonLocalCorrelation &operator=(const NonLocalCorrelation &) = delete;
  void set_parameters(const Parameters &p) { m_params = p; }

  // Result operator()(basis, mo) (nonlocal_correlation.cpp:524-527): D = mo.D (restricted); Vxc nbf x nbf; returns the energy
  double vv10(const double *D, double *Vxc) {
    double e = 0.0;
    if (m_multi) check(occ_xc_multi_vv10(m_multi, D, m_perm.data(), m_params.vv10_b, m_params.vv10_C, Vxc, &e));
    else check(occ_xc_vv10(m_handle, D, m_perm.data(), m_params.vv10_b, m_params.vv10_C, Vxc, &e));
    return e;
  }
  // GradientResult compute_gradient(basis, mo) (nonlocal_correlation.cpp:530-534): gradient 3 x natoms (Mat3N); returns the energy
  double compute_gradient(const double *D, double *gradient) {
    double e = 0.0;
    if (m_multi)
      check(occ_xc_multi_vv10_gradient(m_multi, D, m_perm.data(), m_atom_offsets.data(), m_params.vv10_b, m_params.vv10_C,
                                       gradient, &e));
    else
      check(occ_xc_vv10_gradient(m_handle, D, m_perm.data(), m_atom_offsets.data(), m_params.vv10_b, m_params.vv10_C,
                                 gradient, &e));
    return e;
  }

private:
  occ_xc_handle *m_handle = nullptr;
  occ_xc_multi *m_multi = nullptr;
  int m_nbf;
  Parameters m_params;
  std::vector<size_t> m_perm, m_atom_offsets;
};

} // namespace occ_b200
