// adapter/occ_cosx_b200.h — header-only C++17 host side over the COSX C ABI (include/occ_cosx_b200.h).
//
// Mirrors occ::qm::cosx::SemiNumericalExchange in ESP mode (reference: include/occ/qm/seminumerical_exchange.h:104-205):
//   SemiNumericalExchangeB200::compute_K(kind, D, K, precision)     <->  Mat compute_K(mo, precision, Schwarz)   :481-515
//   ...::compute_K_range_separated(D, alpha, beta, omega, K)        <->  compute_K_range_separated               :830-860
//   ...::set_range_separated_omega / range_separated_omega          <->  same names                              .h:157-158
//   ...::num_grid_points / num_batches / num_atoms                  <->  same names                              :430-452
//   errors                                                          <->  std::runtime_error (:489-493)
// The constructor takes what the reference's constructor leaves in its members: the basis, the Becke-partitioned atom
// grids (m_atom_grids) and m_overlap_projector.  Matrices cross as raw column-major pointers (Eigen's Mat::data()).
#pragma once
#include "../include/occ_cosx_b200.h"
#include "occ_dft_b200.h"
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <vector>

namespace occ_b200 {

class SemiNumericalExchangeB200 {
public:
  // atom_offsets: natoms + 1 entries delimiting the atom grids in pts (3 x N, column-major = xyz interleaved) / weights
  SemiNumericalExchangeB200(const BasisArrays &basis, const std::vector<size_t> &atom_offsets, const double *pts,
                            const double *weights, const double *overlap_projector, double margin = 1.0,
                            int device = -1, const std::vector<int> &devices = {})
      : m_nbf(basis.nbf()), m_natoms(static_cast<int>(atom_offsets.size()) - 1) {
    occ_cosx_options_t opt = {};
    opt.margin = margin;
    opt.device = device;
    occ_xc_basis_t b = basis.view();
    // `devices` non-empty: one shard of grid batches per listed GPU, all driven from this process (occ_cosx_multi_*)
    if (!devices.empty())
      check(occ_cosx_multi_create(&b, m_natoms, atom_offsets.data(), pts, weights, overlap_projector, &opt,
                                  static_cast<int>(devices.size()), devices.data(), &m_multi));
    else
      check(occ_cosx_create(&b, m_natoms, atom_offsets.data(), pts, weights, overlap_projector, &opt, &m_h));
  }
  SemiNumericalExchangeB200(const SemiNumericalExchangeB200 &) = delete;
  SemiNumericalExchangeB200 &operator=(const SemiNumericalExchangeB200 &) = delete;
  ~SemiNumericalExchangeB200() {
    occ_cosx_destroy(m_h);
    occ_cosx_multi_destroy(m_multi);
  }

  // D = mo.D (nbf x nbf restricted, 2 nbf x nbf unrestricted), K of the same shape
  void compute_K(SpinorbitalKind kind, const double *D, double *K,
                 double precision = std::numeric_limits<double>::epsilon()) const {
    const int k = kind == SpinorbitalKind::Unrestricted ? 1 : 0;
    if (m_multi) check(occ_cosx_multi_compute_K(m_multi, k, D, nullptr, 0, 0, precision, m_omega, K));
    else check(occ_cosx_compute_K(m_h, k, D, precision, m_omega, K));
  }
  // the same with the occupied orbitals (mo.Cocc) as well: C = [C_alpha | C_beta], nbf rows, scaled so that 2 D_s = C_s C_s^T
  // (restricted: sqrt(2) Cocc; unrestricted: Cocc_s). F = Phi.D2q runs in the factorised form (occ_cosx_compute_K_mo).
  void compute_K_mo(SpinorbitalKind kind, const double *D, const double *C, int nocc_alpha, int nocc_beta, double *K,
                    double precision = std::numeric_limits<double>::epsilon()) const {
    const int k = kind == SpinorbitalKind::Unrestricted ? 1 : 0;
    if (m_multi) check(occ_cosx_multi_compute_K(m_multi, k, D, C, nocc_alpha, nocc_beta, precision, m_omega, K));
    else check(occ_cosx_compute_K_mo(m_h, k, D, C, nocc_alpha, nocc_beta, precision, m_omega, K));
  }
  // general spinorbitals: D and K are 2 nbf x 2 nbf (compute_K_general, seminumerical_exchange.cpp:658-712)
  void compute_K_general(const double *D, double *K, double precision = std::numeric_limits<double>::epsilon()) const {
    if (m_multi) check(occ_cosx_multi_compute_K(m_multi, 2, D, nullptr, 0, 0, precision, m_omega, K));
    else check(occ_cosx_compute_K(m_h, 2, D, precision, m_omega, K));
  }
  void compute_K_range_separated(const double *D, double alpha, double beta, double omega, double *K,
                                 double precision = std::numeric_limits<double>::epsilon()) const {
    if (m_multi) check(occ_cosx_multi_compute_K_range_separated(m_multi, 0, D, nullptr, 0, 0, alpha, beta, omega, precision, K));
    else check(occ_cosx_compute_K_range_separated(m_h, D, alpha, beta, omega, precision, K));
  }
  void set_range_separated_omega(double omega) { m_omega = omega; }
  double range_separated_omega() const { return m_omega; }

  size_t num_grid_points() const { return info().npts; }
  size_t num_batches() const { return info().nbatches; }
  size_t num_atoms() const { return static_cast<size_t>(m_natoms); }
  int nbf() const { return m_nbf; }
  occ_cosx_info_t info() const {
    occ_cosx_info_t i;
    if (m_multi) { // totals over the shards
      check(occ_cosx_multi_get_info(m_multi, 0, &i));
      for (int r = 1; r < occ_cosx_multi_ndev(m_multi); ++r) {
        occ_cosx_info_t j;
        check(occ_cosx_multi_get_info(m_multi, r, &j));
        i.npts += j.npts;
        i.nbatches += j.nbatches;
      }
    } else {
      check(occ_cosx_get_info(m_h, &i));
    }
    return i;
  }

private:
  occ_cosx_handle *m_h = nullptr;
  occ_cosx_multi *m_multi = nullptr;
  int m_nbf, m_natoms;
  double m_omega = 0.0;
};

} // namespace occ_b200
