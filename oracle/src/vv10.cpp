// oracle/src/vv10.cpp — CPU ORACLE (test infrastructure): the post-SCF VV10 non-local correlation of the reference,
// SURVEY.md 8(f) rank 4.  Restates (paths under /root/reference):
//   src/dft/nonlocal_correlation.cpp:38-118     vv10_kernel  (the pair sum runs over j < i ONLY, in the order the
//                                               points are handed over; self term seeded into f and u; points with
//                                               rho < 1e-8 neither receive nor contribute)
//   src/dft/nonlocal_correlation.cpp:367-420    NonLocalCorrelationFunctional::vv10  (dense evaluate_basis over all
//                                               points, restricted deriv-1 density, E = sum exc . rho,
//                                               Vxc = ktmp + ktmp^T with ktmp = phi^T (1/2 vrho phi + 2 vsigma grad rho . grad phi))
// Pins: the reference's one VV10-bearing assertion set, the wB97X-V isolated-XC case (tests/isolated_xc_tests.cpp:340-385:
// local gradient + VV10 gradient against VV10-free PySCF numbers, margins 7e-3 / 5e-3 / 4e-3 and symmetry 1e-5), is met
// and its deviations (+6.03e-3, +4.49e-3, -3.07e-3; asymmetry 5.4e-6) are reproduced (tests/test_xc_gradient.py); beyond
// that the restatement is checked against an independent numpy transcription of the same lines (tests/test_vv10.py).
// No tighter reference-held number exists for the VV10 energy or Vxc themselves.
#include "oracle.h"
#include "oracle_internal.h"
#include <cmath>
#include <omp.h>
#include <vector>

namespace {
constexpr double kPi = 3.14159265358979323846;
}

extern "C" {

/* rho4: npts x 4 row-major on output (rho, rho_x, rho_y, rho_z) if non-NULL; FUW: 3 x npts if non-NULL */
int orc_vv10(const orc_basis_t *basis, size_t npts, const double *pts, const double *w, const double *D2, double b,
             double C, double *K, double *energy, double *rho4_out, double *fuw_out, int nthreads) {
  using namespace orc;
  const Basis B(basis);
  const orc_basis_t &raw = B.raw;
  const int nbf = B.nbf;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  // upper-triangle-symmetrised D (evaluate_density<1, Restricted> uses D.selfadjointView<Upper>, density.cpp:38)
  std::vector<double> Ds(static_cast<size_t>(nbf) * nbf);
  for (int j = 0; j < nbf; ++j)
    for (int i = 0; i < nbf; ++i) {
      const int r = std::min(i, j), c = std::max(i, j);
      Ds[static_cast<size_t>(j) * nbf + i] = D2[static_cast<size_t>(c) * nbf + r];
    }
  std::vector<double> rho(npts), gx(npts), gy(npts), gz(npts);
  constexpr size_t BLK = 128;
  const size_t nblk = (npts + BLK - 1) / BLK;

  auto collocate_block = [&](size_t p0, int n, std::vector<double> *phi) {
    for (int c = 0; c < 4; ++c) phi[c].assign(static_cast<size_t>(n) * nbf, 0.0);
    std::vector<double> tmp[4];
    for (int s = 0; s < raw.nsh; ++s) {
      const int l = raw.l[s], sz = orc_nbf_shell(l, raw.pure);
      for (int c = 0; c < 4; ++c) tmp[c].assign(static_cast<size_t>(sz) * n, 0.0);
      collocate_shell(l, raw.pure, n, pts + 3 * p0, raw.nprim[s], raw.coefs + raw.prim_off[s], raw.exps + raw.prim_off[s],
                      raw.origin + 3 * s, 1, tmp[0].data(), tmp[1].data(), tmp[2].data(), tmp[3].data());
      const double fac = orc_common_fac(l, raw.pure);
      for (int c = 0; c < 4; ++c)
        for (int f = 0; f < sz; ++f)
          for (int p = 0; p < n; ++p)
            phi[c][static_cast<size_t>(B.first_bf[s] + f) * n + p] = fac * tmp[c][static_cast<size_t>(f) * n + p];
    }
  };

#pragma omp parallel
  {
    std::vector<double> phi[4], X;
#pragma omp for schedule(dynamic)
    for (size_t blk = 0; blk < nblk; ++blk) {
      const size_t p0 = blk * BLK;
      const int n = static_cast<int>(std::min(BLK, npts - p0));
      collocate_block(p0, n, phi);
      X.assign(static_cast<size_t>(n) * nbf, 0.0);
      dgemm('N', 'N', n, nbf, nbf, 1.0, phi[0].data(), n, Ds.data(), nbf, 0.0, X.data(), n);
      for (int p = 0; p < n; ++p) {
        double r = 0, x = 0, y = 0, z = 0;
        for (int f = 0; f < nbf; ++f) {
          const size_t o = static_cast<size_t>(f) * n + p;
          r += phi[0][o] * X[o];
          x += phi[1][o] * X[o];
          y += phi[2][o] * X[o];
          z += phi[3][o] * X[o];
        }
        rho[p0 + p] = r;
        gx[p0 + p] = 2.0 * x;
        gy[p0 + p] = 2.0 * y;
        gz[p0 + p] = 2.0 * z;
      }
    }
  }

  // ---- vv10_kernel (nonlocal_correlation.cpp:38-118) ----
  const double kappa_factor = b * 1.5 * kPi * std::pow(9.0 * kPi, -1.0 / 6.0);
  const double beta = std::pow(3.0 / (b * b), 0.75) / 32.0;
  const double pi43 = 4.0 * kPi / 3.0;
  std::vector<double> rw(npts), om(npts), kap(npts), dom_dr(npts), dom_dg(npts), dk_dr(npts), F(npts, 0.0), U(npts, 0.0),
      W(npts, 0.0);
  for (size_t i = 0; i < npts; ++i) {
    const double g = gx[i] * gx[i] + gy[i] * gy[i] + gz[i] * gz[i]; // grad_rho() returns |grad rho|^2
    rw[i] = rho[i] * w[i];
    double t = g / (rho[i] * rho[i]);
    t = C * t * t;
    om[i] = std::sqrt(t + pi43 * rho[i]);
    dom_dr[i] = (0.5 * pi43 * rho[i] - 2.0 * t) / om[i];
    dom_dg[i] = t * rho[i] / (g * om[i]);
    kap[i] = kappa_factor * std::pow(rho[i], 1.0 / 6.0);
    dk_dr[i] = kap[i] / 6.0;
  }
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < npts; ++i) {
    if (rho[i] < 1e-8) continue;
    const double k = kap[i];
    double f = rw[i] / (2.0 * k * k * k);
    double u = f * (1.0 / k + 1.0 / (2.0 * k));
    double ww = 0.0;
    for (size_t j = 0; j < i; ++j) {
      if (rho[j] < 1e-8) continue;
      const double dx = pts[3 * i] - pts[3 * j], dy = pts[3 * i + 1] - pts[3 * j + 1], dz = pts[3 * i + 2] - pts[3 * j + 2];
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double gi = r2 * om[i] + kap[i], gj = r2 * om[j] + kap[j];
      double t = 2.0 * rw[j] / (gi * gj * (gi + gj));
      f += t;
      t *= 1.0 / gi + 1.0 / (gi + gj);
      u += t;
      ww += t * r2;
    }
    F[i] = -1.5 * f;
    U[i] = u;
    W[i] = ww;
  }
  std::vector<double> vrho(npts), vsig(npts);
  double etot = 0.0;
  for (size_t i = 0; i < npts; ++i) {
    const double exc = (beta + 0.5 * F[i]) * w[i];
    etot += exc * rho[i];
    // points below the density cut have F = U = W = 0; the products with d omega / d kappa are skipped for them
    // (the reference multiplies 0 by a possibly non-finite derivative when rho underflows to exactly 0)
    vrho[i] = rho[i] < 1e-8 ? beta * w[i] : (beta + F[i] + 1.5 * (U[i] * dk_dr[i] + W[i] * dom_dr[i])) * w[i];
    vsig[i] = (rho[i] < 1e-8 || W[i] == 0.0) ? 0.0 : 1.5 * W[i] * dom_dg[i] * w[i];
  }
  *energy = etot;
  if (rho4_out)
    for (size_t i = 0; i < npts; ++i) {
      rho4_out[4 * i] = rho[i];
      rho4_out[4 * i + 1] = gx[i];
      rho4_out[4 * i + 2] = gy[i];
      rho4_out[4 * i + 3] = gz[i];
    }
  if (fuw_out)
    for (size_t i = 0; i < npts; ++i) {
      fuw_out[i] = F[i];
      fuw_out[npts + i] = U[i];
      fuw_out[2 * npts + i] = W[i];
    }

  // ---- Vxc (nonlocal_correlation.cpp:408-419) ----
  std::vector<double> Kt(static_cast<size_t>(nbf) * nbf, 0.0);
#pragma omp parallel
  {
    std::vector<double> phi[4], G, Kl(static_cast<size_t>(nbf) * nbf, 0.0);
#pragma omp for schedule(dynamic)
    for (size_t blk = 0; blk < nblk; ++blk) {
      const size_t p0 = blk * BLK;
      const int n = static_cast<int>(std::min(BLK, npts - p0));
      collocate_block(p0, n, phi);
      G.assign(static_cast<size_t>(n) * nbf, 0.0);
      for (int f = 0; f < nbf; ++f)
        for (int p = 0; p < n; ++p) {
          const size_t o = static_cast<size_t>(f) * n + p, i = p0 + p;
          G[o] = 0.5 * (phi[0][o] * vrho[i]) + 2.0 * (phi[1][o] * (gx[i] * vsig[i])) + 2.0 * (phi[2][o] * (gy[i] * vsig[i])) +
                 2.0 * (phi[3][o] * (gz[i] * vsig[i]));
        }
      dgemm('T', 'N', nbf, nbf, n, 1.0, phi[0].data(), n, G.data(), n, 1.0, Kl.data(), nbf);
    }
#pragma omp critical
    for (size_t k = 0; k < Kt.size(); ++k) Kt[k] += Kl[k];
  }
  for (int j = 0; j < nbf; ++j)
    for (int i = 0; i < nbf; ++i)
      K[static_cast<size_t>(j) * nbf + i] = Kt[static_cast<size_t>(j) * nbf + i] + Kt[static_cast<size_t>(i) * nbf + j];
  return 0;
}
}

// ---------------------------------------------------------------------------------------------------------
// VV10 nuclear gradient (post-SCF, restricted): nonlocal_correlation.cpp:108-190, 193-328, 392-522
// ---------------------------------------------------------------------------------------------------------
extern "C" int orc_vv10_gradient(const orc_basis_t *basis, size_t npts, const double *pts, const double *w, int natoms,
                                 const size_t *atom_offsets, const double *D2, double b, double C, double *grad,
                                 double *energy, double *parts, int nthreads) {
  using namespace orc;
  const Basis B(basis);
  const orc_basis_t &raw = B.raw;
  const int nbf = B.nbf;
  if (!raw.shell_atom) return 1;
  // vv10_kernel_gradient is the same arithmetic as vv10_kernel (:113-190 vs :38-106): take rho, grad rho, F, U, W from it
  std::vector<double> Kdummy(static_cast<size_t>(nbf) * nbf), rho4(4 * npts), fuw(3 * npts);
  if (orc_vv10(basis, npts, pts, w, D2, b, C, Kdummy.data(), energy, rho4.data(), fuw.data(), nthreads)) return 1;
  const double kappa_factor = b * 1.5 * kPi * std::pow(9.0 * kPi, -1.0 / 6.0);
  const double beta = std::pow(3.0 / (b * b), 0.75) / 32.0;
  const double pi43 = 4.0 * kPi / 3.0;
  std::vector<double> c0(npts), cx(npts), cy(npts), cz(npts), om(npts), kap(npts);
  for (size_t i = 0; i < npts; ++i) {
    const double rho = rho4[4 * i], gx = rho4[4 * i + 1], gy = rho4[4 * i + 2], gz = rho4[4 * i + 3];
    const double g = gx * gx + gy * gy + gz * gz;
    double t = g / (rho * rho);
    t = C * t * t;
    om[i] = std::sqrt(t + pi43 * rho);
    kap[i] = kappa_factor * std::pow(rho, 1.0 / 6.0);
    const double F = fuw[i], U = fuw[npts + i], W = fuw[2 * npts + i];
    double vrho = beta * w[i], vsig = 0.0;
    if (rho >= 1e-8) {
      vrho = (beta + F + 1.5 * (U * (kap[i] / 6.0) + W * ((0.5 * pi43 * rho - 2.0 * t) / om[i]))) * w[i];
      vsig = W == 0.0 ? 0.0 : 1.5 * W * (t * rho / (g * om[i])) * w[i];
    }
    c0[i] = 0.5 * vrho; // :446 "vrho = vxc.col(0) * 0.5"
    cx[i] = 2.0 * vsig * gx;
    cy[i] = 2.0 * vsig * gy;
    cz[i] = 2.0 * vsig * gz;
  }
  // ---- potential term: V_d = phi_d^T aow + (phi^T aow_d)^T (:462-496), then the distribution loop (:499-512) ----
  const size_t n2 = static_cast<size_t>(nbf) * nbf;
  std::vector<double> V[3];
  for (int d = 0; d < 3; ++d) V[d].assign(n2, 0.0);
  constexpr size_t BLK = 128;
  const size_t nblk = (npts + BLK - 1) / BLK;
  static const int hess_slot[3][3] = {{4, 5, 6}, {5, 7, 8}, {6, 8, 9}}; // xx xy xz yy yz zz
#pragma omp parallel
  {
    std::vector<double> phi[10], tmp[10], aow, Vl[3];
    for (int d = 0; d < 3; ++d) Vl[d].assign(n2, 0.0);
#pragma omp for schedule(dynamic)
    for (size_t blk = 0; blk < nblk; ++blk) {
      const size_t p0 = blk * BLK;
      const int n = static_cast<int>(std::min(BLK, npts - p0));
      for (int c = 0; c < 10; ++c) phi[c].assign(static_cast<size_t>(n) * nbf, 0.0);
      for (int s = 0; s < raw.nsh; ++s) {
        const int l = raw.l[s], sz = orc_nbf_shell(l, raw.pure);
        double *out[10];
        for (int c = 0; c < 10; ++c) {
          tmp[c].assign(static_cast<size_t>(sz) * n, 0.0);
          out[c] = tmp[c].data();
        }
        collocate_shell_deriv2(l, raw.pure, n, pts + 3 * p0, raw.nprim[s], raw.coefs + raw.prim_off[s],
                               raw.exps + raw.prim_off[s], raw.origin + 3 * s, out);
        const double fac = orc_common_fac(l, raw.pure);
        for (int c = 0; c < 10; ++c)
          for (int f = 0; f < sz; ++f)
            for (int p = 0; p < n; ++p)
              phi[c][static_cast<size_t>(B.first_bf[s] + f) * n + p] = fac * tmp[c][static_cast<size_t>(f) * n + p];
      }
      aow.assign(static_cast<size_t>(n) * nbf, 0.0);
      auto fill = [&](int a, int bx, int by, int bz) { // aow = c0 phi[a] + cx phi[bx] + cy phi[by] + cz phi[bz]
        for (int f = 0; f < nbf; ++f)
          for (int p = 0; p < n; ++p) {
            const size_t o = static_cast<size_t>(f) * n + p, i = p0 + p;
            aow[o] = phi[a][o] * c0[i] + phi[bx][o] * cx[i] + phi[by][o] * cy[i] + phi[bz][o] * cz[i];
          }
      };
      fill(0, 1, 2, 3);
      for (int d = 0; d < 3; ++d)
        dgemm('T', 'N', nbf, nbf, n, 1.0, phi[1 + d].data(), n, aow.data(), n, 1.0, Vl[d].data(), nbf);
      for (int d = 0; d < 3; ++d) {
        fill(1 + d, hess_slot[d][0], hess_slot[d][1], hess_slot[d][2]);
        dgemm('T', 'N', nbf, nbf, n, 1.0, aow.data(), n, phi[0].data(), n, 1.0, Vl[d].data(), nbf); // (phi^T aow_d)^T
      }
    }
#pragma omp critical
    for (int d = 0; d < 3; ++d)
      for (size_t k = 0; k < n2; ++k) V[d][k] += Vl[d][k];
  }
  std::vector<int> bf_atom(nbf);
  for (int s = 0; s < raw.nsh; ++s)
    for (int f = 0; f < orc_nbf_shell(raw.l[s], raw.pure); ++f) bf_atom[B.first_bf[s] + f] = raw.shell_atom[s];
  std::vector<double> gpot(3 * static_cast<size_t>(natoms), 0.0), ggrid(3 * static_cast<size_t>(natoms), 0.0);
  for (int mu = 0; mu < nbf; ++mu)
    for (int nu = 0; nu < nbf; ++nu) {
      const double Dval = D2[static_cast<size_t>(nu) * nbf + mu]; // 2.0 * mo.D(mu, nu)
      if (std::fabs(Dval) < 1e-12) continue;
      for (int d = 0; d < 3; ++d) gpot[3 * bf_atom[mu] + d] -= V[d][static_cast<size_t>(nu) * nbf + mu] * Dval;
    }
  // ---- grid-displacement term (:196-328) ----
  std::vector<int> pt_atom(npts, 0);
  for (int a = 0; a < natoms; ++a)
    for (size_t i = atom_offsets[a]; i < atom_offsets[a + 1]; ++i) pt_atom[i] = a;
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
  {
    std::vector<double> gl(3 * static_cast<size_t>(natoms), 0.0);
#pragma omp for schedule(dynamic, 64)
    for (size_t i = 0; i < npts; ++i) {
      const double rho_i = rho4[4 * i];
      if (rho_i < 1e-8) continue;
      for (size_t j = 0; j < npts; ++j) {
        const double rho_j = rho4[4 * j];
        if (i == j || rho_j < 1e-8) continue;
        const double dx = pts[3 * i] - pts[3 * j], dy = pts[3 * i + 1] - pts[3 * j + 1], dz = pts[3 * i + 2] - pts[3 * j + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (std::sqrt(r2) > 15.0) continue;
        const double gi = r2 * om[i] + kap[i], gj = r2 * om[j] + kap[j];
        const double Theta = 2.0 * rho_i * w[i] * rho_j * w[j];
        const double denom = gi * gj * (gi + gj);
        if (std::fabs(denom) < 1e-20) continue;
        const double dD = 2.0 * om[i] * (gj * (gi + gj) + gi * gj) + 2.0 * om[j] * (gi * (gi + gj) + gi * gj);
        const double dPhi = 1.5 * Theta * dD / (denom * denom);
        const double f[3] = {2.0 * dPhi * dx, 2.0 * dPhi * dy, 2.0 * dPhi * dz};
        for (int d = 0; d < 3; ++d) {
          gl[3 * pt_atom[i] + d] -= f[d];
          gl[3 * pt_atom[j] + d] += f[d];
        }
      }
    }
#pragma omp critical
    for (size_t k = 0; k < gl.size(); ++k) ggrid[k] += gl[k];
  }
  for (size_t k = 0; k < gpot.size(); ++k) grad[k] = gpot[k] + ggrid[k];
  if (parts) {
    std::copy(gpot.begin(), gpot.end(), parts);
    std::copy(ggrid.begin(), ggrid.end(), parts + gpot.size());
  }
  return 0;
}
