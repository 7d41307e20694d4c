// oracle/src/vv10.cpp — CPU ORACLE (test infrastructure): the post-SCF VV10 non-local correlation of the reference,
// SURVEY.md 8(f) rank 4.  Restates (paths under /root/reference):
//   src/dft/nonlocal_correlation.cpp:38-118     vv10_kernel  (the pair sum runs over j < i ONLY, in the order the
//                                               points are handed over; self term seeded into f and u; points with
//                                               rho < 1e-8 neither receive nor contribute)
//   src/dft/nonlocal_correlation.cpp:367-420    NonLocalCorrelationFunctional::vv10  (dense evaluate_basis over all
//                                               points, restricted deriv-1 density, E = sum exc . rho,
//                                               Vxc = ktmp + ktmp^T with ktmp = phi^T (1/2 vrho phi + 2 vsigma grad rho . grad phi))
// "parity unpinned": the reference's only VV10 test is a 1e-3-level wB97X-V gradient comparison
// (tests/dft_gradient_tests.cpp:449-520); this restatement is checked against an independent numpy
// transcription of the same lines in tests/test_vv10.py.
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
