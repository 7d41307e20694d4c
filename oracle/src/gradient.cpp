// oracle/src/gradient.cpp — CPU ORACLE (test infrastructure): the exchange-correlation NUCLEAR GRADIENT,
// SURVEY.md 8(f) rank 1.  Restates (paths under /root/reference):
//   include/occ/dft/dft.h:598-749              DFT::compute_xc_gradient_impl, screened Morton-leaf path (:666-739)
//   src/dft/dft_gradient_kernels.cpp:27-470    accumulate_gradient_contribution_{R,U}<0|1|2>
//   src/dft/dft_gradient_kernels.cpp:472-595   process_grid_block_gradient (density skip, functional, distribution)
//   src/3rdparty/gau2grid/gau2grid_deriv2.c    second AO derivatives (restated generically, pinned against the
//                                              compiled original in tests/test_oracle_gradient.py)
// The V_x/V_y/V_z matrices are formed explicitly and contracted with D exactly as the reference does, including
// its |D| < 1e-12 skip in the distribution loops (:547-549, :583-585).
#include "oracle.h"
#include "oracle_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <omp.h>
#include <vector>

namespace orc {

static inline double ipow(double x, int n) {
  double r = 1.0;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// value, gradient and the six second derivatives (xx, xy, xz, yy, yz, zz) of one shell; component-major, no common_fac
void collocate_shell_deriv2(int l, int pure, size_t npts, const double *xyz, int nprim, const double *coefs,
                            const double *exps, const double *center, double *out[10]) {
  std::vector<std::array<int, 3>> pw;
  cart_powers(l, pw);
  const int nc = ncart(l);
  const int nout = pure ? nsph(l) : nc;
  const std::vector<double> *c2s = (pure && l >= 1) ? &c2s_matrix(l) : nullptr;
  std::vector<double> c(static_cast<size_t>(10) * nc);
  for (size_t p = 0; p < npts; ++p) {
    const double x = xyz[3 * p] - center[0], y = xyz[3 * p + 1] - center[1], z = xyz[3 * p + 2] - center[2];
    const double r2 = x * x + y * y + z * z;
    double S0 = 0, S1 = 0, S2 = 0;
    for (int n = 0; n < nprim; ++n) {
      const double T1 = coefs[n] * std::exp(-exps[n] * r2);
      const double a2 = -2.0 * exps[n];
      S0 += T1;
      S1 += a2 * T1;
      S2 += a2 * a2 * T1;
    }
    const double q[3] = {x, y, z};
    for (int i = 0; i < nc; ++i) {
      const int e[3] = {pw[i][0], pw[i][1], pw[i][2]};
      auto mono = [&](int dx, int dy, int dz) { // d^(dx+dy+dz) A / dx^dx dy^dy dz^dz for A = x^lx y^ly z^lz
        const int d[3] = {dx, dy, dz};
        double v = 1.0;
        for (int k = 0; k < 3; ++k) {
          int pwr = e[k];
          double f = 1.0;
          for (int t = 0; t < d[k]; ++t) {
            f *= pwr;
            --pwr;
          }
          if (pwr < 0) return 0.0;
          v *= f * ipow(q[k], pwr);
        }
        return v;
      };
      const double A = mono(0, 0, 0);
      const double A1[3] = {mono(1, 0, 0), mono(0, 1, 0), mono(0, 0, 1)};
      c[0 * nc + i] = S0 * A;
      for (int k = 0; k < 3; ++k) c[(1 + k) * nc + i] = S1 * q[k] * A + S0 * A1[k];
      int slot = 4;
      for (int a = 0; a < 3; ++a)
        for (int b = a; b < 3; ++b) {
          int d[3] = {0, 0, 0};
          d[a]++;
          d[b]++;
          const double A2 = mono(d[0], d[1], d[2]);
          double v = S2 * q[a] * q[b] * A + S1 * (q[a] * A1[b] + q[b] * A1[a]) + S0 * A2;
          if (a == b) v += S1 * A;
          c[slot * nc + i] = v;
          ++slot;
        }
    }
    for (int comp = 0; comp < 10; ++comp) {
      if (!out[comp]) continue;
      for (int m = 0; m < nout; ++m) {
        double v;
        if (!c2s) {
          v = c[comp * nc + m];
        } else {
          v = 0.0;
          const double *row = c2s->data() + static_cast<size_t>(m) * nc;
          for (int i = 0; i < nc; ++i) v += row[i] * c[comp * nc + i];
        }
        out[comp][static_cast<size_t>(m) * npts + p] = v;
      }
    }
  }
}

namespace {

struct GradWS {
  std::vector<double> phi[10], Dloc, Dphi, rho, aow, V[2][3], tmp, sym;
  std::vector<double> p_rho, p_sigma, p_tau, exc, vrho, vsigma, vtau;
  std::vector<int> sig, bf_atom;
  std::vector<double> grad;
};

void rowdot(int n, int nb, const double *A, const double *B, double fac, double *out, bool accumulate) {
  std::vector<double> acc(n, 0.0);
  for (int j = 0; j < nb; ++j) {
    const double *a = A + static_cast<size_t>(j) * n, *b = B + static_cast<size_t>(j) * n;
    for (int i = 0; i < n; ++i) acc[i] += a[i] * b[i];
  }
  for (int i = 0; i < n; ++i) out[i] = (accumulate ? out[i] : 0.0) + fac * acc[i];
}
// out (+)= coef[i] * phi[i, j]
void axpy_rows(int n, int nb, const double *phi, const double *coef, double *out, bool accumulate) {
  for (int j = 0; j < nb; ++j) {
    const double *a = phi + static_cast<size_t>(j) * n;
    double *o = out + static_cast<size_t>(j) * n;
    for (int i = 0; i < n; ++i) o[i] = (accumulate ? o[i] : 0.0) + coef[i] * a[i];
  }
}

void gradient_leaf(const Basis &B, const double *extents, const double *spts, const double *sw, const orc_leaf_t &leaf,
                   int nfunc, const int *func_ids, int nspin, int deriv, double density_threshold, bool screen,
                   const double *D2, GradWS &ws) {
  const orc_basis_t &b = B.raw;
  const int nbf = B.nbf;
  const int n = static_cast<int>(leaf.count);
  if (n == 0) return;
  const double *pts = spts + 3 * leaf.offset;
  const double *w = sw + leaf.offset;
  ws.sig.clear();
  for (int s = 0; s < b.nsh; ++s) {
    bool keep = true;
    if (screen) {
      const double dx = b.origin[3 * s] - leaf.center[0], dy = b.origin[3 * s + 1] - leaf.center[1],
                   dz = b.origin[3 * s + 2] - leaf.center[2];
      keep = std::sqrt(dx * dx + dy * dy + dz * dz) - extents[s] < leaf.radius;
    }
    if (keep) ws.sig.push_back(s);
  }
  if (ws.sig.empty()) return;
  int nb = 0;
  ws.bf_atom.clear();
  for (int s : ws.sig) {
    const int sz = orc_nbf_shell(b.l[s], b.pure);
    for (int k = 0; k < sz; ++k) ws.bf_atom.push_back(b.shell_atom[s]); // dft.h:702-707
    nb += sz;
  }
  const int ldD = nspin * nbf, ldl = nspin * nb;
  ws.Dloc.resize(static_cast<size_t>(nspin) * nb * nb);
  {
    int lj = 0;
    for (int sj : ws.sig) {
      const int gj = B.first_bf[sj], nj = orc_nbf_shell(b.l[sj], b.pure);
      for (int jj = 0; jj < nj; ++jj)
        for (int sp = 0; sp < nspin; ++sp) {
          int li = 0;
          for (int si : ws.sig) {
            const int gi = B.first_bf[si], ni = orc_nbf_shell(b.l[si], b.pure);
            std::memcpy(&ws.Dloc[static_cast<size_t>(lj + jj) * ldl + sp * nb + li],
                        &D2[static_cast<size_t>(gj + jj) * ldD + sp * nbf + gi], sizeof(double) * ni);
            li += ni;
          }
        }
      lj += nj;
    }
  }
  // ---- AO values: one more derivative than the functional needs (dft.h:626-628) ----
  const int max_deriv = deriv == 0 ? 1 : 2;
  const int nphi = max_deriv == 1 ? 4 : 10;
  for (int c = 0; c < nphi; ++c) ws.phi[c].resize(static_cast<size_t>(n) * nb);
  {
    int bf = 0;
    for (int s : ws.sig) {
      const int L = b.l[s], sz = orc_nbf_shell(L, b.pure);
      const double fac = orc_common_fac(L, b.pure);
      double *out[10];
      for (int c = 0; c < 10; ++c) out[c] = c < nphi ? ws.phi[c].data() + static_cast<size_t>(bf) * n : nullptr;
      collocate_shell_deriv2(L, b.pure, n, pts, b.nprim[s], b.coefs + b.prim_off[s], b.exps + b.prim_off[s],
                             b.origin + 3 * s, out);
      const size_t cnt = static_cast<size_t>(sz) * n;
      for (int c = 0; c < nphi; ++c)
        for (size_t i = 0; i < cnt; ++i) out[c][i] *= fac;
      bf += sz;
    }
  }
  // ---- density (density.cpp:8-166; Laplacian column not needed) ----
  const int ncomp = deriv == 0 ? 1 : (deriv == 1 ? 4 : 6);
  const int ldr = nspin * n;
  ws.rho.assign(static_cast<size_t>(ldr) * ncomp, 0.0);
  ws.Dphi.resize(static_cast<size_t>(n) * nb);
  ws.tmp.resize(static_cast<size_t>(nb) * nb);
  for (int sp = 0; sp < nspin; ++sp) {
    const double *Ds = ws.Dloc.data() + sp * nb;
    const double *Dmat = Ds;
    int ldm = ldl;
    if (nspin == 1) {
      for (int j = 0; j < nb; ++j)
        for (int i = 0; i < nb; ++i)
          ws.tmp[static_cast<size_t>(j) * nb + i] =
              (i <= j) ? Ds[static_cast<size_t>(j) * ldl + i] : Ds[static_cast<size_t>(i) * ldl + j];
      Dmat = ws.tmp.data();
      ldm = nb;
    }
    dgemm('N', 'N', n, nb, nb, 1.0, ws.phi[0].data(), n, Dmat, ldm, 0.0, ws.Dphi.data(), n);
    double *r = ws.rho.data() + sp * n;
    rowdot(n, nb, ws.phi[0].data(), ws.Dphi.data(), 1.0, r, false);
    if (deriv >= 1)
      for (int d = 0; d < 3; ++d)
        rowdot(n, nb, ws.phi[1 + d].data(), ws.Dphi.data(), 2.0, r + static_cast<size_t>(1 + d) * ldr, false);
    if (deriv >= 2) {
      double *tau = r + static_cast<size_t>(5) * ldr;
      for (int d = 0; d < 3; ++d) {
        dgemm('N', 'N', n, nb, nb, 1.0, ws.phi[1 + d].data(), n, Ds, ldl, 0.0, ws.Dphi.data(), n);
        rowdot(n, nb, ws.phi[1 + d].data(), ws.Dphi.data(), 1.0, tau, d > 0);
      }
      for (int i = 0; i < n; ++i) tau[i] *= 0.5;
    }
  }
  // ---- dft_gradient_kernels.cpp:490-493: skip (no electron count in the gradient path) ----
  double mx = ws.rho[0];
  for (int i = 0; i < nspin * n; ++i) mx = std::max(mx, ws.rho[i]);
  if (mx < density_threshold) return;
  // ---- params, functional, weighting (:495-515) ----
  const int ns_sigma = nspin == 1 ? 1 : 3;
  ws.p_rho.resize(static_cast<size_t>(n) * nspin);
  ws.p_sigma.assign(static_cast<size_t>(n) * ns_sigma, 0.0);
  ws.p_tau.assign(static_cast<size_t>(n) * nspin, 0.0);
  auto g = [&](int sp, int d, int i) { return ws.rho[static_cast<size_t>(1 + d) * ldr + sp * n + i]; };
  for (int i = 0; i < n; ++i) {
    for (int sp = 0; sp < nspin; ++sp) ws.p_rho[static_cast<size_t>(i) * nspin + sp] = ws.rho[sp * n + i];
    if (deriv >= 1) {
      if (nspin == 1) {
        ws.p_sigma[i] = g(0, 0, i) * g(0, 0, i) + g(0, 1, i) * g(0, 1, i) + g(0, 2, i) * g(0, 2, i);
      } else {
        ws.p_sigma[3 * i + 0] = g(0, 0, i) * g(0, 0, i) + g(0, 1, i) * g(0, 1, i) + g(0, 2, i) * g(0, 2, i);
        ws.p_sigma[3 * i + 1] = g(0, 0, i) * g(1, 0, i) + g(0, 1, i) * g(1, 1, i) + g(0, 2, i) * g(1, 2, i);
        ws.p_sigma[3 * i + 2] = g(1, 0, i) * g(1, 0, i) + g(1, 1, i) * g(1, 1, i) + g(1, 2, i) * g(1, 2, i);
      }
    }
    if (deriv >= 2)
      for (int sp = 0; sp < nspin; ++sp)
        ws.p_tau[static_cast<size_t>(i) * nspin + sp] = ws.rho[static_cast<size_t>(5) * ldr + sp * n + i];
  }
  ws.exc.assign(n, 0.0);
  ws.vrho.assign(static_cast<size_t>(n) * nspin, 0.0);
  ws.vsigma.assign(static_cast<size_t>(n) * ns_sigma, 0.0);
  ws.vtau.assign(static_cast<size_t>(n) * nspin, 0.0);
  for (int f = 0; f < nfunc; ++f)
    xc_eval_add(func_ids[f], nspin == 2, n, ws.p_rho.data(), ws.p_sigma.data(), ws.p_tau.data(), ws.exc.data(),
                ws.vrho.data(), ws.vsigma.data(), ws.vtau.data());
  for (int i = 0; i < n; ++i) {
    for (int sp = 0; sp < nspin; ++sp) {
      ws.vrho[static_cast<size_t>(i) * nspin + sp] *= w[i];
      ws.vtau[static_cast<size_t>(i) * nspin + sp] *= w[i];
    }
    for (int k = 0; k < ns_sigma; ++k) ws.vsigma[static_cast<size_t>(i) * ns_sigma + k] *= w[i];
  }
  // ---- V_x, V_y, V_z per spin (accumulate_gradient_contribution_{R,U}<deriv>) ----
  // second-derivative slot of the (a, b) pair in phi[4..9] = xx, xy, xz, yy, yz, zz
  static const int H[3][3] = {{4, 5, 6}, {5, 7, 8}, {6, 8, 9}};
  ws.aow.resize(static_cast<size_t>(n) * nb);
  ws.sym.resize(static_cast<size_t>(nb) * nb);
  std::vector<double> c0(n), cd[3], ct(n);
  for (auto &v : cd) v.resize(n);
  for (int sp = 0; sp < nspin; ++sp) {
    for (int d = 0; d < 3; ++d) ws.V[sp][d].assign(static_cast<size_t>(nb) * nb, 0.0);
    if (deriv == 0) {
      // :27-45 / :179-212   V_d = phi_d^T (vrho phi)
      for (int i = 0; i < n; ++i) c0[i] = ws.vrho[static_cast<size_t>(i) * nspin + sp];
      axpy_rows(n, nb, ws.phi[0].data(), c0.data(), ws.aow.data(), false);
      for (int d = 0; d < 3; ++d)
        dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[1 + d].data(), n, ws.aow.data(), n, 1.0, ws.V[sp][d].data(), nb);
      continue;
    }
    const int other = 1 - sp, same = sp == 0 ? 0 : 2;
    for (int i = 0; i < n; ++i) {
      c0[i] = 0.5 * ws.vrho[static_cast<size_t>(i) * nspin + sp];
      for (int d = 0; d < 3; ++d)
        cd[d][i] = nspin == 1 ? 2.0 * ws.vsigma[i] * g(0, d, i)
                              : 2.0 * ws.vsigma[3 * i + same] * g(sp, d, i) + ws.vsigma[3 * i + 1] * g(other, d, i);
      ct[i] = deriv >= 2 ? 0.5 * ws.vtau[static_cast<size_t>(i) * nspin + sp] : 0.0;
    }
    // aow = vrho/2 phi + sum_d' c_d' phi_d' ;  V_d += phi_d^T aow
    axpy_rows(n, nb, ws.phi[0].data(), c0.data(), ws.aow.data(), false);
    for (int d = 0; d < 3; ++d) axpy_rows(n, nb, ws.phi[1 + d].data(), cd[d].data(), ws.aow.data(), true);
    for (int d = 0; d < 3; ++d)
      dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[1 + d].data(), n, ws.aow.data(), n, 1.0, ws.V[sp][d].data(), nb);
    // aow_d = vrho/2 phi_d + sum_d' c_d' phi_dd' ;  V_d += (phi^T aow_d)^T
    for (int d = 0; d < 3; ++d) {
      axpy_rows(n, nb, ws.phi[1 + d].data(), c0.data(), ws.aow.data(), false);
      for (int e = 0; e < 3; ++e) axpy_rows(n, nb, ws.phi[H[d][e]].data(), cd[e].data(), ws.aow.data(), true);
      dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[0].data(), n, ws.aow.data(), n, 0.0, ws.sym.data(), nb);
      double *V = ws.V[sp][d].data();
      for (int j = 0; j < nb; ++j)
        for (int i = 0; i < nb; ++i) V[static_cast<size_t>(j) * nb + i] += ws.sym[static_cast<size_t>(i) * nb + j];
    }
    if (deriv >= 2) {
      // tau terms (:141-157 / :386-415): V_d += sum_e phi_{d e}^T (vtau/2 phi_e), bra only
      for (int e = 0; e < 3; ++e) {
        axpy_rows(n, nb, ws.phi[1 + e].data(), ct.data(), ws.aow.data(), false);
        for (int d = 0; d < 3; ++d)
          dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[H[d][e]].data(), n, ws.aow.data(), n, 1.0, ws.V[sp][d].data(), nb);
      }
    }
  }
  // ---- distribution to atoms (:540-553 restricted, :575-592 unrestricted) ----
  if (nspin == 1) {
    for (int mu = 0; mu < nb; ++mu) {
      const int atom = ws.bf_atom[mu];
      for (int nu = 0; nu < nb; ++nu) {
        const double Dval = 2.0 * ws.Dloc[static_cast<size_t>(nu) * ldl + mu];
        if (std::fabs(Dval) < 1e-12) continue;
        for (int d = 0; d < 3; ++d) ws.grad[3 * atom + d] -= ws.V[0][d][static_cast<size_t>(nu) * nb + mu] * Dval;
      }
    }
  } else {
    for (int mu = 0; mu < nb; ++mu) {
      const int atom = ws.bf_atom[mu];
      for (int nu = 0; nu < nb; ++nu) {
        const double Da = ws.Dloc[static_cast<size_t>(nu) * ldl + mu], Db = ws.Dloc[static_cast<size_t>(nu) * ldl + nb + mu];
        if (std::fabs(Da + Db) < 1e-12) continue;
        for (int d = 0; d < 3; ++d)
          ws.grad[3 * atom + d] -= 2.0 * (ws.V[0][d][static_cast<size_t>(nu) * nb + mu] * Da +
                                          ws.V[1][d][static_cast<size_t>(nu) * nb + mu] * Db);
      }
    }
  }
}

} // namespace
} // namespace orc

extern "C" {

void orc_collocate_shell_deriv2(int l, int pure, size_t npts, const double *xyz, int nprim, const double *coefs,
                                const double *exps, const double *center, double *out10 /* 10 x nout x npts */) {
  const int nout = orc_nbf_shell(l, pure);
  double *out[10];
  for (int c = 0; c < 10; ++c) out[c] = out10 + static_cast<size_t>(c) * nout * npts;
  orc::collocate_shell_deriv2(l, pure, npts, xyz, nprim, coefs, exps, center, out);
}

int orc_xc_gradient(const orc_basis_t *basis, size_t npts, const double *spts, const double *sw, size_t nleaves,
                    const orc_leaf_t *leaves, int nfunc, const int *func_ids, int nspin, int deriv,
                    double density_threshold, double screening_threshold, const double *D2, int natoms, double *grad,
                    int nthreads) {
  using namespace orc;
  (void)npts;
  for (int f = 0; f < nfunc; ++f)
    if (orc_xc_family(func_ids[f]) < 0) return 1;
  if (!basis->shell_atom) return 2;
  if (nthreads < 1) nthreads = 1;
  Basis B(basis);
  const bool screen = screening_threshold > 0.0;
  std::vector<double> extents(basis->nsh, 0.0);
  if (screen) orc_decay_cutoff(basis, screening_threshold, extents.data());
  for (int l = 0; l <= ORC_MAX_L; ++l)
    if (basis->pure) c2s_matrix(l);
  std::vector<GradWS> wss(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    GradWS &ws = wss[omp_get_thread_num()];
    ws.grad.assign(3 * static_cast<size_t>(natoms), 0.0);
#pragma omp for schedule(dynamic, 1)
    for (long li = 0; li < static_cast<long>(nleaves); ++li)
      gradient_leaf(B, extents.data(), spts, sw, leaves[li], nfunc, func_ids, nspin, deriv, density_threshold, screen, D2,
                    ws);
  }
  std::fill(grad, grad + 3 * static_cast<size_t>(natoms), 0.0);
  for (auto &ws : wss)
    for (size_t i = 0; i < ws.grad.size(); ++i) grad[i] += ws.grad[i];
  return 0;
}

} // extern "C"
