// oracle/src/build.cpp — CPU ORACLE (test infrastructure): the per-SCF-iteration Vxc/Exc build.
// Restates (paths under /root/reference):
//   include/occ/dft/dft.h:190-388              DFT::compute_K_dft, screened Morton-leaf path (:288-372)
//   src/dft/dft_kernels.cpp:8-70               kernels::process_grid_block
//   include/occ/dft/dft_kernels.h:15-59        impl::set_params (sigma from grad rho)
//   src/gto/gto.cpp:212-287                    evaluate_basis_subset (collocation + common_fac)
//   src/gto/density.cpp:8-166                  evaluate_density<deriv, R|U>
//   include/occ/dft/functional.h:438-470       Result::operator+= / weight_by
//   src/dft/xc_potential_matrix.cpp:8-135      xc_potential_matrix<R|U, deriv>
// OpenMP over leaves stands in for tbb::parallel_for (dft.h:303); thread-local K copies and the final
// serial reduction mirror dft.h:221-232,375-380.  Eigen's GEMM is replaced by OpenBLAS (when bound
// through orc_set_blas) or a simple blocked kernel: summation order is the only difference.
// Deviation (documented in DESIGN.md): the Laplacian column rho(:,4) (density.cpp:84-100) is not formed —
// the reference never consumes it (xc_potential_matrix.cpp:42-57 ignores vlaplacian; r2SCAN has no
// Laplacian dependence) — so AO second derivatives are not evaluated.
#include "oracle.h"
#include "oracle_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <omp.h>
#include <vector>

namespace orc {

typedef void (*cblas_dgemm_t)(int order, int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda,
                              const double *B, int ldb, double beta, double *C, int ldc);
static cblas_dgemm_t g_cblas_dgemm = nullptr;

static void dgemm_plain(char ta, char tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B,
                        int ldb, double beta, double *C, int ldc) {
  for (int j = 0; j < n; ++j) {
    double *c = C + static_cast<size_t>(j) * ldc;
    if (beta == 0.0)
      for (int i = 0; i < m; ++i) c[i] = 0.0;
    else if (beta != 1.0)
      for (int i = 0; i < m; ++i) c[i] *= beta;
  }
  const bool tA = (ta == 'T'), tB = (tb == 'T');
  if (!tA) {
    const int KB = 64;
    for (int k0 = 0; k0 < k; k0 += KB) {
      const int k1 = std::min(k, k0 + KB);
      for (int j = 0; j < n; ++j) {
        double *c = C + static_cast<size_t>(j) * ldc;
        for (int p = k0; p < k1; ++p) {
          const double b = alpha * (tB ? B[j + static_cast<size_t>(p) * ldb] : B[p + static_cast<size_t>(j) * ldb]);
          const double *a = A + static_cast<size_t>(p) * lda;
          for (int i = 0; i < m; ++i) c[i] += a[i] * b;
        }
      }
    }
  } else {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        const double *a = A + static_cast<size_t>(i) * lda;
        double s = 0.0;
        if (!tB) {
          const double *b = B + static_cast<size_t>(j) * ldb;
          for (int p = 0; p < k; ++p) s += a[p] * b[p];
        } else {
          for (int p = 0; p < k; ++p) s += a[p] * B[j + static_cast<size_t>(p) * ldb];
        }
        C[i + static_cast<size_t>(j) * ldc] += alpha * s;
      }
  }
}

void dgemm(char ta, char tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
           double beta, double *C, int ldc) {
  if (m == 0 || n == 0) return;
  if (g_cblas_dgemm) {
    g_cblas_dgemm(102 /*ColMajor*/, ta == 'T' ? 112 : 111, tb == 'T' ? 112 : 111, m, n, k, alpha, A, lda, B, ldb, beta,
                  C, ldc);
    return;
  }
  dgemm_plain(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

struct Workspace {
  std::vector<double> phi[4], Dloc, Dphi, rho, tmp, Kblk, aow;
  std::vector<double> p_rho, p_sigma, p_tau, exc, vrho, vsigma, vtau;
  std::vector<int> sig;
  std::vector<double> K;
  double energy = 0, na = 0, nbeta = 0;
  double t_colloc = 0, t_density = 0, t_func = 0, t_vxc = 0;
  double sum_n_nb2 = 0, sum_nb2 = 0;
  size_t nskipped = 0;
  int nb_min = 1 << 30, nb_max = 0;
};

// rowsum(A o B) for column-major (n x nb) matrices, accumulated into out[n] with factor
static void rowdot(int n, int nb, const double *A, const double *B, double fac, double *out, bool accumulate) {
  if (!accumulate)
    for (int i = 0; i < n; ++i) out[i] = 0.0;
  std::vector<double> acc(n, 0.0);
  for (int j = 0; j < nb; ++j) {
    const double *a = A + static_cast<size_t>(j) * n, *b = B + static_cast<size_t>(j) * n;
    for (int i = 0; i < n; ++i) acc[i] += a[i] * b[i];
  }
  for (int i = 0; i < n; ++i) out[i] += fac * acc[i];
}

// C(n x nb) = sum_d coef_d[i] * phi_d[i, j]
static void scale_rows_add(int n, int nb, const double *phi, const double *coef, double *out, bool accumulate) {
  for (int j = 0; j < nb; ++j) {
    const double *a = phi + static_cast<size_t>(j) * n;
    double *o = out + static_cast<size_t>(j) * n;
    if (accumulate)
      for (int i = 0; i < n; ++i) o[i] += coef[i] * a[i];
    else
      for (int i = 0; i < n; ++i) o[i] = coef[i] * a[i];
  }
}

static void process_leaf(const Basis &B, const double *extents, const double *spts, const double *sw,
                         const orc_leaf_t &leaf, int nfunc, const int *func_ids, int nspin, int deriv,
                         double density_threshold, bool screen, const double *D2, Workspace &ws) {
  const orc_basis_t &b = B.raw;
  const int nbf = B.nbf;
  const int n = static_cast<int>(leaf.count);
  if (n == 0) return;
  const double *pts = spts + 3 * leaf.offset;
  const double *w = sw + leaf.offset;

  // ---- dft.h:313-320 significant shells ----
  ws.sig.clear();
  for (int s = 0; s < b.nsh; ++s) {
    if (!screen) {
      ws.sig.push_back(s);
      continue;
    }
    const double dx = b.origin[3 * s] - leaf.center[0], dy = b.origin[3 * s + 1] - leaf.center[1],
                 dz = b.origin[3 * s + 2] - leaf.center[2];
    const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (dist - extents[s] < leaf.radius) ws.sig.push_back(s);
  }
  if (ws.sig.empty()) return;
  // ---- dft.h:322-335 packed segments ----
  int nb = 0;
  for (int s : ws.sig) nb += orc_nbf_shell(b.l[s], b.pure);
  ws.sum_n_nb2 += static_cast<double>(n) * nb * nb;
  ws.sum_nb2 += static_cast<double>(nb) * nb;
  ws.nb_min = std::min(ws.nb_min, nb);
  ws.nb_max = std::max(ws.nb_max, nb);

  // ---- dft.h:340-348 gather Dloc (nspin*nb x nb) ----
  ws.Dloc.resize(static_cast<size_t>(nspin) * nb * nb);
  const int ldD = nspin * nbf, ldl = nspin * nb;
  {
    int lj = 0;
    for (int sj : ws.sig) {
      const int gj = B.first_bf[sj], nj = orc_nbf_shell(b.l[sj], b.pure);
      for (int jj = 0; jj < nj; ++jj) {
        for (int sp = 0; sp < nspin; ++sp) {
          int li = 0;
          for (int si : ws.sig) {
            const int gi = B.first_bf[si], ni = orc_nbf_shell(b.l[si], b.pure);
            std::memcpy(&ws.Dloc[static_cast<size_t>(lj + jj) * ldl + sp * nb + li],
                        &D2[static_cast<size_t>(gj + jj) * ldD + sp * nbf + gi], sizeof(double) * ni);
            li += ni;
          }
        }
      }
      lj += nj;
    }
  }

  // ---- gto.cpp:212-287 collocation on the shell subset ----
  double t0 = omp_get_wtime();
  const int ncomp_phi = deriv >= 1 ? 4 : 1;
  for (int c = 0; c < ncomp_phi; ++c) ws.phi[c].resize(static_cast<size_t>(n) * nb);
  {
    int bf = 0;
    for (int s : ws.sig) {
      const int L = b.l[s], sz = orc_nbf_shell(L, b.pure);
      const double fac = orc_common_fac(L, b.pure);
      double *o = ws.phi[0].data() + static_cast<size_t>(bf) * n;
      double *ox = deriv >= 1 ? ws.phi[1].data() + static_cast<size_t>(bf) * n : nullptr;
      double *oy = deriv >= 1 ? ws.phi[2].data() + static_cast<size_t>(bf) * n : nullptr;
      double *oz = deriv >= 1 ? ws.phi[3].data() + static_cast<size_t>(bf) * n : nullptr;
      collocate_shell(L, b.pure, n, pts, b.nprim[s], b.coefs + b.prim_off[s], b.exps + b.prim_off[s],
                      b.origin + 3 * s, deriv >= 1 ? 1 : 0, o, ox, oy, oz);
      const size_t cnt = static_cast<size_t>(sz) * n;
      for (size_t i = 0; i < cnt; ++i) o[i] *= fac;
      if (deriv >= 1)
        for (size_t i = 0; i < cnt; ++i) {
          ox[i] *= fac;
          oy[i] *= fac;
          oz[i] *= fac;
        }
      bf += sz;
    }
  }
  double t1 = omp_get_wtime();
  ws.t_colloc += t1 - t0;

  // ---- density.cpp:8-166 ----
  const int ncomp = deriv == 0 ? 1 : (deriv == 1 ? 4 : 6);
  ws.rho.assign(static_cast<size_t>(nspin) * n * ncomp, 0.0); // (nspin*n) x ncomp column-major
  const int ldr = nspin * n;
  ws.Dphi.resize(static_cast<size_t>(n) * nb);
  ws.tmp.resize(static_cast<size_t>(nb) * nb);
  for (int sp = 0; sp < nspin; ++sp) {
    const double *Ds = ws.Dloc.data() + sp * nb; // ld = ldl
    const double *Dmat = Ds;
    int ldm = ldl;
    if (nspin == 1) {
      // restricted: D.selfadjointView<Upper>() — only the upper triangle is read (density.cpp:14,38,78)
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
      // tau = 1/2 sum_d rowsum(phi_d o (phi_d D)) with the FULL D (density.cpp:90-100)
      double *tau = r + static_cast<size_t>(5) * ldr;
      for (int d = 0; d < 3; ++d) {
        dgemm('N', 'N', n, nb, nb, 1.0, ws.phi[1 + d].data(), n, Ds, ldl, 0.0, ws.Dphi.data(), n);
        rowdot(n, nb, ws.phi[1 + d].data(), ws.Dphi.data(), 1.0, tau, d > 0);
      }
      for (int i = 0; i < n; ++i) tau[i] *= 0.5;
    }
  }
  double t2 = omp_get_wtime();
  ws.t_density += t2 - t1;

  // ---- dft_kernels.cpp:27-40 electron counts, then the leaf-level density skip ----
  {
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < n; ++i) sa += ws.rho[i] * w[i];
    ws.na += sa;
    if (nspin == 2) {
      for (int i = 0; i < n; ++i) sb += ws.rho[n + i] * w[i];
      ws.nbeta += sb;
    }
    double mx = ws.rho[0];
    for (int i = 0; i < nspin * n; ++i) mx = std::max(mx, ws.rho[i]);
    if (mx < density_threshold) {
      ws.nskipped++;
      return;
    }
  }

  // ---- dft_kernels.h:15-59 set_params (libxc interleaved layouts) ----
  const int ns_sigma = nspin == 1 ? 1 : 3;
  ws.p_rho.resize(static_cast<size_t>(n) * nspin);
  ws.p_sigma.resize(static_cast<size_t>(n) * ns_sigma);
  ws.p_tau.resize(static_cast<size_t>(n) * nspin);
  for (int i = 0; i < n; ++i) {
    for (int sp = 0; sp < nspin; ++sp) ws.p_rho[static_cast<size_t>(i) * nspin + sp] = ws.rho[sp * n + i];
    if (deriv >= 1) {
      auto g = [&](int sp, int d) { return ws.rho[static_cast<size_t>(1 + d) * ldr + sp * n + i]; };
      if (nspin == 1) {
        ws.p_sigma[i] = g(0, 0) * g(0, 0) + g(0, 1) * g(0, 1) + g(0, 2) * g(0, 2);
      } else {
        ws.p_sigma[3 * i + 0] = g(0, 0) * g(0, 0) + g(0, 1) * g(0, 1) + g(0, 2) * g(0, 2);
        ws.p_sigma[3 * i + 1] = g(0, 0) * g(1, 0) + g(0, 1) * g(1, 1) + g(0, 2) * g(1, 2);
        ws.p_sigma[3 * i + 2] = g(1, 0) * g(1, 0) + g(1, 1) * g(1, 1) + g(1, 2) * g(1, 2);
      }
    }
    if (deriv >= 2)
      for (int sp = 0; sp < nspin; ++sp)
        ws.p_tau[static_cast<size_t>(i) * nspin + sp] = ws.rho[static_cast<size_t>(5) * ldr + sp * n + i];
  }
  // ---- functional.cpp:18-65 summed over components (functional.h:438-459), then weight_by (:460-470) ----
  ws.exc.assign(n, 0.0);
  ws.vrho.assign(static_cast<size_t>(n) * nspin, 0.0);
  ws.vsigma.assign(static_cast<size_t>(n) * ns_sigma, 0.0);
  ws.vtau.assign(static_cast<size_t>(n) * nspin, 0.0);
  for (int f = 0; f < nfunc; ++f)
    xc_eval_add(func_ids[f], nspin == 2, n, ws.p_rho.data(), ws.p_sigma.data(), ws.p_tau.data(), ws.exc.data(),
                ws.vrho.data(), ws.vsigma.data(), ws.vtau.data());
  for (int i = 0; i < n; ++i) {
    ws.exc[i] *= w[i];
    for (int sp = 0; sp < nspin; ++sp) {
      ws.vrho[static_cast<size_t>(i) * nspin + sp] *= w[i];
      ws.vtau[static_cast<size_t>(i) * nspin + sp] *= w[i];
    }
    for (int k = 0; k < ns_sigma; ++k) ws.vsigma[static_cast<size_t>(i) * ns_sigma + k] *= w[i];
  }
  double t3 = omp_get_wtime();
  ws.t_func += t3 - t2;

  // ---- xc_potential_matrix.cpp ----
  {
    double e = 0.0;
    for (int i = 0; i < n; ++i) {
      double rt = ws.rho[i];
      if (nspin == 2) rt += ws.rho[n + i];
      e += rt * ws.exc[i];
    }
    ws.energy += e;
  }
  ws.Kblk.assign(static_cast<size_t>(nspin) * nb * nb, 0.0); // (nspin*nb) x nb, ld = ldl
  ws.aow.resize(static_cast<size_t>(n) * nb);
  std::vector<double> coef(n), ktmp;
  for (int sp = 0; sp < nspin; ++sp) {
    double *Ks = ws.Kblk.data() + sp * nb;
    if (deriv == 0) {
      // :8-16 / :59-71  V = phi^T diag(w vrho) phi
      for (int i = 0; i < n; ++i) coef[i] = ws.vrho[static_cast<size_t>(i) * nspin + sp];
      scale_rows_add(n, nb, ws.phi[0].data(), coef.data(), ws.aow.data(), false);
      dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[0].data(), n, ws.aow.data(), n, 0.0, Ks, ldl);
    } else {
      ktmp.assign(static_cast<size_t>(nb) * nb, 0.0);
      if (nspin == 1) {
        // :18-40  ktmp = phi^T [ 1/2 vrho phi + 2 vsigma (grad rho . grad phi) ],  V = ktmp + ktmp^T
        for (int i = 0; i < n; ++i) coef[i] = 0.5 * ws.vrho[i];
        scale_rows_add(n, nb, ws.phi[0].data(), coef.data(), ws.aow.data(), false);
        for (int d = 0; d < 3; ++d) {
          for (int i = 0; i < n; ++i) coef[i] = 2 * (ws.rho[static_cast<size_t>(1 + d) * ldr + i] * ws.vsigma[i]);
          scale_rows_add(n, nb, ws.phi[1 + d].data(), coef.data(), ws.aow.data(), true);
        }
        dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[0].data(), n, ws.aow.data(), n, 0.0, ktmp.data(), nb);
      } else {
        // :73-107  LDA-like term directly, gradient term symmetrised
        for (int i = 0; i < n; ++i) coef[i] = ws.vrho[static_cast<size_t>(i) * 2 + sp];
        scale_rows_add(n, nb, ws.phi[0].data(), coef.data(), ws.aow.data(), false);
        dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[0].data(), n, ws.aow.data(), n, 0.0, Ks, ldl);
        const int same = sp == 0 ? 0 : 2, other = 1 - sp;
        for (int d = 0; d < 3; ++d) {
          for (int i = 0; i < n; ++i)
            coef[i] = ws.rho[static_cast<size_t>(1 + d) * ldr + sp * n + i] * (2 * ws.vsigma[3 * i + same]) +
                      ws.rho[static_cast<size_t>(1 + d) * ldr + other * n + i] * ws.vsigma[3 * i + 1];
          scale_rows_add(n, nb, ws.phi[1 + d].data(), coef.data(), ws.aow.data(), d > 0);
        }
        dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[0].data(), n, ws.aow.data(), n, 0.0, ktmp.data(), nb);
      }
      for (int j = 0; j < nb; ++j)
        for (int i = 0; i < nb; ++i)
          Ks[static_cast<size_t>(j) * ldl + i] += ktmp[static_cast<size_t>(j) * nb + i] + ktmp[static_cast<size_t>(i) * nb + j];
      if (deriv >= 2) {
        // :42-57 / :109-135  + sum_d phi_d^T diag(1/2 w vtau) phi_d   (vlaplacian is never used)
        for (int i = 0; i < n; ++i) coef[i] = 0.5 * ws.vtau[static_cast<size_t>(i) * nspin + sp];
        for (int d = 0; d < 3; ++d) {
          scale_rows_add(n, nb, ws.phi[1 + d].data(), coef.data(), ws.aow.data(), false);
          dgemm('T', 'N', nb, nb, n, 1.0, ws.phi[1 + d].data(), n, ws.aow.data(), n, 1.0, Ks, ldl);
        }
      }
    }
  }
  double t4 = omp_get_wtime();
  ws.t_vxc += t4 - t3;

  // ---- dft.h:361-369 scatter into the thread-local K ----
  {
    int lj = 0;
    for (int sj : ws.sig) {
      const int gj = B.first_bf[sj], nj = orc_nbf_shell(b.l[sj], b.pure);
      for (int jj = 0; jj < nj; ++jj)
        for (int sp = 0; sp < nspin; ++sp) {
          int li = 0;
          for (int si : ws.sig) {
            const int gi = B.first_bf[si], ni = orc_nbf_shell(b.l[si], b.pure);
            double *dst = &ws.K[static_cast<size_t>(gj + jj) * ldD + sp * nbf + gi];
            const double *src = &ws.Kblk[static_cast<size_t>(lj + jj) * ldl + sp * nb + li];
            for (int ii = 0; ii < ni; ++ii) dst[ii] += src[ii];
            li += ni;
          }
        }
      lj += nj;
    }
  }
}

} // namespace orc

extern "C" {

int orc_set_blas(const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 0;
  void *f = dlsym(h, "scipy_cblas_dgemm");
  if (!f) f = dlsym(h, "cblas_dgemm");
  if (!f) return 0;
  // one OpenBLAS thread per caller: parallelism is over leaves (mirrors tbb::parallel_for, dft.h:303)
  typedef void (*setnt_t)(int);
  setnt_t setnt = reinterpret_cast<setnt_t>(dlsym(h, "scipy_openblas_set_num_threads"));
  if (!setnt) setnt = reinterpret_cast<setnt_t>(dlsym(h, "openblas_set_num_threads"));
  if (setnt) setnt(1);
  orc::g_cblas_dgemm = reinterpret_cast<orc::cblas_dgemm_t>(f);
  return 1;
}

int orc_leaf_shell_list(const orc_basis_t *b, const double *extents, const orc_leaf_t *leaf, int *out) {
  int cnt = 0;
  for (int s = 0; s < b->nsh; ++s) {
    const double dx = b->origin[3 * s] - leaf->center[0], dy = b->origin[3 * s + 1] - leaf->center[1],
                 dz = b->origin[3 * s + 2] - leaf->center[2];
    if (std::sqrt(dx * dx + dy * dy + dz * dz) - extents[s] < leaf->radius) out[cnt++] = s;
  }
  return cnt;
}

int orc_xc_build(const orc_basis_t *basis, size_t npts, const double *spts, const double *sw, size_t nleaves,
                 const orc_leaf_t *leaves, int nfunc, const int *func_ids, int nspin, int deriv,
                 double density_threshold, double screening_threshold, const double *D2, double *K, double *exc,
                 double *n_alpha, double *n_beta, int nthreads, orc_build_stats_t *stats) {
  using namespace orc;
  (void)npts;
  for (int f = 0; f < nfunc; ++f)
    if (orc_xc_family(func_ids[f]) < 0) return 1;
  if (nthreads < 1) nthreads = 1;
  Basis B(basis);
  const size_t ksz = static_cast<size_t>(nspin) * B.nbf * B.nbf;
  const double tstart = omp_get_wtime();
  const bool screen = screening_threshold > 0.0;
  std::vector<double> extents(basis->nsh, 0.0);
  if (screen) orc_decay_cutoff(basis, screening_threshold, extents.data()); // dft.h:293-294, every build
  for (int l = 0; l <= ORC_MAX_L && l <= 8; ++l)
    if (basis->pure) c2s_matrix(l); // build tables before the parallel region
  std::vector<Workspace> wss(nthreads);
  // timing split for the CPU baseline: the leaf loop scales with the leaves handed in, the rest (extents, the
  // per-thread full-size K copies of dft.h:221-232 and their serial reduction, dft.h:375-380) is a fixed cost per build
  double t_loop0 = 0.0, t_loop1 = 0.0;
#pragma omp parallel num_threads(nthreads)
  {
    Workspace &ws = wss[omp_get_thread_num()];
    ws.K.assign(ksz, 0.0);
#pragma omp barrier
#pragma omp master
    t_loop0 = omp_get_wtime();
#pragma omp for schedule(dynamic, 1)
    for (long li = 0; li < static_cast<long>(nleaves); ++li)
      process_leaf(B, extents.data(), spts, sw, leaves[li], nfunc, func_ids, nspin, deriv, density_threshold, screen,
                   D2, ws);
#pragma omp master
    t_loop1 = omp_get_wtime();
  }
  // dft.h:375-380 serial reduction over thread copies
  std::memset(K, 0, sizeof(double) * ksz);
  double e = 0, na = 0, nb = 0;
  orc_build_stats_t st;
  std::memset(&st, 0, sizeof(st));
  st.nb_min = 1 << 30;
  for (auto &ws : wss) {
    for (size_t i = 0; i < ksz; ++i) K[i] += ws.K[i];
    e += ws.energy;
    na += ws.na;
    nb += ws.nbeta;
    st.sum_n_nb2 += ws.sum_n_nb2;
    st.sum_nb2 += ws.sum_nb2;
    st.nskipped += ws.nskipped;
    st.nb_min = std::min(st.nb_min, ws.nb_min);
    st.nb_max = std::max(st.nb_max, ws.nb_max);
    st.t_colloc += ws.t_colloc;
    st.t_density += ws.t_density;
    st.t_func += ws.t_func;
    st.t_vxc += ws.t_vxc;
  }
  st.nleaves = nleaves;
  st.t_total = omp_get_wtime() - tstart;
  st.t_leaf_loop = t_loop1 - t_loop0;
  st.t_fixed = st.t_total - st.t_leaf_loop;
  *exc = e;
  *n_alpha = na;
  *n_beta = nb;
  if (stats) *stats = st;
  return 0;
}

} // extern "C"
