// oracle/src/gto.cpp — CPU ORACLE (test infrastructure): shell normalisation, AO collocation, decay extents.
// Follows (paths under /root/reference):
//   src/gto/shell.cpp:20-23,35-64,180-186   gint / gto_norm / normalize_contracted_gto / incorporate_shell_norm
//   src/gto/gto.cpp:7-16                    common_fac
//   src/gto/gto.cpp:18-67                   evaluate_decay_cutoff
//   src/3rdparty/gau2grid/gau2grid_phi.c, gau2grid_deriv1.c, gau2grid_transform.c  (arithmetic of the
//   collocation; restated here as a generic-L routine, cross-checked against the compiled original)
#include "oracle.h"
#include "oracle_internal.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// ---- src/gto/shell.cpp:20-23 ----
static double gint(int n, double alpha) {
  double n1_2 = 0.5 * (n + 1);
  return std::tgamma(n1_2) / (2 * std::pow(alpha, n1_2));
}
// ---- src/gto/shell.cpp:35-50 ----
static double gto_norm(int l, double alpha) { return 1.0 / std::sqrt(gint(l * 2 + 2, 2 * alpha)); }

int ncart(int l) { return (l + 1) * (l + 2) / 2; }
int nsph(int l) { return 2 * l + 1; }

// Cartesian exponents in CCA order (gau2grid shell order: lx descending, then ly descending).
void cart_powers(int l, std::vector<std::array<int, 3>> &out) {
  out.clear();
  for (int lx = l; lx >= 0; --lx)
    for (int ly = l - lx; ly >= 0; --ly) out.push_back({lx, ly, l - lx - ly});
}

static double binom(int n, int k) {
  if (k < 0 || k > n) return 0.0;
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return r;
}
static double fact(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}

// Racah-normalised real solid harmonics expanded in Cartesian monomials (Helgaker, Jorgensen, Olsen,
// "Molecular Electronic-Structure Theory", eq. 6.4.47-6.4.50), rows m = -l..l (CCA order), columns in
// cart_powers order. Reproduces the tables of gau2grid_transform.c:44-196 (checked in tests).
const std::vector<double> &c2s_matrix(int l) {
  static std::vector<double> cache[ORC_MAX_L + 1];
  static bool built[ORC_MAX_L + 1] = {false};
  if (built[l]) return cache[l];
  std::vector<std::array<int, 3>> pw;
  cart_powers(l, pw);
  const int nc = ncart(l), ns = nsph(l);
  std::vector<double> M(static_cast<size_t>(ns) * nc, 0.0);
  auto cart_index = [&](int lx, int ly, int lz) {
    for (int i = 0; i < nc; ++i)
      if (pw[i][0] == lx && pw[i][1] == ly && pw[i][2] == lz) return i;
    return -1;
  };
  for (int m = -l; m <= l; ++m) {
    const int am = std::abs(m);
    const double N = (1.0 / (std::pow(2.0, am) * fact(l))) *
                     std::sqrt(2.0 * fact(l + am) * fact(l - am) / (m == 0 ? 2.0 : 1.0));
    const int two_vm = (m < 0) ? 1 : 0; // 2*v_m
    for (int t = 0; t <= (l - am) / 2; ++t)
      for (int u = 0; u <= t; ++u)
        // v runs over v_m, v_m+1, ... <= floor(|m|/2 - v_m) + v_m ; iterate 2v
        for (int two_v = two_vm; two_v <= 2 * ((am - two_vm) / 2) + two_vm; two_v += 2) {
          const int vi = (two_v - two_vm) / 2; // v - v_m
          const double sign = ((t + vi) % 2 == 0) ? 1.0 : -1.0;
          const double C = sign * std::pow(0.25, t) * binom(l, t) * binom(l - t, am + t) * binom(t, u) *
                           binom(am, two_v);
          const int lx = 2 * t + am - 2 * u - two_v;
          const int ly = 2 * u + two_v;
          const int lz = l - 2 * t - am;
          if (lx < 0 || ly < 0 || lz < 0) continue;
          M[static_cast<size_t>(m + l) * nc + cart_index(lx, ly, lz)] += N * C;
        }
  }
#pragma omp critical(orc_c2s)
  {
    if (!built[l]) {
      cache[l] = M;
      built[l] = true;
    }
  }
  return cache[l];
}

static inline double ipow(double x, int n) {
  double r = 1.0;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// One shell on npts points. out arrays are component-major [comp*npts + pt], no common_fac.
void collocate_shell(int l, int pure, size_t npts, const double *xyz, int nprim, const double *coefs,
                     const double *exps, const double *center, int deriv, double *phi, double *phi_x,
                     double *phi_y, double *phi_z) {
  std::vector<std::array<int, 3>> pw;
  cart_powers(l, pw);
  const int nc = ncart(l);
  const int nout = pure ? nsph(l) : nc;
  const std::vector<double> *c2s = (pure && l >= 1) ? &c2s_matrix(l) : nullptr;
  std::vector<double> c(nc), cx(nc), cy(nc), cz(nc);
  for (size_t p = 0; p < npts; ++p) {
    // gau2grid_deriv1.c:474-488 (xyz_stride = 3)
    const double xc = xyz[3 * p + 0] - center[0];
    const double yc = xyz[3 * p + 1] - center[1];
    const double zc = xyz[3 * p + 2] - center[2];
    double R2 = xc * xc;
    R2 += yc * yc;
    R2 += zc * zc;
    double S0 = 0.0, S1 = 0.0;
    // gau2grid_deriv1.c:492-505
    for (int n = 0; n < nprim; ++n) {
      const double T1 = coefs[n] * std::exp(-1.0 * exps[n] * R2);
      S0 += T1;
      S1 += (-2.0 * exps[n]) * T1;
    }
    const double SX = S1 * xc, SY = S1 * yc, SZ = S1 * zc;
    for (int i = 0; i < nc; ++i) {
      const int lx = pw[i][0], ly = pw[i][1], lz = pw[i][2];
      const double A = ipow(xc, lx) * ipow(yc, ly) * ipow(zc, lz);
      c[i] = S0 * A;
      if (deriv >= 1) {
        // gau2grid_deriv1.c:520-581: d/dx (S0 A) = S1 x A + S0 dA/dx
        const double AX = lx ? lx * ipow(xc, lx - 1) * ipow(yc, ly) * ipow(zc, lz) : 0.0;
        const double AY = ly ? ly * ipow(xc, lx) * ipow(yc, ly - 1) * ipow(zc, lz) : 0.0;
        const double AZ = lz ? lz * ipow(xc, lx) * ipow(yc, ly) * ipow(zc, lz - 1) : 0.0;
        cx[i] = SX * A + S0 * AX;
        cy[i] = SY * A + S0 * AY;
        cz[i] = SZ * A + S0 * AZ;
      }
    }
    if (!c2s) {
      // pure with l == 0 is the identity (gau2grid_transform.c:23-31)
      for (int i = 0; i < nout; ++i) {
        phi[static_cast<size_t>(i) * npts + p] = c[i];
        if (deriv >= 1) {
          phi_x[static_cast<size_t>(i) * npts + p] = cx[i];
          phi_y[static_cast<size_t>(i) * npts + p] = cy[i];
          phi_z[static_cast<size_t>(i) * npts + p] = cz[i];
        }
      }
    } else {
      for (int m = 0; m < nout; ++m) {
        const double *row = c2s->data() + static_cast<size_t>(m) * nc;
        double v = 0, vx = 0, vy = 0, vz = 0;
        for (int i = 0; i < nc; ++i) {
          if (row[i] == 0.0) continue;
          v += row[i] * c[i];
          if (deriv >= 1) {
            vx += row[i] * cx[i];
            vy += row[i] * cy[i];
            vz += row[i] * cz[i];
          }
        }
        phi[static_cast<size_t>(m) * npts + p] = v;
        if (deriv >= 1) {
          phi_x[static_cast<size_t>(m) * npts + p] = vx;
          phi_y[static_cast<size_t>(m) * npts + p] = vy;
          phi_z[static_cast<size_t>(m) * npts + p] = vz;
        }
      }
    }
  }
}

Basis::Basis(const orc_basis_t *b) : raw(*b) {
  first_bf.resize(b->nsh);
  nbf = 0;
  max_shell_size = 0;
  for (int s = 0; s < b->nsh; ++s) {
    first_bf[s] = nbf;
    int sz = orc_nbf_shell(b->l[s], b->pure);
    nbf += sz;
    if (sz > max_shell_size) max_shell_size = sz;
  }
}

} // namespace orc

extern "C" {

int orc_nbf_shell(int l, int pure) { return pure ? orc::nsph(l) : orc::ncart(l); }

// src/gto/gto.cpp:7-16
double orc_common_fac(int l, int pure) {
  switch (l) {
  case 0:
    return 0.282094791773878143;
  case 1:
    return 0.488602511902919921;
  default:
    return pure ? std::sqrt((2 * l + 1) / (4 * M_PI)) : 1.0;
  }
}

// src/gto/shell.cpp:180-186 (primitive norms) then :52-64 (contracted norm)
void orc_incorporate_shell_norm(int l, int nprim, const double *exps, double *c) {
  for (int i = 0; i < nprim; ++i) c[i] *= orc::gto_norm(l, exps[i]);
  double s1 = 0.0;
  for (int j = 0; j < nprim; ++j) {
    double t = 0.0;
    for (int i = 0; i < nprim; ++i) t += c[i] * orc::gint(l * 2 + 2, exps[i] + exps[j]);
    s1 += t * c[j];
  }
  s1 = 1.0 / std::sqrt(s1);
  for (int i = 0; i < nprim; ++i) c[i] *= s1;
}

void orc_collocate_shell(int l, int pure, size_t npts, const double *xyz, int nprim, const double *coefs,
                         const double *exps, const double *center, int deriv, double *phi, double *phi_x,
                         double *phi_y, double *phi_z) {
  orc::collocate_shell(l, pure, npts, xyz, nprim, coefs, exps, center, deriv, phi, phi_x, phi_y, phi_z);
}

// src/gto/gto.cpp:18-67.  The probe points are carried through "pts += origin ... pts -= origin"
// exactly as the reference does (the round-off of that shuttle is part of its result), shells are
// visited atom by atom in atom_to_shell order, the probe line is +x only.
void orc_decay_cutoff(const orc_basis_t *b, double threshold, double *extents) {
  const double max_radius = 30.0, increment = 0.1, min_radius = 0.5;
  const int npts = static_cast<int>((max_radius - min_radius) / increment) + 1;
  std::vector<double> pts(3 * static_cast<size_t>(npts), 0.0);
  // Vec::LinSpaced(npts, lo, hi): lo + i*(hi-lo)/(npts-1), last element exactly hi
  for (int i = 0; i < npts; ++i) {
    double v = (i == npts - 1) ? max_radius : min_radius + i * ((max_radius - min_radius) / (npts - 1));
    pts[3 * i] = 1.0 * v;
  }
  int natoms = 0;
  for (int s = 0; s < b->nsh; ++s)
    if (b->shell_atom[s] + 1 > natoms) natoms = b->shell_atom[s] + 1;
  int max_shell_size = 0;
  for (int s = 0; s < b->nsh; ++s) max_shell_size = std::max(max_shell_size, orc_nbf_shell(b->l[s], b->pure));
  std::vector<double> values(static_cast<size_t>(npts) * max_shell_size);
  for (int a = 0; a < natoms; ++a) {
    for (int s = 0; s < b->nsh; ++s) {
      if (b->shell_atom[s] != a) continue;
      const double *o = b->origin + 3 * s;
      for (int i = 0; i < npts; ++i)
        for (int d = 0; d < 3; ++d) pts[3 * i + d] += o[d];
      const int L = b->l[s];
      const int size = orc_nbf_shell(L, b->pure);
      const double fac = orc_common_fac(L, b->pure);
      orc::collocate_shell(L, b->pure, npts, pts.data(), b->nprim[s], b->coefs + b->prim_off[s],
                           b->exps + b->prim_off[s], o, 0, values.data(), nullptr, nullptr, nullptr);
      for (int i = 0; i < npts; ++i) {
        extents[s] = i * increment + min_radius;
        double v = 0.0;
        for (int c = 0; c < size; ++c) v = std::max(v, std::fabs(values[static_cast<size_t>(c) * npts + i] * fac));
        if (v < threshold) break;
      }
      for (int i = 0; i < npts; ++i)
        for (int d = 0; d < 3; ++d) pts[3 * i + d] -= o[d];
    }
  }
}

} // extern "C"
