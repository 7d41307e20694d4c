// oracle/src/cosx.cpp — CPU ORACLE (test infrastructure): seminumerical exchange (COSX), SURVEY 8(f) rank 2.
// Follows (paths under /root/reference):
//   include/occ/ints/boys.h:17-64,66-140,142-182,191-330   Boys function: asymptotic form, Chebyshev table (order 7,
//                                                          252 segments on [0,36), 16 orders), Maclaurin reference
//   include/occ/ints/ecoeffs.h:150-198                     McMurchie-Davidson E coefficients (1D, dynamic L)
//   include/occ/ints/rints.h:24-29,158-227                 Hermite Coulomb integrals R_tuv (dynamic L)
//   include/occ/ints/kernels.h:11-18,256-340               sph. factor, E-matrix, primitive-pair precompute
//   include/occ/ints/rints_specialized.h:1451-1500         long-range erf(omega r)/r substitution
//   src/ints/esp.cpp:16-47,85-146,150-252                  spherical E-matrices, pair centre/extent, evaluate, overlap
//   include/occ/qm/integral_engine.h:250-281               significant shell pairs (overlap-block norm >= 1e-12)
//   src/qm/seminumerical_exchange.cpp:293-331,336-380,660-815   pair registration, numerical overlap, grid sweep
// Parity: the Boys / E / R pieces are checked against the reference's own headers compiled into oracle/_ref
// (tests/test_cosx_oracle.py); the assembled K has no golden vector in the reference (its tests only print
// |K - K_exact|), so it is pinned by analytic nuclear-attraction values and by quadrature of the collocated AOs.
#include "oracle.h"
#include "oracle_internal.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {
namespace cosx {

// ---------------------------------------------------------------- Boys function (boys.h)
constexpr int kOrder = 7, kMaxM = 16, kMaxT = 36, kSegments = 252, kN = kOrder + 1;
constexpr double kDelta = static_cast<double>(kMaxT) / kSegments;
constexpr double kSqrtPiOver4 = 0.88622692545275801365;
constexpr double kPi = 3.14159265358979323846;

// boys.h:142-160
static double boys_maclaurin(double T, int m) {
  constexpr double epsilon = 1e-17;
  double denom = m + 0.5;
  double term = std::exp(-T) / (2.0 * denom);
  double old_term = 0.0;
  double sum = term;
  do {
    denom += 1.0;
    old_term = term;
    term = old_term * T / denom;
    sum += term;
  } while (term > sum * epsilon || old_term < term);
  return sum;
}
// boys.h:162-175
static double boys_recursive(double T, int m) {
  if (T == 0.0) return 1.0 / (2 * m + 1);
  double sqrtT = std::sqrt(T);
  double Fm = kSqrtPiOver4 / sqrtT * std::erf(sqrtT);
  double expT = std::exp(-T);
  double inv2T = 0.5 / T;
  for (int i = 0; i < m; ++i) Fm = ((2 * i + 1) * Fm - expT) * inv2T;
  return Fm;
}
// boys.h:177-182
static double boys_reference(double T, int m) {
  if (T < m * 0.25 + 1) return boys_maclaurin(T, m);
  return boys_recursive(T, m);
}

// boys.h:191-330 — Chebyshev interpolation of order 7 per segment, converted to monomials in
// xd = x/Delta - seg - 1/2.  The 8 x 8 collocation system A c = f (A[row][k] = T_k(2 x_row)) is solved by LU with
// partial pivoting, exactly as the reference does (same pivot order, same substitution order).
static std::vector<double> g_table;
static bool g_table_built = false;
// called once at every extern entry, before any parallel region
static void ensure_boys_table() {
  std::vector<double> &table = g_table;
  bool &built = g_table_built;
  if (!built) {
    std::vector<double> t(static_cast<size_t>(kN) * kMaxM * (kSegments + 1), 0.0);
    double nodes[kN];
    for (int k = 0; k < kN; ++k) nodes[k] = std::cos(((2 * k + 1) * kPi) / (2.0 * kN)) / 2.0;
    double Tp[kN][kN] = {};
    Tp[0][0] = 1.0;
    Tp[1][1] = 2.0;
    for (int k = 2; k < kN; ++k)
      for (int j = 0; j < kN; ++j) {
        if (j > 0) Tp[k][j] += 4.0 * Tp[k - 1][j - 1];
        Tp[k][j] -= Tp[k - 2][j];
      }
    double LU[kN][kN];
    for (int row = 0; row < kN; ++row) {
      const double x2 = 2.0 * nodes[row];
      LU[row][0] = 1.0;
      LU[row][1] = x2;
      for (int k = 2; k < kN; ++k) LU[row][k] = 2.0 * x2 * LU[row][k - 1] - LU[row][k - 2];
    }
    int perm[kN];
    for (int i = 0; i < kN; ++i) perm[i] = i;
    for (int k = 0; k < kN; ++k) {
      int max_row = k;
      double max_val = std::abs(LU[k][k]);
      for (int i = k + 1; i < kN; ++i)
        if (std::abs(LU[i][k]) > max_val) {
          max_val = std::abs(LU[i][k]);
          max_row = i;
        }
      if (max_row != k) {
        for (int j = 0; j < kN; ++j) std::swap(LU[k][j], LU[max_row][j]);
        std::swap(perm[k], perm[max_row]);
      }
      for (int i = k + 1; i < kN; ++i) {
        LU[i][k] /= LU[k][k];
        for (int j = k + 1; j < kN; ++j) LU[i][j] -= LU[i][k] * LU[k][j];
      }
    }
    for (int seg = 0; seg < kSegments; ++seg) {
      const double a = seg * kDelta;
      for (int m = 0; m < kMaxM; ++m) {
        double f[kN], y[kN], cheb[kN], mono[kN] = {};
        for (int node = 0; node < kN; ++node) f[node] = boys_maclaurin(a + (nodes[node] + 0.5) * kDelta, m);
        for (int i = 0; i < kN; ++i) {
          y[i] = f[perm[i]];
          for (int j = 0; j < i; ++j) y[i] -= LU[i][j] * y[j];
        }
        for (int i = kN - 1; i >= 0; --i) {
          cheb[i] = y[i];
          for (int j = i + 1; j < kN; ++j) cheb[i] -= LU[i][j] * cheb[j];
          cheb[i] /= LU[i][i];
        }
        for (int k = 0; k < kN; ++k)
          for (int j = 0; j <= k; ++j) mono[j] += cheb[k] * Tp[k][j];
        std::memcpy(&t[(static_cast<size_t>(seg) * kMaxM + m) * kN], mono, sizeof(mono));
      }
    }
    table.swap(t);
    built = true;
  }
}

// boys.h:27-50 (upward in m from the asymptotic F_0) and 108-140 (Horner in xd)
static void boys_evaluate(double x, int n, double *F) {
  if (x >= static_cast<double>(kMaxT)) {
    const double one_over_x = 1.0 / x;
    double Fm = kSqrtPiOver4 * std::sqrt(one_over_x);
    F[0] = Fm;
    for (int i = 1; i < n; ++i) {
      Fm = Fm * (static_cast<double>(i) - 0.5) * one_over_x;
      F[i] = Fm;
    }
    return;
  }
  const double *table = g_table.data();
  const double x_over_delta = x / kDelta;
  const int seg = static_cast<int>(x_over_delta);
  const double xd = x_over_delta - static_cast<double>(seg) - 0.5;
  const double *base = table + static_cast<size_t>(seg) * kMaxM * kN;
  for (int i = 0; i < n; ++i) {
    const double *p = base + i * kN;
    double r = p[kOrder];
    for (int k = kOrder - 1; k >= 0; --k) r = r * xd + p[k];
    F[i] = r;
  }
}

// ---------------------------------------------------------------- E coefficients (ecoeffs.h:150-198)
struct E1D {
  int la, lb, P;
  std::vector<double> d;
  double &at(int i, int j, int t) { return d[t + (P + 1) * (j + (lb + 1) * i)]; }
  double at(int i, int j, int t) const { return d[t + (P + 1) * (j + (lb + 1) * i)]; }
};
static void e_coeffs_1d(int la, int lb, double a, double b, double XAB, E1D &E) {
  E.la = la;
  E.lb = lb;
  E.P = la + lb;
  E.d.assign(static_cast<size_t>(la + 1) * (lb + 1) * (E.P + 1), 0.0);
  const double p = a + b;
  const double mu = a * b / p;
  const double inv_2p = 0.5 / p;
  const double XPA = b / p * XAB;
  const double XPB = -a / p * XAB;
  E.at(0, 0, 0) = std::exp(-mu * XAB * XAB);
  for (int j = 0; j < lb; ++j) {
    for (int t = 0; t <= j; ++t) {
      double val = XPB * E.at(0, j, t);
      if (t > 0) val += inv_2p * E.at(0, j, t - 1);
      if (t < j) val += (t + 1) * E.at(0, j, t + 1);
      E.at(0, j + 1, t) += val;
    }
    E.at(0, j + 1, j + 1) = inv_2p * E.at(0, j, j);
  }
  for (int i = 0; i < la; ++i)
    for (int j = 0; j <= lb; ++j) {
      const int max_t = i + j;
      for (int t = 0; t <= max_t; ++t) {
        double val = XPA * E.at(i, j, t);
        if (t > 0) val += inv_2p * E.at(i, j, t - 1);
        if (t < max_t) val += (t + 1) * E.at(i, j, t + 1);
        E.at(i + 1, j, t) += val;
      }
      E.at(i + 1, j, max_t + 1) = inv_2p * E.at(i, j, max_t);
    }
}

// ---------------------------------------------------------------- R integrals (rints.h:15-29,158-227)
static inline int nhermsum(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }
static inline int hermite_index(int t, int u, int v) {
  const int L = t + u + v;
  const int uv = u + v;
  return nhermsum(L - 1) + uv * (uv + 1) / 2 + v;
}
static void r_ints(int L, double p, double PCx, double PCy, double PCz, double *R /* nhermsum(L) */) {
  const double Tp = p * (PCx * PCx + PCy * PCy + PCz * PCz);
  const double neg_2p = -2.0 * p;
  double Fm[ORC_MAX_L * 2 + 2];
  boys_evaluate(Tp, L + 1, Fm);
  const int nsz = nhermsum(L);
  std::vector<double> R_all(static_cast<size_t>(L + 2) * nsz, 0.0);
  auto R_m = [&](int m, int t, int u, int v) -> double & { return R_all[m * nsz + hermite_index(t, u, v)]; };
  double pw = 1.0;
  for (int m = 0; m <= L; ++m) {
    R_m(m, 0, 0, 0) = pw * Fm[m];
    pw *= neg_2p;
  }
  for (int t = 0; t < L; ++t)
    for (int m = 0; m <= L - t - 1; ++m) {
      double val = PCx * R_m(m + 1, t, 0, 0);
      if (t > 0) val += t * R_m(m + 1, t - 1, 0, 0);
      R_m(m, t + 1, 0, 0) = val;
    }
  for (int t = 0; t <= L; ++t)
    for (int u = 0; u < L - t; ++u)
      for (int m = 0; m <= L - t - u - 1; ++m) {
        double val = PCy * R_m(m + 1, t, u, 0);
        if (u > 0) val += u * R_m(m + 1, t, u - 1, 0);
        R_m(m, t, u + 1, 0) = val;
      }
  for (int t = 0; t <= L; ++t)
    for (int u = 0; u <= L - t; ++u)
      for (int v = 0; v < L - t - u; ++v)
        for (int m = 0; m <= L - t - u - v - 1; ++m) {
          double val = PCz * R_m(m + 1, t, u, v);
          if (v > 0) val += v * R_m(m + 1, t, u, v - 1);
          R_m(m, t, u, v + 1) = val;
        }
  for (int i = 0; i < nsz; ++i) R[i] = R_all[i];
}

// ---------------------------------------------------------------- shell pairs (kernels.h:230-340, esp.cpp:16-47,226-252)
struct Prim {
  double p, P[3], prefactor;
  std::vector<double> E; // [nab x nherm] row-major; spherical pairs: already transformed
};
struct Pair {
  int la, lb, nab, nherm, spherical;
  std::vector<Prim> prims;
  double center[3], extent;
};
static double sph_factor(int l) { return l == 0 ? 0.282094791773878143 : (l == 1 ? 0.488602511902919921 : 1.0); }

// gto.cpp:289-400 cartesian_to_spherical_transformation_matrix: fully normalised real solid harmonics for l >= 2
// (= sqrt((2l+1)/4pi) x the Racah-normalised gau2grid table), the (y, z, x) permutation for l = 1, 1 for l = 0.
static std::vector<double> c2s_full(int l) {
  const int ns = l == 0 ? 1 : 2 * l + 1, nc = ncart(l);
  std::vector<double> T(static_cast<size_t>(ns) * nc, 0.0);
  if (l == 0) {
    T[0] = 1.0;
    return T;
  }
  const std::vector<double> &M = c2s_matrix(l);
  const double f = l >= 2 ? std::sqrt((2 * l + 1) / (4.0 * kPi)) : 1.0;
  for (size_t i = 0; i < T.size(); ++i) T[i] = f * M[i];
  return T;
}

static Pair make_pair(int la, int lb, int na_prim, int nb_prim, const double *ea, const double *eb, const double *ca,
                      const double *cb, const double *A, const double *B, int spherical) {
  Pair pr;
  pr.la = la;
  pr.lb = lb;
  const int L = la + lb, nca = ncart(la), ncb = ncart(lb), nabc = nca * ncb;
  pr.nherm = nhermsum(L);
  pr.spherical = spherical;
  const int nsa = la == 0 ? 1 : 2 * la + 1, nsb = lb == 0 ? 1 : 2 * lb + 1;
  pr.nab = spherical ? nsa * nsb : nabc;
  std::vector<double> Tab;
  if (spherical) { // esp.cpp:16-47
    const std::vector<double> Ta = c2s_full(la), Tb = c2s_full(lb);
    Tab.assign(static_cast<size_t>(nsa) * nsb * nabc, 0.0);
    for (int i = 0; i < nsa; ++i)
      for (int j = 0; j < nsb; ++j)
        for (int k = 0; k < nca; ++k)
          for (int l = 0; l < ncb; ++l)
            Tab[static_cast<size_t>(i * nsb + j) * nabc + k * ncb + l] = Ta[i * nca + k] * Tb[j * ncb + l];
  }
  const double sph_fac = sph_factor(la) * sph_factor(lb);
  std::vector<std::array<int, 3>> pa, pb;
  cart_powers(la, pa);
  cart_powers(lb, pb);
  for (int ia = 0; ia < na_prim; ++ia)
    for (int ib = 0; ib < nb_prim; ++ib) {
      const double a = ea[ia], b = eb[ib];
      Prim pm;
      pm.p = a + b;
      for (int d = 0; d < 3; ++d) pm.P[d] = (a * A[d] + b * B[d]) / pm.p;
      pm.prefactor = 2.0 * kPi / pm.p * sph_fac * ca[ia] * cb[ib];
      E1D Ex, Ey, Ez;
      e_coeffs_1d(la, lb, a, b, B[0] - A[0], Ex); // ecoeffs.h:114-123: XAB = B - A
      e_coeffs_1d(la, lb, a, b, B[1] - A[1], Ey);
      e_coeffs_1d(la, lb, a, b, B[2] - A[2], Ez);
      std::vector<double> Ec(static_cast<size_t>(nabc) * pr.nherm, 0.0); // kernels.h:256-290
      int row = 0;
      for (const auto &pwa : pa)
        for (const auto &pwb : pb) {
          for (int t = 0; t <= pwa[0] + pwb[0]; ++t)
            for (int u = 0; u <= pwa[1] + pwb[1]; ++u)
              for (int v = 0; v <= pwa[2] + pwb[2]; ++v)
                Ec[static_cast<size_t>(row) * pr.nherm + hermite_index(t, u, v)] =
                    Ex.at(pwa[0], pwb[0], t) * Ey.at(pwa[1], pwb[1], u) * Ez.at(pwa[2], pwb[2], v);
          ++row;
        }
      if (spherical) { // esp.cpp:226-252: E_sph = T_ab . E_cart
        pm.E.assign(static_cast<size_t>(pr.nab) * pr.nherm, 0.0);
        for (int r = 0; r < pr.nab; ++r)
          for (int k = 0; k < nabc; ++k) {
            const double tv = Tab[static_cast<size_t>(r) * nabc + k];
            if (tv == 0.0) continue;
            for (int h = 0; h < pr.nherm; ++h) pm.E[static_cast<size_t>(r) * pr.nherm + h] += tv * Ec[static_cast<size_t>(k) * pr.nherm + h];
          }
      } else {
        pm.E.swap(Ec);
      }
      pr.prims.push_back(std::move(pm));
    }
  // esp.cpp:85-100: centre / extent of the most diffuse primitive pair (first minimum wins)
  double p_min = std::numeric_limits<double>::max();
  pr.center[0] = pr.center[1] = pr.center[2] = 0.0;
  for (const auto &pm : pr.prims)
    if (pm.p < p_min) {
      p_min = pm.p;
      std::memcpy(pr.center, pm.P, sizeof(pr.center));
    }
  pr.extent = std::sqrt(27.631021115928547 / p_min);
  return pr;
}

// esp.cpp:102-125 + rints_specialized.h:1451-1500: V[pt, ab] = sum_prim prefactor * sum_h R_h(pt) E[ab, h]
static void pair_evaluate(const Pair &pr, size_t npts, const double *pts /* 3 x npts */, double omega, double *V,
                          std::vector<double> &ws) {
  const int L = pr.la + pr.lb;
  ws.resize(npts * pr.nherm);
  std::fill(V, V + npts * pr.nab, 0.0);
  for (const auto &pm : pr.prims) {
    double p_used = pm.p, rs_scale = 1.0;
    if (omega > 0.0) {
      const double w2 = omega * omega;
      rs_scale = w2 / (pm.p + w2);
      p_used = pm.p * rs_scale;
    }
    for (size_t pt = 0; pt < npts; ++pt)
      r_ints(L, p_used, pm.P[0] - pts[3 * pt], pm.P[1] - pts[3 * pt + 1], pm.P[2] - pts[3 * pt + 2],
             &ws[pt * pr.nherm]);
    if (omega > 0.0) {
      const double s = std::sqrt(rs_scale);
      for (auto &x : ws) x *= s;
    }
    for (size_t pt = 0; pt < npts; ++pt) {
      const double *R = &ws[pt * pr.nherm];
      double *Vp = V + pt * pr.nab;
      for (int ab = 0; ab < pr.nab; ++ab) {
        const double *Er = &pm.E[static_cast<size_t>(ab) * pr.nherm];
        double s = 0.0;
        for (int h = 0; h < pr.nherm; ++h) s += R[h] * Er[h];
        Vp[ab] += pm.prefactor * s;
      }
    }
  }
}

// esp.cpp:127-146
static void pair_overlap(const Pair &pr, double *S) {
  for (int ab = 0; ab < pr.nab; ++ab) S[ab] = 0.0;
  for (const auto &pm : pr.prims) {
    const double f = pm.prefactor * (std::sqrt(kPi / pm.p) / 2.0);
    for (int ab = 0; ab < pr.nab; ++ab) S[ab] += f * pm.E[static_cast<size_t>(ab) * pr.nherm];
  }
}

static Pair shell_pair(const orc_basis_t &b, int p, int q) {
  return make_pair(b.l[p], b.l[q], b.nprim[p], b.nprim[q], b.exps + b.prim_off[p], b.exps + b.prim_off[q],
                   b.coefs + b.prim_off[p], b.coefs + b.prim_off[q], b.origin + 3 * p, b.origin + 3 * q, b.pure);
}

struct PairList {
  std::vector<std::array<int, 2>> pq; // p >= q, insertion order of seminumerical_exchange.cpp:293-331
  std::vector<Pair> pairs;
};

// integral_engine.h:250-281 (threshold 1e-12 on the Frobenius norm of the overlap block; same-origin pairs always kept)
// + seminumerical_exchange.cpp:293-331
static PairList significant_pairs(const Basis &B) {
  const orc_basis_t &b = B.raw;
  PairList out;
  std::vector<double> S;
  for (int p = 0; p < b.nsh; ++p)
    for (int q = 0; q <= p; ++q) {
      const bool same = b.origin[3 * p] == b.origin[3 * q] && b.origin[3 * p + 1] == b.origin[3 * q + 1] &&
                        b.origin[3 * p + 2] == b.origin[3 * q + 2];
      Pair pr = shell_pair(b, p, q);
      if (!same) {
        S.resize(pr.nab);
        pair_overlap(pr, S.data());
        double n2 = 0.0;
        for (double x : S) n2 += x * x;
        if (std::sqrt(n2) < 1e-12) continue;
      }
      out.pq.push_back({p, q});
      out.pairs.push_back(std::move(pr));
    }
  return out;
}

} // namespace cosx
} // namespace orc

using namespace orc;
using namespace orc::cosx;

extern "C" {

double orc_cosx_boys(double x, int m) {
  ensure_boys_table();
  double F[kMaxM];
  boys_evaluate(x, m + 1, F);
  return F[m];
}
double orc_cosx_boys_reference(double x, int m) { return boys_reference(x, m); }

void orc_cosx_ecoeffs_1d(int la, int lb, double a, double b, double XAB, double *out) {
  E1D E;
  e_coeffs_1d(la, lb, a, b, XAB, E);
  std::memcpy(out, E.d.data(), E.d.size() * sizeof(double));
}

void orc_cosx_rints(int L, double p, double PCx, double PCy, double PCz, double *out) {
  ensure_boys_table();
  r_ints(L, p, PCx, PCy, PCz, out);
}

/* One shell pair on npts points: V[pt * nab + ab] (esp.cpp:102-125). Returns nab. */
int orc_cosx_esp_pair(int la, int lb, int na_prim, int nb_prim, const double *ea, const double *eb,
                      const double *ca, const double *cb, const double *A, const double *B, int spherical,
                      size_t npts, const double *pts, double omega, double *V) {
  ensure_boys_table();
  Pair pr = make_pair(la, lb, na_prim, nb_prim, ea, eb, ca, cb, A, B, spherical);
  std::vector<double> ws;
  if (V) pair_evaluate(pr, npts, pts, omega, V, ws);
  return pr.nab;
}

/* Significant shell pairs in the reference's registration order; pq_out may be NULL to query the count. */
size_t orc_cosx_pairs(const orc_basis_t *basis, int *pq_out, double *center_out, double *extent_out) {
  Basis B(basis);
  PairList pl = significant_pairs(B);
  if (pq_out)
    for (size_t i = 0; i < pl.pq.size(); ++i) {
      pq_out[2 * i] = pl.pq[i][0];
      pq_out[2 * i + 1] = pl.pq[i][1];
      if (center_out) std::memcpy(center_out + 3 * i, pl.pairs[i].center, 3 * sizeof(double));
      if (extent_out) extent_out[i] = pl.pairs[i].extent;
    }
  return pl.pq.size();
}

/* Analytic overlap matrix (nbf x nbf, column-major) assembled from ESPEvaluator::evaluate_overlap blocks
 * (esp.cpp:127-146). The reference takes m_overlap from libcint; the two agree to rounding (tests/ints_tests.cpp:285-350). */
void orc_cosx_overlap(const orc_basis_t *basis, double *S) {
  Basis B(basis);
  const orc_basis_t &b = B.raw;
  const int nbf = B.nbf;
  std::fill(S, S + static_cast<size_t>(nbf) * nbf, 0.0);
  std::vector<double> blk;
  for (int p = 0; p < b.nsh; ++p)
    for (int q = 0; q <= p; ++q) {
      Pair pr = shell_pair(b, p, q);
      blk.resize(pr.nab);
      pair_overlap(pr, blk.data());
      const int np = orc_nbf_shell(b.l[p], b.pure), nq = orc_nbf_shell(b.l[q], b.pure);
      for (int i = 0; i < np; ++i)
        for (int j = 0; j < nq; ++j) {
          const double v = blk[i * nq + j];
          S[(B.first_bf[p] + i) + static_cast<size_t>(B.first_bf[q] + j) * nbf] = v;
          S[(B.first_bf[q] + j) + static_cast<size_t>(B.first_bf[p] + i) * nbf] = v;
        }
    }
}

static void collocate_all(const Basis &B, size_t npt, const double *pts, double *phi /* npt x nbf col-major */) {
  const orc_basis_t &b = B.raw;
  for (int s = 0; s < b.nsh; ++s) {
    double *out = phi + static_cast<size_t>(B.first_bf[s]) * npt;
    collocate_shell(b.l[s], b.pure, npt, pts, b.nprim[s], b.coefs + b.prim_off[s], b.exps + b.prim_off[s],
                    b.origin + 3 * s, 0, out, nullptr, nullptr, nullptr);
    const double fac = orc_common_fac(b.l[s], b.pure);
    const int n = orc_nbf_shell(b.l[s], b.pure);
    for (size_t i = 0; i < npt * n; ++i) out[i] *= fac;
  }
}

/* seminumerical_exchange.cpp:336-380 — numerical overlap on the COSX grid, blocks of 64 per atom grid.
 * The block bounds are the reference's: u = min(npt_total - 1, (block + 1) * 64), so the LAST point of every atom
 * grid is left out. atom_off[natoms + 1] delimits the atom grids inside pts (3 x N interleaved) / w. */
void orc_cosx_numerical_overlap(const orc_basis_t *basis, int natoms, const size_t *atom_off, const double *pts,
                                const double *w, double *Sn) {
  Basis B(basis);
  const int nbf = B.nbf;
  std::fill(Sn, Sn + static_cast<size_t>(nbf) * nbf, 0.0);
  std::vector<double> phi, wphi;
  for (int a = 0; a < natoms; ++a) {
    const size_t npt_total = atom_off[a + 1] - atom_off[a];
    const size_t num_blocks = npt_total / 64 + 1;
    for (size_t block = 0; block < num_blocks; ++block) {
      const long l = static_cast<long>(block * 64);
      const long u = static_cast<long>(std::min(npt_total - 1, (block + 1) * 64));
      const long npt = u - l;
      if (npt <= 0) continue;
      const double *bp = pts + 3 * (atom_off[a] + l);
      const double *bw = w + atom_off[a] + l;
      phi.resize(static_cast<size_t>(npt) * nbf);
      wphi.resize(phi.size());
      collocate_all(B, npt, bp, phi.data());
      for (int c = 0; c < nbf; ++c)
        for (long i = 0; i < npt; ++i) wphi[c * npt + i] = phi[c * npt + i] * bw[i];
      dgemm('T', 'N', nbf, nbf, static_cast<int>(npt), 1.0, phi.data(), static_cast<int>(npt), wphi.data(),
            static_cast<int>(npt), 1.0, Sn, nbf);
    }
  }
}

typedef struct {
  size_t nbatches, nskipped_batches;
  size_t pair_evals;     /* (batch, pair) combinations evaluated */
  double dense_flops;    /* 2 * npt * nherm * nab per primitive pair evaluated (the reference's R.E^T product) */
  size_t npairs;
} orc_cosx_stats_t;

/* seminumerical_exchange.cpp:660-815 — compute_K_for_density: the raw (unsymmetrised) K for one projected density
 * D2q (nbf x nbf column-major). Batches are blocks of 128 consecutive points of each atom grid. */
int orc_cosx_K_for_density(const orc_basis_t *basis, int natoms, const size_t *atom_off, const double *pts,
                           const double *w, const double *D2q, double precision, double omega, double margin,
                           double *K, orc_cosx_stats_t *stats) {
  ensure_boys_table();
  Basis B(basis);
  const orc_basis_t &b = B.raw;
  const int nbf = B.nbf;
  PairList pl = significant_pairs(B);
  const size_t npair = pl.pq.size();
  // density screen (seminumerical_exchange.cpp:758-764) does not depend on the batch
  std::vector<char> dens_ok(npair);
  for (size_t k = 0; k < npair; ++k) {
    const int p = pl.pq[k][0], q = pl.pq[k][1];
    const int bp = B.first_bf[p], bq = B.first_bf[q];
    const int sp = orc_nbf_shell(b.l[p], b.pure), sq = orc_nbf_shell(b.l[q], b.pure);
    double mx = 0.0;
    for (int i = 0; i < sp; ++i)
      for (int j = 0; j < sq; ++j) {
        mx = std::max(mx, std::abs(D2q[(bp + i) + static_cast<size_t>(bq + j) * nbf]));
        if (p != q) mx = std::max(mx, std::abs(D2q[(bq + j) + static_cast<size_t>(bp + i) * nbf]));
      }
    dens_ok[k] = !(mx < 1e-6);
  }
  struct Batch {
    size_t start, count;
  };
  std::vector<Batch> batches;
  for (int a = 0; a < natoms; ++a) {
    const size_t n = atom_off[a + 1] - atom_off[a];
    for (size_t s = 0; s < n; s += 128) batches.push_back({atom_off[a] + s, std::min<size_t>(128, n - s)});
  }
  std::fill(K, K + static_cast<size_t>(nbf) * nbf, 0.0);
  size_t nskip = 0, nevals = 0;
  double flops = 0.0;
#pragma omp parallel reduction(+ : nskip, nevals, flops)
  {
    std::vector<double> Kloc(static_cast<size_t>(nbf) * nbf, 0.0), phi, wao, Fg, Gg, V, ws;
#pragma omp for schedule(dynamic, 1)
    for (long bi = 0; bi < static_cast<long>(batches.size()); ++bi) {
      const size_t npt = batches[bi].count;
      const double *bp = pts + 3 * batches[bi].start;
      const double *bw = w + batches[bi].start;
      double c[3] = {0, 0, 0};
      for (size_t i = 0; i < npt; ++i)
        for (int d = 0; d < 3; ++d) c[d] += bp[3 * i + d];
      for (int d = 0; d < 3; ++d) c[d] /= static_cast<double>(npt);
      double radius = 0.0;
      for (size_t i = 0; i < npt; ++i) {
        const double dx = bp[3 * i] - c[0], dy = bp[3 * i + 1] - c[1], dz = bp[3 * i + 2] - c[2];
        radius = std::max(radius, std::sqrt(dx * dx + dy * dy + dz * dz));
      }
      phi.resize(npt * nbf);
      collocate_all(B, npt, bp, phi.data());
      if (*std::max_element(phi.begin(), phi.end()) < precision) {
        ++nskip;
        continue;
      }
      wao.resize(phi.size());
      for (int col = 0; col < nbf; ++col)
        for (size_t i = 0; i < npt; ++i) wao[col * npt + i] = phi[col * npt + i] * bw[i];
      Fg.assign(npt * nbf, 0.0);
      dgemm('N', 'N', static_cast<int>(npt), nbf, nbf, 1.0, wao.data(), static_cast<int>(npt), D2q, nbf, 0.0,
            Fg.data(), static_cast<int>(npt));
      Gg.assign(npt * nbf, 0.0);
      for (size_t k = 0; k < npair; ++k) {
        const Pair &pr = pl.pairs[k];
        const double dx = pr.center[0] - c[0], dy = pr.center[1] - c[1], dz = pr.center[2] - c[2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) - pr.extent >= radius + margin) continue;
        if (!dens_ok[k]) continue;
        const int p = pl.pq[k][0], q = pl.pq[k][1];
        const int bfp = B.first_bf[p], bfq = B.first_bf[q];
        const int sp = orc_nbf_shell(b.l[p], b.pure), sq = orc_nbf_shell(b.l[q], b.pure);
        V.resize(npt * pr.nab);
        pair_evaluate(pr, npt, bp, omega, V.data(), ws);
        ++nevals;
        flops += 2.0 * npt * pr.nherm * pr.nab * pr.prims.size();
        for (size_t pt = 0; pt < npt; ++pt) {
          const double *Vp = &V[pt * pr.nab]; // (sp x sq) row-major
          for (int j = 0; j < sq; ++j) {
            double s = 0.0;
            for (int i = 0; i < sp; ++i) s += Fg[(bfp + i) * npt + pt] * Vp[i * sq + j];
            Gg[(bfq + j) * npt + pt] -= s;
          }
          if (p != q)
            for (int i = 0; i < sp; ++i) {
              double s = 0.0;
              for (int j = 0; j < sq; ++j) s += Fg[(bfq + j) * npt + pt] * Vp[i * sq + j];
              Gg[(bfp + i) * npt + pt] -= s;
            }
        }
      }
      dgemm('T', 'N', nbf, nbf, static_cast<int>(npt), -1.0, phi.data(), static_cast<int>(npt), Gg.data(),
            static_cast<int>(npt), 1.0, Kloc.data(), nbf);
    }
#pragma omp critical(orc_cosx_K)
    for (size_t i = 0; i < Kloc.size(); ++i) K[i] += Kloc[i];
  }
  if (stats) {
    stats->nbatches = batches.size();
    stats->nskipped_batches = nskip;
    stats->pair_evals = nevals;
    stats->dense_flops = flops;
    stats->npairs = npair;
  }
  return 0;
}

} // extern "C"
