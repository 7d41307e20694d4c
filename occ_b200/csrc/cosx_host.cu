// cosx_host.cu — host orchestration of the B200 seminumerical exchange (COSX) build; C ABI of include/occ_cosx_b200.h.
//
// Reference: occ::qm::cosx::SemiNumericalExchange (src/qm/seminumerical_exchange.cpp), ESP mode.
//   create():   what the reference precomputes per geometry — significant shell pairs (integral_engine.h:250-281,
//               seminumerical_exchange.cpp:293-331), their McMurchie-Davidson E tables, prefactors, centres and extents
//               (include/occ/ints/ecoeffs.h:150-198, kernels.h:293-340, src/ints/esp.cpp:85-100), the grid batches with
//               their bounding spheres (:676-720) and the Boys table (boys.h:191-330) — built on the host, once.
//   compute():  per SCF iteration, per chunk of batches: AO values (k_collocate) -> batch skip flags -> F = Phi D2q
//               (k_dgemm, DMMA; factorised as (Phi (P C)) C^T when the caller passes the occupied orbitals) -> ESP +
//               contraction kernels, one launch per (la, lb) class (k_cosx_esp) -> K' += Phi^T G (k_cosx_kgemm, DMMA)
//               -> K = c (K' + K'^T).
// Every product, including the plain batched GEMMs F = Phi D2q / P C / P (2 D), runs on this library's own DMMA
// kernels (k_dgemm.cu, k_cosx_kgemm.cu): no cuBLAS.  No CPU fallback.
#include "../../include/occ_cosx_b200.h"
#include "cosx_internal.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>
#include <thread>
#include <vector>

using namespace xcb;

int occ_xc_set_error(const char *msg); // xc_host.cu
static int fail(const std::string &m) { return occ_xc_set_error(m.c_str()); }
#define CUDA_TRY(expr)                                                                                       \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return fail(std::string("CUDA error: ") + cudaGetErrorString(e__) + " in " #expr);                    \
  } while (0)

namespace {

constexpr double kPi = 3.14159265358979323846;

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }
inline int nfun(int l, int pure) { return pure ? 2 * l + 1 : ncart(l); }

// include/occ/ints/ecoeffs.h:150-198 — E_t^{ij} of one Cartesian direction, layout t + (la+lb+1) (j + (lb+1) i)
void e_table(int la, int lb, double a, double b, double XAB, double *E) {
  const int P1 = la + lb + 1;
  auto at = [&](int i, int j, int t) -> double & { return E[t + P1 * (j + (lb + 1) * i)]; };
  std::fill(E, E + (la + 1) * (lb + 1) * P1, 0.0);
  const double p = a + b, mu = a * b / p, inv_2p = 0.5 / p;
  const double XPA = b / p * XAB, XPB = -a / p * XAB;
  at(0, 0, 0) = std::exp(-mu * XAB * XAB);
  for (int j = 0; j < lb; ++j) {
    for (int t = 0; t <= j; ++t) {
      double val = XPB * at(0, j, t);
      if (t > 0) val += inv_2p * at(0, j, t - 1);
      if (t < j) val += (t + 1) * at(0, j, t + 1);
      at(0, j + 1, t) += val;
    }
    at(0, j + 1, j + 1) = inv_2p * at(0, j, j);
  }
  for (int i = 0; i < la; ++i)
    for (int j = 0; j <= lb; ++j) {
      const int max_t = i + j;
      for (int t = 0; t <= max_t; ++t) {
        double val = XPA * at(i, j, t);
        if (t > 0) val += inv_2p * at(i, j, t - 1);
        if (t < max_t) val += (t + 1) * at(i, j, t + 1);
        at(i + 1, j, t) += val;
      }
      at(i + 1, j, max_t + 1) = inv_2p * at(i, j, max_t);
    }
}

// gto.cpp:289-400: fully normalised c2s (l >= 2: sqrt((2l+1)/4pi) x Racah table; l = 1: (y, z, x); l = 0: 1)
std::vector<double> c2s_full(int l) {
  const int ns = 2 * l + 1, nc = ncart(l);
  std::vector<double> T(static_cast<size_t>(ns) * nc, 0.0);
  const double *M = c2s_racah_table(l);
  const double f = l >= 2 ? std::sqrt((2 * l + 1) / (4.0 * kPi)) : 1.0;
  for (size_t i = 0; i < T.size(); ++i) T[i] = f * M[i];
  return T;
}

// boys.h:142-160
double boys_maclaurin(double T, int m) {
  double denom = m + 0.5, term = std::exp(-T) / (2.0 * denom), old_term = 0.0, sum = term;
  do {
    denom += 1.0;
    old_term = term;
    term = old_term * T / denom;
    sum += term;
  } while (term > sum * 1e-17 || old_term < term);
  return sum;
}

// boys.h:191-330: order-7 Chebyshev interpolant per segment of width 1/7 on [0, 36), 16 orders, stored as monomials in
// xd = 7x - segment - 1/2. The Chebyshev coefficients come from the discrete orthogonality of the 8 Chebyshev nodes
// (the reference solves the equivalent 8 x 8 collocation system by LU).
std::vector<double> boys_table() {
  constexpr int N = 8, MaxM = 16, Segments = 252;
  const double Delta = 36.0 / Segments;
  std::vector<double> t(static_cast<size_t>(N) * MaxM * (Segments + 1), 0.0);
  double nodes[N], Tp[N][N] = {};
  for (int k = 0; k < N; ++k) nodes[k] = std::cos(((2 * k + 1) * kPi) / (2.0 * N)) / 2.0;
  Tp[0][0] = 1.0;
  Tp[1][1] = 2.0; // T_1(2 x) = 2 x
  for (int k = 2; k < N; ++k)
    for (int j = 0; j < N; ++j) {
      if (j > 0) Tp[k][j] += 4.0 * Tp[k - 1][j - 1];
      Tp[k][j] -= Tp[k - 2][j];
    }
  for (int seg = 0; seg < Segments; ++seg)
    for (int m = 0; m < MaxM; ++m) {
      double f[N], cheb[N], mono[N] = {};
      for (int n = 0; n < N; ++n) f[n] = boys_maclaurin(seg * Delta + (nodes[n] + 0.5) * Delta, m);
      for (int k = 0; k < N; ++k) {
        double s = 0.0;
        for (int n = 0; n < N; ++n) s += f[n] * std::cos(k * (2 * n + 1) * kPi / (2.0 * N));
        cheb[k] = s * (k == 0 ? 1.0 : 2.0) / N;
      }
      for (int k = 0; k < N; ++k)
        for (int j = 0; j <= k; ++j) mono[j] += cheb[k] * Tp[k][j];
      std::memcpy(&t[(static_cast<size_t>(seg) * MaxM + m) * N], mono, sizeof(mono));
    }
  return t;
}

struct HostPair {
  CosxPair d;
  int la, lb;
};

} // namespace

struct occ_cosx_handle {
  int device = 0, nbf = 0, nsh = 0, pure = 0, natoms = 0, nbp = 0;
  size_t npts = 0;
  int ntiles = 0, tiles_per_chunk = 1, nruns = 1;
  double margin = 1.0, density_threshold = 1e-6;
  bool collect = false;
  std::vector<void *> allocs;
  size_t alloc_bytes = 0;
  BuildParams B = {};      // what k_collocate reads
  CollocItem *d_items = nullptr;
  int *d_counts = nullptr, *d_skip = nullptr, *d_shell_size = nullptr;
  double *d_tc[4] = {nullptr, nullptr, nullptr, nullptr};
  double *d_pw_ovl = nullptr; // weights with the last point of every atom grid zeroed (numerical overlap quirk)
  CosxPair *d_pairs = nullptr;
  double *d_pool = nullptr, *d_boys = nullptr, *d_c2s = nullptr;
  unsigned char *d_active = nullptr;
  unsigned long long *d_stats = nullptr;
  double *d_phi = nullptr, *d_F = nullptr, *d_G = nullptr;
  double *d_D = nullptr, *d_D2q = nullptr, *d_Kraw = nullptr, *d_K = nullptr, *d_proj = nullptr;
  double *d_Dg = nullptr, *d_Kg = nullptr, *d_Kraw4 = nullptr; // general spinorbitals (allocated on first use)
  double *d_C = nullptr, *d_Cq = nullptr, *d_psi = nullptr;    // factorised F (occ_cosx_compute_K_mo; allocated on first use)
  int mo_cap = 0;                                               // orbitals per spin the three buffers hold
  size_t npairs = 0, nprim_pairs = 0;
  struct ClassRange {
    int la, lb, pair0, npairs;
  };
  std::vector<ClassRange> classes;
  std::vector<double> S_analytic;
  std::vector<int> h_counts;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  occ_cosx_info_t info = {};
};

template <class T> static int dev_alloc(occ_cosx_handle *h, T **ptr, size_t count) {
  void *p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  h->alloc_bytes += bytes;
  *ptr = static_cast<T *>(p);
  return 0;
}
template <class T> static void dev_free(occ_cosx_handle *h, T **ptr) { // a buffer that is being replaced by a larger one
  if (!*ptr) return;
  auto it = std::find(h->allocs.begin(), h->allocs.end(), static_cast<void *>(*ptr));
  if (it != h->allocs.end()) h->allocs.erase(it);
  cudaFree(*ptr);
  *ptr = nullptr;
}
template <class T> static int dev_upload(occ_cosx_handle *h, T **ptr, const std::vector<T> &v) {
  if (dev_alloc(h, ptr, v.size())) return 1;
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" {

void occ_cosx_destroy(occ_cosx_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void *p : h->allocs) cudaFree(p);
  for (auto &e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int occ_cosx_create(const occ_xc_basis_t *basis, int natoms, const size_t *atom_offsets, const double *pts_xyz,
                    const double *weights, const double *overlap_projector, const occ_cosx_options_t *opt,
                    occ_cosx_handle **out) {
  if (!basis || !atom_offsets || !pts_xyz || !weights || !out || natoms < 1) return fail("occ_cosx_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("occ_cosx_create: no CUDA device available (this library has no CPU fallback)");
  int device = (opt && opt->device >= 0) ? opt->device : -1;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= ndev) return fail("occ_cosx_create: device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  const int world = (opt && opt->world_size > 1) ? opt->world_size : 1;
  const int rank = world > 1 ? opt->rank : 0;
  if (rank < 0 || rank >= world) return fail("occ_cosx_create: rank outside [0, world_size)");
  const int nsh = basis->nsh, pure = basis->pure ? 1 : 0;
  std::vector<int> first_bf(nsh), shell_size(nsh);
  int nbf = 0;
  for (int s = 0; s < nsh; ++s) {
    if (basis->l[s] < 0 || basis->l[s] > COSX_MAX_L)
      return fail("occ_cosx_create: shells up to l = 4 (g) are supported by the COSX build");
    first_bf[s] = nbf;
    shell_size[s] = nfun(basis->l[s], pure);
    nbf += shell_size[s];
  }
  occ_cosx_handle *h = new occ_cosx_handle;
  h->device = device;
  h->nbf = nbf;
  h->nsh = nsh;
  h->pure = pure;
  h->natoms = natoms;
  h->nbp = (nbf + NB_PAD - 1) / NB_PAD * NB_PAD;
  h->margin = (opt && opt->margin > 0.0) ? opt->margin : 1.0;
  h->density_threshold = (opt && opt->density_threshold > 0.0) ? opt->density_threshold : 1e-6;
  h->collect = opt && opt->collect_stats;
  const double pair_threshold = (opt && opt->pair_threshold > 0.0) ? opt->pair_threshold : 1e-12;
  const size_t batch_bytes = (opt && opt->batch_bytes) ? opt->batch_bytes : (size_t(4096) << 20);
  h->npts = atom_offsets[natoms] - atom_offsets[0];
#define CHECK(x)                                                                                             \
  do {                                                                                                       \
    if (x) {                                                                                                 \
      occ_cosx_destroy(h);                                                                                   \
      return 1;                                                                                              \
    }                                                                                                        \
  } while (0)

  // ---- shell pairs, grouped by (la >= lb) class, sorted by the resident shell inside a class ----
  std::vector<double> Tfull[COSX_MAX_L + 1];
  for (int l = 0; l <= COSX_MAX_L; ++l) Tfull[l] = c2s_full(l);
  constexpr int NCLS = (COSX_MAX_L + 1) * (COSX_MAX_L + 1);
  std::vector<HostPair> hp[NCLS];
  std::vector<double> pool[NCLS];
  h->S_analytic.assign(static_cast<size_t>(nbf) * nbf, 0.0);
  std::vector<double> Ex, Ey, Ez, Sc, Ss, tmp;
  for (int p0 = 0; p0 < nsh; ++p0)
    for (int q0 = 0; q0 <= p0; ++q0) {
      // roles: the higher-l shell is the "p" side of the kernel (its G row stays in registers across the run of pairs
      // that share it; the per-pair column traffic is then that of the lower-l shell). V is symmetric under the swap.
      const bool swap_roles = basis->l[p0] < basis->l[q0];
      const int p = swap_roles ? q0 : p0, q = swap_roles ? p0 : q0;
      const int la = basis->l[p], lb = basis->l[q], L1 = la + lb + 1, ne = (la + 1) * (lb + 1) * L1;
      const double *A = basis->origin + 3 * p, *Bc = basis->origin + 3 * q;
      const int cls = la * (COSX_MAX_L + 1) + lb, stride = cosx_stride(la, lb);
      const int nca = ncart(la), ncb = ncart(lb);
      const double sph_fac = (la == 0 ? 0.282094791773878143 : (la == 1 ? 0.488602511902919921 : 1.0)) *
                             (lb == 0 ? 0.282094791773878143 : (lb == 1 ? 0.488602511902919921 : 1.0)); // kernels.h:11-18
      std::vector<double> rec_all;
      Sc.assign(static_cast<size_t>(nca) * ncb, 0.0);
      double p_min = 1.0e300, cen[3] = {0, 0, 0};
      Ex.resize(ne);
      Ey.resize(ne);
      Ez.resize(ne);
      for (int ia = 0; ia < basis->nprim[p]; ++ia)
        for (int ib = 0; ib < basis->nprim[q]; ++ib) {
          const double a = basis->exps[basis->prim_off[p] + ia], b = basis->exps[basis->prim_off[q] + ib];
          const double ca = basis->coefs[basis->prim_off[p] + ia], cb = basis->coefs[basis->prim_off[q] + ib];
          const double pe = a + b;
          const double P[3] = {(a * A[0] + b * Bc[0]) / pe, (a * A[1] + b * Bc[1]) / pe, (a * A[2] + b * Bc[2]) / pe};
          const double pref = 2.0 * kPi / pe * sph_fac * ca * cb; // kernels.h:325
          e_table(la, lb, a, b, Bc[0] - A[0], Ex.data());
          e_table(la, lb, a, b, Bc[1] - A[1], Ey.data());
          e_table(la, lb, a, b, Bc[2] - A[2], Ez.data());
          if (pe < p_min) { // esp.cpp:85-100: first minimum wins
            p_min = pe;
            std::memcpy(cen, P, sizeof(cen));
          }
          // overlap block (esp.cpp:127-146): prefactor sqrt(pi/p)/2 E_000
          const double of = pref * std::sqrt(kPi / pe) / 2.0;
          int ra = 0;
          for (int ax = la; ax >= 0; --ax)
            for (int ay = la - ax; ay >= 0; --ay, ++ra) {
              const int az = la - ax - ay;
              int rb = 0;
              for (int bx = lb; bx >= 0; --bx)
                for (int by = lb - bx; by >= 0; --by, ++rb) {
                  const int bz = lb - bx - by;
                  Sc[ra * ncb + rb] += of * Ex[(ax * (lb + 1) + bx) * L1] * Ey[(ay * (lb + 1) + by) * L1] *
                                       Ez[(az * (lb + 1) + bz) * L1];
                }
            }
          const size_t o = rec_all.size();
          rec_all.resize(o + stride);
          rec_all[o] = pe;
          rec_all[o + 1] = P[0];
          rec_all[o + 2] = P[1];
          rec_all[o + 3] = P[2];
          rec_all[o + 4] = pref;
          if (la == 0 && lb == 0) { // pre-combined: V = [4] R_000
            rec_all[o + 4] = pref * Ex[0] * Ey[0] * Ez[0];
          } else if (la == 1 && lb == 0) { // pre-combined: V_a = cA_a R_000 + cB_a R_(1 in a); E^d[(i * 1 + 0) * 2 + t], i = 0, 1
            const double *E[3] = {Ex.data(), Ey.data(), Ez.data()};
            for (int a = 0; a < 3; ++a) {
              double other = 1.0;
              for (int d = 0; d < 3; ++d)
                if (d != a) other *= E[d][0];
              rec_all[o + COSX_PRIM_HDR + a] = pref * E[a][2] * other;     // t = 0 of (i = 1, j = 0)
              rec_all[o + COSX_PRIM_HDR + 3 + a] = pref * E[a][3] * other; // t = 1
            }
          } else if (cosx_dense(la, lb)) { // dense pre-combined products, loop order of prim_eval / cosx_dense_count
            size_t k = o + COSX_PRIM_HDR;
            for (int ax = la; ax >= 0; --ax)
              for (int ay = la - ax; ay >= 0; --ay) {
                const int az = la - ax - ay;
                for (int bx = lb; bx >= 0; --bx)
                  for (int by = lb - bx; by >= 0; --by) {
                    const int bz = lb - bx - by;
                    for (int t = 0; t <= ax + bx; ++t)
                      for (int u = 0; u <= ay + by; ++u)
                        for (int v = 0; v <= az + bz; ++v)
                          rec_all[k++] = pref * Ex[(ax * (lb + 1) + bx) * L1 + t] * Ey[(ay * (lb + 1) + by) * L1 + u] *
                                         Ez[(az * (lb + 1) + bz) * L1 + v];
                  }
              }
          } else {
            for (int i = 0; i < ne; ++i) {
              rec_all[o + COSX_PRIM_HDR + i] = pref * Ex[i];
              rec_all[o + COSX_PRIM_HDR + ne + i] = Ey[i];
              rec_all[o + COSX_PRIM_HDR + 2 * ne + i] = Ez[i];
            }
          }
        }
      // block in the basis' own (Cartesian or spherical) functions
      const int np = shell_size[p], nq = shell_size[q];
      if (pure) {
        const std::vector<double> &Ta = Tfull[la], &Tb = Tfull[lb];
        tmp.assign(static_cast<size_t>(np) * ncb, 0.0);
        for (int i = 0; i < np; ++i)
          for (int k = 0; k < nca; ++k)
            for (int j = 0; j < ncb; ++j) tmp[i * ncb + j] += Ta[i * nca + k] * Sc[k * ncb + j];
        Ss.assign(static_cast<size_t>(np) * nq, 0.0);
        for (int i = 0; i < np; ++i)
          for (int j = 0; j < nq; ++j)
            for (int k = 0; k < ncb; ++k) Ss[i * nq + j] += tmp[i * ncb + k] * Tb[j * ncb + k];
      } else {
        Ss = Sc;
      }
      double n2 = 0.0;
      for (int i = 0; i < np; ++i)
        for (int j = 0; j < nq; ++j) {
          const double v = Ss[i * nq + j];
          n2 += v * v;
          h->S_analytic[(first_bf[p] + i) + static_cast<size_t>(first_bf[q] + j) * nbf] = v;
          h->S_analytic[(first_bf[q] + j) + static_cast<size_t>(first_bf[p] + i) * nbf] = v;
        }
      const bool same = A[0] == Bc[0] && A[1] == Bc[1] && A[2] == Bc[2]; // integral_engine.h:262-266
      if (!same && std::sqrt(n2) < pair_threshold) continue;
      HostPair e;
      e.la = la;
      e.lb = lb;
      e.d.p = p;
      e.d.q = q;
      e.d.bfp = first_bf[p];
      e.d.bfq = first_bf[q];
      e.d.nprim = basis->nprim[p] * basis->nprim[q];
      e.d.off = static_cast<int>(pool[cls].size());
      e.d.cx = cen[0];
      e.d.cy = cen[1];
      e.d.cz = cen[2];
      e.d.extent = std::sqrt(27.631021115928547 / p_min);
      pool[cls].insert(pool[cls].end(), rec_all.begin(), rec_all.end());
      hp[cls].push_back(e);
      h->nprim_pairs += e.d.nprim;
    }
  std::vector<CosxPair> pairs_all;
  std::vector<double> pool_all;
  for (int cls = 0; cls < NCLS; ++cls) {
    if (hp[cls].empty()) continue;
    occ_cosx_handle::ClassRange cr = {cls / (COSX_MAX_L + 1), cls % (COSX_MAX_L + 1), static_cast<int>(pairs_all.size()), static_cast<int>(hp[cls].size())};
    const size_t base = pool_all.size();
    if (base + pool[cls].size() > 0x7fffffffULL) {
      occ_cosx_destroy(h);
      return fail("occ_cosx_create: primitive pool exceeds 2^31 doubles");
    }
    std::stable_sort(hp[cls].begin(), hp[cls].end(), [](const HostPair &a, const HostPair &b) {
      return a.d.p != b.d.p ? a.d.p < b.d.p : a.d.q < b.d.q;
    });
    for (auto &e : hp[cls]) {
      e.d.off += static_cast<int>(base);
      pairs_all.push_back(e.d);
    }
    pool_all.insert(pool_all.end(), pool[cls].begin(), pool[cls].end());
    h->classes.push_back(cr);
  }
  h->npairs = pairs_all.size();

  // ---- grid batches: blocks of 128 consecutive points of every atom grid (:676-699), dealt round-robin to ranks ----
  std::vector<double> px, py, pz, pw, pwo, tc[4];
  std::vector<TileDesc> tiles;
  size_t gbatch = 0;
  const size_t base0 = atom_offsets[0];
  for (int a = 0; a < natoms; ++a) {
    const size_t n = atom_offsets[a + 1] - atom_offsets[a];
    for (size_t s0 = 0; s0 < n; s0 += TILE_P, ++gbatch) {
      if (static_cast<int>(gbatch % world) != rank) continue;
      const size_t cnt = std::min<size_t>(TILE_P, n - s0), g0 = atom_offsets[a] - base0 + s0;
      double c[3] = {0, 0, 0};
      for (size_t i = 0; i < cnt; ++i)
        for (int d = 0; d < 3; ++d) c[d] += pts_xyz[3 * (g0 + i) + d];
      for (int d = 0; d < 3; ++d) c[d] /= static_cast<double>(cnt);
      double rad = 0.0;
      for (size_t i = 0; i < cnt; ++i) {
        const double dx = pts_xyz[3 * (g0 + i)] - c[0], dy = pts_xyz[3 * (g0 + i) + 1] - c[1],
                     dz = pts_xyz[3 * (g0 + i) + 2] - c[2];
        rad = std::max(rad, std::sqrt(dx * dx + dy * dy + dz * dz));
      }
      for (int d = 0; d < 3; ++d) tc[d].push_back(c[d]);
      tc[3].push_back(rad);
      for (size_t i = 0; i < TILE_P; ++i) {
        const size_t g = g0 + (i < cnt ? i : 0);
        px.push_back(pts_xyz[3 * g]);
        py.push_back(pts_xyz[3 * g + 1]);
        pz.push_back(pts_xyz[3 * g + 2]);
        pw.push_back(i < cnt ? weights[g] : 0.0);
        // compute_overlap_matrix (:349-352) never reaches the last point of an atom grid
        pwo.push_back((i < cnt && s0 + i + 1 < n) ? weights[g] : 0.0);
      }
      TileDesc t = {};
      t.leaf = static_cast<int>(tiles.size());
      t.pt0 = static_cast<int>(tiles.size()) * TILE_P;
      t.count = static_cast<int>(cnt);
      t.nb = nbf;
      t.nbp = h->nbp;
      t.sh_off = 0;
      t.nsh = nsh;
      tiles.push_back(t);
      h->h_counts.push_back(static_cast<int>(cnt));
    }
  }
  h->ntiles = static_cast<int>(tiles.size());
  const size_t slab_bytes = static_cast<size_t>(h->nbp) * TILE_P * sizeof(double);
  h->tiles_per_chunk = static_cast<int>(std::max<size_t>(1, std::min<size_t>(batch_bytes / (3 * slab_bytes), std::max(h->ntiles, 1))));
  for (int t = 0; t < h->ntiles; ++t)
    tiles[t].scr_off = static_cast<long long>(t % h->tiles_per_chunk) * h->nbp * TILE_P;
  h->nruns = (nsh + 15) / 16;
  std::vector<CollocItem> items;
  for (int t = 0; t < h->ntiles; ++t)
    for (int r = 0; r < h->nruns; ++r) items.push_back({t, r * 16, std::min(nsh, (r + 1) * 16), r == h->nruns - 1 ? 1 : 0});

  // ---- device state ----
  CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess && fail("occ_cosx_create: stream"));
  for (auto &e : h->ev) CHECK(cudaEventCreate(&e) != cudaSuccess && fail("occ_cosx_create: event"));
  upload_default_c2s();
  {
    int *d_l = nullptr, *d_np = nullptr, *d_po = nullptr, *d_fb = nullptr, *d_list = nullptr;
    double *d_or = nullptr, *d_ex = nullptr, *d_co = nullptr;
    TileDesc *d_tiles = nullptr;
    size_t nprim_tot = 0;
    for (int s = 0; s < nsh; ++s) nprim_tot = std::max<size_t>(nprim_tot, basis->prim_off[s] + basis->nprim[s]);
    std::vector<int> list(nsh);
    for (int s = 0; s < nsh; ++s) list[s] = s;
    CHECK(dev_upload(h, &d_l, std::vector<int>(basis->l, basis->l + nsh)));
    CHECK(dev_upload(h, &d_np, std::vector<int>(basis->nprim, basis->nprim + nsh)));
    CHECK(dev_upload(h, &d_po, std::vector<int>(basis->prim_off, basis->prim_off + nsh)));
    CHECK(dev_upload(h, &d_fb, first_bf));
    CHECK(dev_upload(h, &d_list, list));
    CHECK(dev_upload(h, &d_or, std::vector<double>(basis->origin, basis->origin + 3 * nsh)));
    CHECK(dev_upload(h, &d_ex, std::vector<double>(basis->exps, basis->exps + nprim_tot)));
    CHECK(dev_upload(h, &d_co, std::vector<double>(basis->coefs, basis->coefs + nprim_tot)));
    CHECK(dev_upload(h, &d_tiles, tiles));
    h->B.basis = {d_l, d_np, d_po, d_fb, d_or, d_ex, d_co, pure};
    h->B.tiles = d_tiles;
    h->B.shell_list = d_list;
    h->B.shell_col = d_fb;
    h->B.nbf = nbf;
    h->B.nspin = 1;
    h->B.deriv = 0;
    h->B.ncomp = 1;
  }
  {
    double *d = nullptr;
    CHECK(dev_upload(h, &d, px));
    h->B.px = d;
    CHECK(dev_upload(h, &d, py));
    h->B.py = d;
    CHECK(dev_upload(h, &d, pz));
    h->B.pz = d;
    CHECK(dev_upload(h, &d, pw));
    h->B.pw = d;
  }
  CHECK(dev_upload(h, &h->d_pw_ovl, pwo));
  for (int d = 0; d < 4; ++d) CHECK(dev_upload(h, &h->d_tc[d], tc[d]));
  CHECK(dev_upload(h, &h->d_items, items));
  CHECK(dev_upload(h, &h->d_counts, h->h_counts));
  CHECK(dev_alloc(h, &h->d_skip, h->ntiles));
  CHECK(dev_upload(h, &h->d_shell_size, shell_size));
  CHECK(dev_upload(h, &h->d_pairs, pairs_all));
  CHECK(dev_upload(h, &h->d_pool, pool_all));
  { // device copy in order-major layout [m][segment][8]: a class (la, lb) only ever reads the rows of order m = la + lb
    // (lower orders come from the downward recursion), so its 253 x 64 B of table are one contiguous 16 KB region that
    // stays L1-resident next to the kernels' shared memory instead of 64-byte pieces strided by 1 KiB
    const std::vector<double> seg_major = boys_table();
    constexpr int N = 8, MaxM = 16;
    const size_t nseg = seg_major.size() / (N * MaxM);
    static_assert(MaxM == COSX_BOYS_ORDERS, "device table layout");
    std::vector<double> order_major(seg_major.size() + COSX_BOYS_SEGMENTS);
    for (int sg = 0; sg < COSX_BOYS_SEGMENTS; ++sg) // e^-T at the segment midpoints (k_cosx_esp.cuh add_seeds)
      order_major[seg_major.size() + sg] = std::exp(-(sg + 0.5) / 7.0);
    for (size_t sg = 0; sg < nseg; ++sg)
      for (int m = 0; m < MaxM; ++m)
        std::memcpy(&order_major[(m * nseg + sg) * N], &seg_major[(sg * MaxM + m) * N], N * sizeof(double));
    if (nseg != static_cast<size_t>(COSX_BOYS_SEGMENTS)) {
      occ_cosx_destroy(h);
      return fail("occ_cosx_create: Boys table size mismatch");
    }
    CHECK(dev_upload(h, &h->d_boys, order_major));
  }
  {
    std::vector<double> c2s(9 + 30 + 70 + 135, 0.0);
    const double p1[9] = {0, 1, 0, 0, 0, 1, 1, 0, 0}; // gto.cpp:296-305: (py, pz, px)
    std::copy(p1, p1 + 9, c2s.begin());
    std::copy(Tfull[2].begin(), Tfull[2].end(), c2s.begin() + 9);
    std::copy(Tfull[3].begin(), Tfull[3].end(), c2s.begin() + 39);
    std::copy(Tfull[4].begin(), Tfull[4].end(), c2s.begin() + 109);
    CHECK(dev_upload(h, &h->d_c2s, c2s));
  }
  CHECK(dev_alloc(h, &h->d_active, h->npairs));
  CHECK(dev_alloc(h, &h->d_stats, 64)); // 2 per (la, lb) class (25 classes), [56] = k_cosx_kgemm blocks
  const size_t slab_d = static_cast<size_t>(h->nbp) * TILE_P * h->tiles_per_chunk;
  CHECK(dev_alloc(h, &h->d_phi, slab_d));
  CHECK(dev_alloc(h, &h->d_F, slab_d));
  CHECK(dev_alloc(h, &h->d_G, slab_d));
  const size_t n2 = static_cast<size_t>(nbf) * nbf;
  CHECK(dev_alloc(h, &h->d_D, 2 * n2));
  CHECK(dev_alloc(h, &h->d_D2q, n2));
  CHECK(dev_alloc(h, &h->d_Kraw, n2));
  CHECK(dev_alloc(h, &h->d_K, 2 * n2));
  CHECK(dev_alloc(h, &h->d_proj, n2));
  h->B.scratch = h->d_phi;
  *out = h;
  if (occ_cosx_set_projector(h, overlap_projector)) {
    *out = nullptr;
    occ_cosx_destroy(h);
    return 1;
  }
#undef CHECK
  h->info.nbf = nbf;
  h->info.nsh = nsh;
  h->info.natoms = natoms;
  h->info.npts = h->npts;
  h->info.nbatches = h->ntiles;
  h->info.npairs = h->npairs;
  h->info.nprimitive_pairs = h->nprim_pairs;
  h->info.device_bytes = h->alloc_bytes;
  return 0;
}

int occ_cosx_set_projector(occ_cosx_handle *h, const double *P) {
  if (!h) return fail("occ_cosx_set_projector: null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t n = h->nbf;
  std::vector<double> I;
  if (!P) {
    I.assign(n * n, 0.0);
    for (size_t i = 0; i < n; ++i) I[i + i * n] = 1.0;
    P = I.data();
  }
  CUDA_TRY(cudaMemcpy(h->d_proj, P, n * n * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

// seminumerical_exchange_sweep (:660-815) for one projected density already on the device
// Cq / Cs (nbf x nocc each, or NULL): the factorised form D2q = Cq . Cs^T of the same matrix — F_t = (Phi_t Cq) Cs^T then
// costs 4 n nbf nocc flops instead of 2 n nbf^2; D2q itself is still read by the density screen of the pairs (:758-764)
static int sweep(occ_cosx_handle *h, const double *D2q, double precision, double w_full, double w_lr, double omega,
                 double *Kraw, const double *Cq = nullptr, const double *Cs = nullptr, int nocc = 0) {
  const int nbf = h->nbf, nbp = h->nbp;
  cudaStream_t st = h->stream;
  const size_t n2 = static_cast<size_t>(nbf) * nbf;
  CUDA_TRY(cudaMemsetAsync(Kraw, 0, n2 * sizeof(double), st));
  launch_cosx_pair_flags(h->d_pairs, h->d_shell_size, static_cast<int>(h->npairs), D2q, nbf, h->density_threshold,
                         h->d_active, st);
  h->info.launches += 1;
  if (h->collect) CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, 64 * sizeof(unsigned long long), st));
  CosxParams P = {};
  P.pairs = h->d_pairs;
  P.pool = h->d_pool;
  P.pair_active = h->d_active;
  P.tile_skip = h->d_skip;
  P.counts = h->d_counts;
  P.tcx = h->d_tc[0];
  P.tcy = h->d_tc[1];
  P.tcz = h->d_tc[2];
  P.trad = h->d_tc[3];
  P.px = h->B.px;
  P.py = h->B.py;
  P.pz = h->B.pz;
  P.pw = h->B.pw;
  P.F = h->d_F;
  P.G = h->d_G;
  P.boys = h->d_boys;
  P.c2s = h->d_c2s;
  P.nbp = nbp;
  P.pure = h->pure;
  P.margin = h->margin;
  P.w_full = w_full;
  P.w_lr = w_lr;
  P.omega = omega;
  {
    const char *dbg = std::getenv("OCC_COSX_DEBUG");
    P.debug = dbg ? std::atoi(dbg) : 0;
  }
  P.stats = h->collect ? h->d_stats : nullptr;
  const long long slab = static_cast<long long>(nbp) * TILE_P;
  float ms[4] = {0, 0, 0, 0};
  for (int t0 = 0; t0 < h->ntiles; t0 += h->tiles_per_chunk) {
    const int nt = std::min(h->tiles_per_chunk, h->ntiles - t0);
    if (h->collect) cudaEventRecord(h->ev[0], st);
    launch_collocate(h->B, h->d_items + static_cast<size_t>(t0) * h->nruns, nt * h->nruns, st);
    launch_cosx_tile_flags(h->d_phi, h->d_counts + t0, nt, nbf, nbp, precision, h->d_skip + t0, st);
    h->info.launches += 2;
    if (h->collect) cudaEventRecord(h->ev[1], st);
    // F_t = Phi_t . D2q for every batch of the chunk (:733-734, the weight is applied when F is read)
    if (Cq && nocc > 0) {
      const long long pslab = static_cast<long long>(nocc) * TILE_P;
      // Psi_t = Phi_t (P C), F_t = Psi_t C^T: this library's DMMA GEMM (k_dgemm.cu), one batched launch each
      launch_dgemm(false, TILE_P, nocc, nbf, 1.0, h->d_phi, TILE_P, slab, Cq, nbf, 0, h->d_psi, TILE_P, pslab, nt, st);
      launch_dgemm(true, TILE_P, nbf, nocc, 1.0, h->d_psi, TILE_P, pslab, Cs, nbf, 0, h->d_F, TILE_P, slab, nt, st);
      h->info.launches += 2;
    } else if (Cq) { // no occupied orbital of this spin: D2q = 0
      CUDA_TRY(cudaMemsetAsync(h->d_F, 0, static_cast<size_t>(nt) * slab * sizeof(double), st));
    } else {
      launch_dgemm(false, TILE_P, nbf, nbf, 1.0, h->d_phi, TILE_P, slab, D2q, nbf, 0, h->d_F, TILE_P, slab, nt, st);
      h->info.launches += 1;
    }
    CUDA_TRY(cudaMemsetAsync(h->d_G, 0, static_cast<size_t>(nt) * slab * sizeof(double), st));
    if (h->collect) cudaEventRecord(h->ev[2], st);
    P.tile0 = t0;
    for (const auto &c : h->classes) {
      launch_cosx_esp(P, c.la, c.lb, c.pair0, c.npairs, nt, st);
      h->info.launches += 1;
    }
    if (h->collect) cudaEventRecord(h->ev[3], st);
    // K' += Phi_t^T . G_t (:797; G here is the reference's -Gg): k_cosx_kgemm, one launch per chunk
    launch_cosx_kgemm(h->d_phi, h->d_G, h->d_skip + t0, nt, nbf, nbp, Kraw, h->collect ? h->d_stats + 56 : nullptr, st);
    h->info.launches += 1;
    if (h->collect) {
      cudaEventRecord(h->ev[4], st);
      CUDA_TRY(cudaEventSynchronize(h->ev[4]));
      for (int i = 0; i < 4; ++i) {
        float m = 0;
        cudaEventElapsedTime(&m, h->ev[i], h->ev[i + 1]);
        ms[i] += m;
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  if (h->collect) {
    unsigned long long st_h[64];
    CUDA_TRY(cudaMemcpyAsync(st_h, h->d_stats, sizeof(st_h), cudaMemcpyDeviceToHost, st));
    int skip_n = 0;
    std::vector<int> sk(h->ntiles);
    CUDA_TRY(cudaMemcpyAsync(sk.data(), h->d_skip, sizeof(int) * h->ntiles, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int v : sk) skip_n += v;
    h->info.skipped_batches += skip_n;
    for (int cls = 0; cls < (COSX_MAX_L + 1) * (COSX_MAX_L + 1); ++cls) {
      const int la = cls / (COSX_MAX_L + 1), lb = cls % (COSX_MAX_L + 1), L = la + lb;
      const double nherm = (L + 1) * (L + 2) * (L + 3) / 6.0, nab = static_cast<double>(nfun(la, h->pure)) * nfun(lb, h->pure);
      h->info.pair_evals += st_h[2 * cls];
      h->info.dense_flops += 2.0 * nherm * nab * static_cast<double>(st_h[2 * cls + 1]);
    }
    h->info.kgemm_blocks += st_h[56];
    h->info.ms_collocate += ms[0];
    h->info.ms_fgemm += ms[1];
    h->info.ms_esp += ms[2];
    h->info.ms_kgemm += ms[3];
  }
  double npt_local = 0.0;
  for (int c : h->h_counts) npt_local += c;
  h->info.gemm_flops += 2.0 * npt_local * nbf * nbf + (Cq ? 4.0 * npt_local * nbf * nocc : 2.0 * npt_local * nbf * nbf);
  return 0;
}

static void reset_info(occ_cosx_handle *h) {
  h->info.pair_evals = 0;
  h->info.skipped_batches = 0;
  h->info.dense_flops = h->info.gemm_flops = 0.0;
  h->info.ms_collocate = h->info.ms_fgemm = h->info.ms_esp = h->info.ms_kgemm = 0.0;
  h->info.launches = 0;
  h->info.gemm_calls = 0;
  h->info.kgemm_blocks = 0;
}

int occ_cosx_K_for_density_device(occ_cosx_handle *h, const double *D2q_dev, double precision, double w_full,
                                  double w_lr, double omega, double *K_dev) {
  if (!h || !D2q_dev || !K_dev) return fail("occ_cosx_K_for_density_device: null argument");
  if (w_lr != 0.0 && !(omega > 0.0)) return fail("occ_cosx: the long-range operator needs omega > 0");
  CUDA_TRY(cudaSetDevice(h->device));
  reset_info(h);
  if (sweep(h, D2q_dev, precision, w_full, w_lr, omega, K_dev)) return 1;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int occ_cosx_K_for_density(occ_cosx_handle *h, const double *D2q, double precision, double omega, double *K) {
  if (!h || !D2q || !K) return fail("occ_cosx_K_for_density: null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t n2 = static_cast<size_t>(h->nbf) * h->nbf;
  CUDA_TRY(cudaMemcpyAsync(h->d_D2q, D2q, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  reset_info(h);
  if (sweep(h, h->d_D2q, precision, omega > 0.0 ? 0.0 : 1.0, omega > 0.0 ? 1.0 : 0.0, omega, h->d_Kraw)) return 1;
  CUDA_TRY(cudaMemcpyAsync(K, h->d_Kraw, n2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

// compute_K_restricted / _unrestricted / _general (:517-545, 617-660, 662-712) and compute_K_range_separated (:830-860)
// the build up to the symmetrised K on the device (*dK_out); nothing is synchronised
static int compute_K_device(occ_cosx_handle *h, int kind, const double *D, double precision, double w_full, double w_lr,
                            double omega, double **dK_out, const double *C = nullptr, int nocc_a = 0, int nocc_b = 0) {
  if (!h || !D) return fail("occ_cosx_compute_K: null argument");
  if (C && kind == 2) return fail("occ_cosx_compute_K_mo: restricted (0) or unrestricted (1) only");
  if (C && (nocc_a < 0 || nocc_b < 0 || nocc_a > h->nbf || nocc_b > h->nbf))
    return fail("occ_cosx_compute_K_mo: occupied counts outside [0, nbf]");
  if (kind < 0 || kind > 2)
    return fail("occ_cosx_compute_K: kind must be 0 (restricted), 1 (unrestricted) or 2 (general)");
  CUDA_TRY(cudaSetDevice(h->device));
  const int nbf = h->nbf;
  const int rows = kind == 0 ? nbf : 2 * nbf, cols = kind == 2 ? 2 * nbf : nbf; // qm::matrix_dimensions
  const size_t n2 = static_cast<size_t>(nbf) * nbf;
  cudaStream_t st = h->stream;
  if (kind == 2 && !h->d_Dg) { // general: 2 nbf x 2 nbf in / out, four raw K' blocks
    if (dev_alloc(h, &h->d_Dg, 4 * n2) || dev_alloc(h, &h->d_Kg, 4 * n2) || dev_alloc(h, &h->d_Kraw4, 4 * n2)) return 1;
  }
  double *dD = kind == 2 ? h->d_Dg : h->d_D, *dK = kind == 2 ? h->d_Kg : h->d_K;
  CUDA_TRY(cudaMemcpyAsync(dD, D, static_cast<size_t>(rows) * cols * sizeof(double), cudaMemcpyHostToDevice, st));
  reset_info(h);
  const int nocc_s[2] = {nocc_a, kind == 1 ? nocc_b : 0};
  if (C) { // orbitals on the device: [C_alpha | C_beta], nbf rows
    const int cap = std::max(1, std::max(nocc_s[0], nocc_s[1]));
    if (cap > h->mo_cap) {
      CUDA_TRY(cudaStreamSynchronize(st)); // nothing in flight may still read the buffers that are about to go
      dev_free(h, &h->d_C);
      dev_free(h, &h->d_Cq);
      dev_free(h, &h->d_psi);
      if (dev_alloc(h, &h->d_C, 2 * static_cast<size_t>(nbf) * cap) || dev_alloc(h, &h->d_Cq, static_cast<size_t>(nbf) * cap) ||
          dev_alloc(h, &h->d_psi, static_cast<size_t>(h->tiles_per_chunk) * TILE_P * cap))
        return 1;
      h->mo_cap = cap;
    }
    const size_t ncol = static_cast<size_t>(nocc_s[0]) + nocc_s[1];
    if (ncol) CUDA_TRY(cudaMemcpyAsync(h->d_C, C, ncol * nbf * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  // Cq = P . C_s, so that Cq C_s^T = P (2 D_s)
  auto project_orbitals = [&](const double *Cs, int nocc) -> int {
    if (nocc == 0) return 0;
    launch_dgemm(false, nbf, nocc, nbf, 1.0, h->d_proj, nbf, 0, Cs, nbf, 0, h->d_Cq, nbf, 0, 1, st);
    h->info.launches += 1;
    return 0;
  };
  // D2q = P . (2 D_block): the factor 2 is D2 = 2 D for restricted (:524) and the stored 1/2 otherwise (:632-633, 676-679)
  auto project = [&](const double *Dblk) -> int {
    launch_dgemm(false, nbf, nbf, nbf, 2.0, h->d_proj, nbf, 0, Dblk, rows, 0, h->d_D2q, nbf, 0, 1, st);
    h->info.launches += 1;
    return 0;
  };
  if (kind < 2) {
    for (int s = 0; s <= kind; ++s) {
      if (project(dD + s * nbf)) return 1;
      const double *Cs = C ? h->d_C + static_cast<size_t>(s) * nocc_s[0] * nbf : nullptr;
      if (C && project_orbitals(Cs, nocc_s[s])) return 1;
      if (sweep(h, h->d_D2q, precision, w_full, w_lr, omega, h->d_Kraw, C ? h->d_Cq : nullptr, Cs, nocc_s[s])) return 1;
      launch_cosx_symmetrize(h->d_Kraw, h->d_Kraw, nbf, kind == 0 ? 0.25 : 0.5, dK + s * nbf, rows, st);
      h->info.launches += 1;
    }
  } else {
    // blocks aa, ab, ba, bb of the column-major 2 nbf x 2 nbf matrix as qm::block defines them (include/occ/qm/opmatrix.h:
    // 14-28): ab is the BOTTOM-LEFT block (rows/2, 0), ba the TOP-RIGHT one (0, cols/2)
    const size_t off[4] = {0, static_cast<size_t>(nbf), static_cast<size_t>(nbf) * rows,
                           static_cast<size_t>(nbf) * rows + nbf};
    for (int blk = 0; blk < 4; ++blk) {
      if (project(dD + off[blk])) return 1;
      if (sweep(h, h->d_D2q, precision, w_full, w_lr, omega, h->d_Kraw4 + blk * n2)) return 1;
    }
    const double *Kaa = h->d_Kraw4, *Kab = h->d_Kraw4 + n2, *Kba = h->d_Kraw4 + 2 * n2, *Kbb = h->d_Kraw4 + 3 * n2;
    launch_cosx_symmetrize(Kaa, Kaa, nbf, 0.5, dK + off[0], rows, st); // :697-698
    launch_cosx_symmetrize(Kab, Kba, nbf, 0.5, dK + off[1], rows, st); // :700  Kab_sym = 0.5 (Kab + Kba^T)
    launch_cosx_symmetrize(Kba, Kab, nbf, 0.5, dK + off[2], rows, st); // :701  Kba_sym = 0.5 (Kba + Kab^T)
    launch_cosx_symmetrize(Kbb, Kbb, nbf, 0.5, dK + off[3], rows, st);
    h->info.launches += 4;
  }
  *dK_out = dK;
  return 0;
}

static int compute_K_impl(occ_cosx_handle *h, int kind, const double *D, double precision, double w_full, double w_lr,
                          double omega, double *K, const double *C = nullptr, int nocc_a = 0, int nocc_b = 0) {
  if (!K) return fail("occ_cosx_compute_K: null argument");
  double *dK = nullptr;
  if (compute_K_device(h, kind, D, precision, w_full, w_lr, omega, &dK, C, nocc_a, nocc_b)) return 1;
  const int nbf = h->nbf;
  const size_t count = static_cast<size_t>(kind == 0 ? nbf : 2 * nbf) * (kind == 2 ? 2 * nbf : nbf);
  CUDA_TRY(cudaMemcpyAsync(K, dK, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int occ_cosx_compute_K(occ_cosx_handle *h, int kind, const double *D, double precision, double omega, double *K) {
  if (omega < 0.0) return fail("occ_cosx_compute_K: omega must be >= 0");
  return compute_K_impl(h, kind, D, precision, omega > 0.0 ? 0.0 : 1.0, omega > 0.0 ? 1.0 : 0.0, omega, K);
}

int occ_cosx_compute_K_range_separated(occ_cosx_handle *h, const double *D, double alpha, double beta, double omega,
                                       double precision, double *K) {
  if (!(omega > 0.0)) return fail("occ_cosx_compute_K_range_separated: omega must be > 0");
  return compute_K_impl(h, 0, D, precision, alpha + beta, -beta, omega, K);
}

int occ_cosx_compute_K_range_separated_kind(occ_cosx_handle *h, int kind, const double *D, double alpha, double beta,
                                            double omega, double precision, double *K) {
  if (!(omega > 0.0)) return fail("occ_cosx_compute_K_range_separated: omega must be > 0");
  if (kind != 0 && kind != 1) // seminumerical_exchange.cpp:862-893 handles Restricted and Unrestricted only
    return fail("occ_cosx_compute_K_range_separated: not implemented for general spinorbitals");
  return compute_K_impl(h, kind, D, precision, alpha + beta, -beta, omega, K);
}

int occ_cosx_compute_K_range_separated_mo_kind(occ_cosx_handle *h, int kind, const double *D, const double *C,
                                               int nocc_alpha, int nocc_beta, double alpha, double beta, double omega,
                                               double precision, double *K) {
  if (!C) return fail("occ_cosx_compute_K_range_separated_mo: null argument");
  if (!(omega > 0.0)) return fail("occ_cosx_compute_K_range_separated_mo: omega must be > 0");
  if (kind != 0 && kind != 1) return fail("occ_cosx_compute_K_range_separated_mo: not implemented for general spinorbitals");
  return compute_K_impl(h, kind, D, precision, alpha + beta, -beta, omega, K, C, nocc_alpha, kind == 1 ? nocc_beta : 0);
}

int occ_cosx_compute_K_mo(occ_cosx_handle *h, int kind, const double *D, const double *C, int nocc_alpha, int nocc_beta,
                          double precision, double omega, double *K) {
  if (!C) return fail("occ_cosx_compute_K_mo: null argument");
  if (omega < 0.0) return fail("occ_cosx_compute_K_mo: omega must be >= 0");
  return compute_K_impl(h, kind, D, precision, omega > 0.0 ? 0.0 : 1.0, omega > 0.0 ? 1.0 : 0.0, omega, K, C, nocc_alpha,
                        nocc_beta);
}

int occ_cosx_compute_K_range_separated_mo(occ_cosx_handle *h, const double *D, const double *C, int nocc, double alpha,
                                          double beta, double omega, double precision, double *K) {
  if (!C) return fail("occ_cosx_compute_K_range_separated_mo: null argument");
  if (!(omega > 0.0)) return fail("occ_cosx_compute_K_range_separated_mo: omega must be > 0");
  return compute_K_impl(h, 0, D, precision, alpha + beta, -beta, omega, K, C, nocc, 0);
}

int occ_cosx_analytic_overlap(occ_cosx_handle *h, double *S) {
  if (!h || !S) return fail("occ_cosx_analytic_overlap: null argument");
  std::memcpy(S, h->S_analytic.data(), h->S_analytic.size() * sizeof(double));
  return 0;
}

// compute_overlap_matrix (:336-380): Sn = sum_batches Phi^T (w Phi), the last point of every atom grid left out
int occ_cosx_numerical_overlap(occ_cosx_handle *h, double *Sn) {
  if (!h || !Sn) return fail("occ_cosx_numerical_overlap: null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  const int nbf = h->nbf, nbp = h->nbp;
  const size_t n2 = static_cast<size_t>(nbf) * nbf;
  cudaStream_t st = h->stream;
  CUDA_TRY(cudaMemsetAsync(h->d_Kraw, 0, n2 * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(h->d_skip, 0, sizeof(int) * h->ntiles, st)); // no batch is skipped here (:352-377)
  for (int t0 = 0; t0 < h->ntiles; t0 += h->tiles_per_chunk) {
    const int nt = std::min(h->tiles_per_chunk, h->ntiles - t0);
    launch_collocate(h->B, h->d_items + static_cast<size_t>(t0) * h->nruns, nt * h->nruns, st);
    launch_cosx_scale_weights(h->d_F, h->d_phi, h->d_pw_ovl + static_cast<size_t>(t0) * TILE_P, nt, nbf, nbp, st);
    launch_cosx_kgemm(h->d_phi, h->d_F, h->d_skip + t0, nt, nbf, nbp, h->d_Kraw, nullptr, st);
  }
  CUDA_TRY(cudaMemcpyAsync(Sn, h->d_Kraw, n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int occ_cosx_get_info(const occ_cosx_handle *h, occ_cosx_info_t *info) {
  if (!h || !info) return fail("occ_cosx_get_info: null argument");
  *info = h->info;
  return 0;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Several GPUs of one box driven from ONE host process (the layout of occ_xc_multi_*, csrc/xc_multi.cu): one handle per
// device with the grid batches dealt round-robin (options.rank / world_size), a host thread per device for the call,
// every GPU uploads D (and C) over its own PCIe link, and GPU r sums its contiguous slice of the symmetrised partial K
// matrices through peer loads in rank order and writes the slice to the host matrix.
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr int COSX_MAX_DEV = 16;
struct CosxPeers {
  const double *p[COSX_MAX_DEV];
  int n;
};
__global__ void k_cosx_reduce_slice(CosxPeers parts, double *__restrict__ out, size_t begin, size_t end) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = begin + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += stride) {
    double s = parts.p[0][i];
    for (int r = 1; r < parts.n; ++r) s += parts.p[r][i];
    out[i] = s;
  }
}
// run job(r) for every rank on its own thread; first error message wins
template <class F> int for_each_rank(int n, F job) {
  std::vector<int> rc(n, 0);
  std::vector<std::string> err(n);
  std::vector<std::thread> th;
  for (int r = 0; r < n; ++r)
    th.emplace_back([&, r] {
      rc[r] = job(r);
      if (rc[r]) err[r] = occ_xc_last_error();
    });
  for (auto &t : th) t.join();
  for (int r = 0; r < n; ++r)
    if (rc[r]) return occ_xc_set_error(err[r].c_str());
  return 0;
}
} // namespace

struct occ_cosx_multi {
  int n = 0;
  int device[COSX_MAX_DEV];
  occ_cosx_handle *h[COSX_MAX_DEV];
  double *d_red[COSX_MAX_DEV]; // reduced slice (full-size buffer; only the rank's slice is written)
  size_t red_count = 0;
};

extern "C" {

void occ_cosx_multi_destroy(occ_cosx_multi *m) {
  if (!m) return;
  for (int r = 0; r < m->n; ++r) {
    if (m->d_red[r]) {
      cudaSetDevice(m->device[r]);
      cudaFree(m->d_red[r]);
    }
    occ_cosx_destroy(m->h[r]);
  }
  delete m;
}

int occ_cosx_multi_create(const occ_xc_basis_t *basis, int natoms, const size_t *atom_offsets, const double *pts_xyz,
                          const double *weights, const double *overlap_projector, const occ_cosx_options_t *options,
                          int ndev, const int *devices, occ_cosx_multi **out) {
  if (!out) return fail("occ_cosx_multi_create: null argument");
  if (ndev < 1 || ndev > COSX_MAX_DEV) return fail("occ_cosx_multi_create: between 1 and 16 devices");
  int avail = 0;
  if (cudaGetDeviceCount(&avail) != cudaSuccess || avail == 0)
    return fail("occ_cosx_multi_create: no CUDA device available (this library has no CPU fallback)");
  occ_cosx_multi *m = new occ_cosx_multi();
  m->n = ndev;
  for (int r = 0; r < COSX_MAX_DEV; ++r) {
    m->h[r] = nullptr;
    m->d_red[r] = nullptr;
    m->device[r] = 0;
  }
  for (int r = 0; r < ndev; ++r) {
    m->device[r] = devices ? devices[r] : r;
    if (m->device[r] < 0 || m->device[r] >= avail) {
      occ_cosx_multi_destroy(m);
      return fail("occ_cosx_multi_create: device ordinal out of range");
    }
  }
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      if (m->device[a] == m->device[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->device[a], m->device[b]);
      if (!can) {
        occ_cosx_multi_destroy(m);
        return fail("occ_cosx_multi_create: the devices cannot access each other's memory; use one handle per process "
                    "with options.rank / world_size and sum the partial matrices instead");
      }
      cudaSetDevice(m->device[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->device[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        occ_cosx_multi_destroy(m);
        return fail(std::string("occ_cosx_multi_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      }
      cudaGetLastError();
    }
  const int rc = for_each_rank(ndev, [&](int r) -> int {
    occ_cosx_options_t opt;
    if (options) opt = *options;
    else std::memset(&opt, 0, sizeof(opt));
    opt.device = m->device[r];
    opt.rank = r;
    opt.world_size = ndev;
    return occ_cosx_create(basis, natoms, atom_offsets, pts_xyz, weights, overlap_projector, &opt, &m->h[r]);
  });
  if (rc) {
    occ_cosx_multi_destroy(m);
    return 1;
  }
  *out = m;
  return 0;
}

int occ_cosx_multi_ndev(const occ_cosx_multi *m) { return m ? m->n : 0; }

int occ_cosx_multi_get_info(const occ_cosx_multi *m, int rank, occ_cosx_info_t *info) {
  if (!m || rank < 0 || rank >= m->n) return fail("occ_cosx_multi_get_info: bad argument");
  return occ_cosx_get_info(m->h[rank], info);
}

int occ_cosx_multi_set_projector(occ_cosx_multi *m, const double *overlap_projector) {
  if (!m) return fail("occ_cosx_multi_set_projector: null argument");
  for (int r = 0; r < m->n; ++r)
    if (occ_cosx_set_projector(m->h[r], overlap_projector)) return 1;
  return 0;
}

int occ_cosx_multi_analytic_overlap(occ_cosx_multi *m, double *S) {
  if (!m) return fail("occ_cosx_multi_analytic_overlap: null argument");
  return occ_cosx_analytic_overlap(m->h[0], S);
}

// the shards' partial numerical overlaps summed on the host in rank order (a once-per-geometry nbf x nbf matrix)
int occ_cosx_multi_numerical_overlap(occ_cosx_multi *m, double *Sn) {
  if (!m || !Sn) return fail("occ_cosx_multi_numerical_overlap: null argument");
  const size_t n2 = static_cast<size_t>(m->h[0]->nbf) * m->h[0]->nbf;
  std::vector<std::vector<double>> part(m->n, std::vector<double>(n2));
  if (for_each_rank(m->n, [&](int r) -> int { return occ_cosx_numerical_overlap(m->h[r], part[r].data()); })) return 1;
  std::copy(part[0].begin(), part[0].end(), Sn);
  for (int r = 1; r < m->n; ++r)
    for (size_t i = 0; i < n2; ++i) Sn[i] += part[r][i];
  return 0;
}

// K = w_full K(1/r) + w_lr K(erf(omega r)/r) over all devices; C may be NULL (D form)
static int multi_compute(occ_cosx_multi *m, int kind, const double *D, const double *C, int nocc_a, int nocc_b,
                         double precision, double w_full, double w_lr, double omega, double *K) {
  if (!m || !D || !K) return fail("occ_cosx_multi_compute_K: null argument");
  const int n = m->n, nbf = m->h[0]->nbf;
  const size_t count = static_cast<size_t>(kind == 0 ? nbf : 2 * nbf) * (kind == 2 ? 2 * nbf : nbf);
  std::vector<double *> dK(n, nullptr);
  // stage 1: every rank builds its symmetrised partial K on its device
  if (for_each_rank(n, [&](int r) -> int {
        occ_cosx_handle *h = m->h[r];
        if (compute_K_device(h, kind, D, precision, w_full, w_lr, omega, &dK[r], C, nocc_a, nocc_b)) return 1;
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaGetLastError());
        return 0;
      }))
    return 1;
  if (m->red_count < count) {
    for (int r = 0; r < n; ++r) {
      CUDA_TRY(cudaSetDevice(m->device[r]));
      if (m->d_red[r]) cudaFree(m->d_red[r]);
      m->d_red[r] = nullptr;
      CUDA_TRY(cudaMalloc(&m->d_red[r], count * sizeof(double)));
    }
    m->red_count = count;
  }
  // stage 2: rank r sums slice r over all partial matrices (peer loads, rank order) and writes it to the host matrix
  return for_each_rank(n, [&](int r) -> int {
    occ_cosx_handle *h = m->h[r];
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t per = ((count + n - 1) / n + 1) & ~size_t(1);
    const size_t lo = std::min(count, per * r), hi = std::min(count, per * (r + 1));
    if (hi > lo) {
      CosxPeers parts;
      parts.n = n;
      for (int p = 0; p < n; ++p) parts.p[p] = dK[p];
      const unsigned grid = static_cast<unsigned>(std::min<size_t>((hi - lo + 255) / 256, 148 * 8));
      k_cosx_reduce_slice<<<grid, 256, 0, h->stream>>>(parts, m->d_red[r], lo, hi);
      CUDA_TRY(cudaMemcpyAsync(K + lo, m->d_red[r] + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
  });
}

int occ_cosx_multi_compute_K(occ_cosx_multi *m, int kind, const double *D, const double *C, int nocc_alpha, int nocc_beta,
                             double precision, double omega, double *K) {
  if (omega < 0.0) return fail("occ_cosx_multi_compute_K: omega must be >= 0");
  return multi_compute(m, kind, D, C, nocc_alpha, nocc_beta, precision, omega > 0.0 ? 0.0 : 1.0, omega > 0.0 ? 1.0 : 0.0, omega, K);
}

int occ_cosx_multi_compute_K_range_separated(occ_cosx_multi *m, int kind, const double *D, const double *C, int nocc_alpha,
                                             int nocc_beta, double alpha, double beta, double omega, double precision,
                                             double *K) {
  if (kind == 2) return fail("occ_cosx_multi_compute_K_range_separated: restricted (0) or unrestricted (1) only");
  if (omega <= 0.0) return fail("occ_cosx_multi_compute_K_range_separated: omega must be > 0");
  return multi_compute(m, kind, D, C, nocc_alpha, nocc_beta, precision, alpha + beta, -beta, omega, K);
}

} // extern "C"
