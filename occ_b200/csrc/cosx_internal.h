// cosx_internal.h — device-side layout of the B200 seminumerical exchange (COSX) build.
//
//   batch = tile = <=128 consecutive points of one atom grid (seminumerical_exchange.cpp:676-699)
//   Phi / F / G of a tile are column-major 128 x nbp slabs (same layout as the XC build's AO slabs):
//     Phi[col*128 + p] AO values, F = Phi . D2q (unweighted; the quadrature weight is applied when F is read),
//     G = sum over shell pairs of (w F) . V  — the reference's -Gg (seminumerical_exchange.cpp:787-795)
//   shell pairs are grouped by class (la, lb); each class is one launch of an ESP kernel instantiated for it.
#pragma once
#include "xc_internal.h"

namespace xcb {

constexpr int COSX_MAX_L = 4;
constexpr int COSX_TILE_GROUP = 8; // batches whose F / G slabs share the L2 while all their pair runs execute

struct CosxPair {
  int p, q;       // shells, p >= q (registration order of seminumerical_exchange.cpp:293-331)
  int bfp, bfq;   // first basis function of p / q
  int nprim;      // primitive pairs
  int off;        // first double of this pair in the primitive pool
  double cx, cy, cz, extent; // esp.cpp:85-100
};

// primitive pool record of a class (la, lb), ne = (la+1)(lb+1)(la+lb+1):
//   [0] p  [1..3] P  [4] unused  [5 .. 5+ne) prefactor * E^x  [5+ne ..) E^y  [5+2ne ..) E^z      (ecoeffs.h layout)
constexpr int COSX_PRIM_HDR = 5;
// doubles per primitive record of class (la >= lb).  The s-s and p-s classes carry PRE-COMBINED coefficients instead of the
// three 1D tables (the products E^x E^y E^z are the same for all 128 points of a batch):
//   (0,0): [4] = prefactor E^x_000 E^y_000 E^z_000                                   -> V = [4] R_000
//   (1,0): [5..7] = cA_a, [8..10] = cB_a (a = x, y, z of the p shell)                 -> V_a = cA_a R_000 + cB_a R_(1 in a)
//          cA_a = prefactor E^a_{1,0;0} prod_{d != a} E^d_{0,0;0},  cB_a = the same with E^a_{1,0;1}
#ifndef COSX_DENSE_MAXL
#define COSX_DENSE_MAXL 2
#endif
#ifndef COSX_DENSE_30
#define COSX_DENSE_30 0 // the dense record also for the (3,0) class (56 products against 48 table entries): measured equal
#endif
__host__ __device__ constexpr bool cosx_dense(int la, int lb) {
  return (la + lb >= 2 && la + lb <= COSX_DENSE_MAXL) || (COSX_DENSE_30 && la == 3 && lb == 0);
}
// number of (a, b, t, u, v) products of a class in the dense pre-combined form: sum over Cartesian components of
// (ax + bx + 1)(ay + by + 1)(az + bz + 1); component order lx descending, then ly (kernels.h:256-290)
__host__ __device__ constexpr int cosx_dense_count(int la, int lb) {
  int n = 0;
  for (int ax = la; ax >= 0; --ax)
    for (int ay = la - ax; ay >= 0; --ay)
      for (int bx = lb; bx >= 0; --bx)
        for (int by = lb - bx; by >= 0; --by)
          n += (ax + bx + 1) * (ay + by + 1) * ((la - ax - ay) + (lb - bx - by) + 1);
  return n;
}
//   2 <= L <= COSX_DENSE_MAXL, i.e. (1,1), (2,0) (and (2,1), (3,0)): [5 ..) = prefactor E^x_t E^y_u E^z_v for every (a, b) and every (t, u, v) it reaches, in the
//          loop order of cosx_dense_count / prim_eval                                  -> V_ab = sum c_k R_tuv
__host__ __device__ constexpr int cosx_stride(int la, int lb) {
  return (la == 0 && lb == 0) ? COSX_PRIM_HDR
         : (la == 1 && lb == 0) ? COSX_PRIM_HDR + 6
         : cosx_dense(la, lb) ? COSX_PRIM_HDR + cosx_dense_count(la, lb)
                                : COSX_PRIM_HDR + 3 * (la + 1) * (lb + 1) * (la + lb + 1);
}
constexpr int COSX_BOYS_ORDERS = 16;
constexpr int COSX_BOYS_EXP_OFFSET = COSX_BOYS_ORDERS * 253 * 8; // behind the Boys rows: exp(-(segment + 1/2) / 7), 253 values
constexpr int COSX_BOYS_SEGMENTS = 253; // rows per order of the device Boys table (boys.h: 252 segments on [0, 36) + 1)

struct CosxParams {
  const CosxPair *pairs;
  const double *pool;
  const unsigned char *pair_active; // density screen of the current D2q (1 = evaluate)
  const int *tile_skip;             // 1 = batch failed max(phi) < precision
  const int *counts;                // valid points per tile
  const double *tcx, *tcy, *tcz, *trad; // batch bounding spheres
  const double *px, *py, *pz, *pw;  // padded SoA points / weights of all tiles
  const double *F;                  // chunk slabs
  double *G;
  const double *boys;               // Chebyshev table of boys.h, order-major: [m][segment][8]
  const double *c2s;                // spherical tables: l = 1 at 0 (3x3), l = 2 at 9 (5x6), l = 3 at 39 (7x10), l = 4 at 109 (9x15)
  int tile0;                        // first tile of the chunk (slabs are chunk-relative)
  int nbp;
  int pure;
  double margin;
  double w_full, w_lr, omega;
  int debug;                        // timing experiments (env OCC_COSX_DEBUG): 1 = no G atomics, 2 = no primitive loop
  unsigned long long *stats;        // [0] pair evaluations, [1] sum npt*nprim*nherm*nab of them; may be NULL
};

void launch_cosx_esp(const CosxParams &P, int la, int lb, int pair0, int npairs, int ntiles, void *stream);
void launch_cosx_tile_flags(const double *phi, const int *counts, int ntiles, int nbf, int nbp, double precision,
                            int *tile_skip, void *stream);
void launch_cosx_pair_flags(const CosxPair *pairs, const int *shell_size, int npairs, const double *D2q, int nbf,
                            double threshold, unsigned char *active, void *stream);
void launch_cosx_symmetrize(const double *Kraw, const double *KrawT, int nbf, double factor, double *K, int ldK,
                            void *stream); // K = factor (Kraw + KrawT^T)
void launch_cosx_scale_weights(double *phi_w, const double *phi, const double *pw, int ntiles, int nbf, int nbp,
                               void *stream);
// K' += Phi_t^T . G_t for the ntiles batches of a chunk (k_cosx_kgemm.cu); stats[0] += non-zero (batch, column block)s
void launch_cosx_kgemm(const double *phi, const double *G, const int *tile_skip, int ntiles, int nbf, int nbp, double *K,
                       unsigned long long *stats, void *stream);
void upload_default_c2s();              // xc_host.cu: the collocation kernel's constant table
const double *c2s_racah_table(int l);   // xc_host.cu: (2l+1) x ncart(l), Racah-normalised, CCA order

} // namespace xcb
