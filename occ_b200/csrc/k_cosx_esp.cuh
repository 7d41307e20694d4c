// k_cosx_esp.cuh — the ESP + contraction kernel of the COSX build, one instantiation per shell-pair class (la, lb);
// the instantiations are spread over k_cosx_la{0..3}.cu so that they compile in parallel. See k_cosx.cu for the design.
#pragma once
#include "cosx_internal.h"
#include <cuda_runtime.h>

namespace xcb {

__host__ __device__ constexpr int cx_ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int cx_nhermsum(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }
__host__ __device__ constexpr int cx_hidx(int t, int u, int v) {
  // rints.h:24-29
  return cx_nhermsum(t + u + v - 1) + (u + v) * (u + v + 1) / 2 + v;
}

// fully normalised cartesian -> spherical tables of gto.cpp:289-400 in global memory: l = 1 (the (y, z, x)
// permutation) at 0, l = 2 at 9, l = 3 at 39
template <int L> __device__ __forceinline__ const double *c2s_of(const double *tab) {
  return tab + (L == 1 ? 0 : (L == 2 ? 9 : 39));
}

// F~[c] = w * sum_m T[m][c] F[m]  (spherical) or w * F[c] (Cartesian): the transposed c2s applied to the row of F
template <int L> __device__ __forceinline__ void load_F(double (&Ft)[cx_ncart(L)], const double *col0, int bf, double w,
                                                        bool pure, const double *__restrict__ c2s) {
  constexpr int NC = cx_ncart(L), NS = 2 * L + 1;
  const double *src = col0 + static_cast<size_t>(bf) * TILE_P;
  if (!pure || L == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) Ft[c] = w * src[static_cast<size_t>(c) * TILE_P];
  } else {
    const double *T = c2s_of<L>(c2s);
#pragma unroll
    for (int c = 0; c < NC; ++c) Ft[c] = 0.0;
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      const double f = w * src[static_cast<size_t>(m) * TILE_P];
#pragma unroll
      for (int c = 0; c < NC; ++c) Ft[c] += __ldg(T + m * NC + c) * f;
    }
  }
}

template <int L> __device__ __forceinline__ void flush_G(const double (&Gt)[cx_ncart(L)], double *col0, int bf, bool pure,
                                                         const double *__restrict__ c2s) {
  constexpr int NC = cx_ncart(L), NS = 2 * L + 1;
  double *dst = col0 + static_cast<size_t>(bf) * TILE_P;
  if (!pure || L == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) atomicAdd(dst + static_cast<size_t>(c) * TILE_P, Gt[c]);
  } else {
    const double *T = c2s_of<L>(c2s);
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      double g = 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) g += __ldg(T + m * NC + c) * Gt[c];
      atomicAdd(dst + static_cast<size_t>(m) * TILE_P, g);
    }
  }
}

// seeds S[m] += scale * (-2p)^m F_m(p r^2)   (rints.h:176-181).  Boys: boys.h:17-64 asymptotic form for T >= 36; below,
// the order-7 table polynomial for m = L and F_m = (2T F_{m+1} + e^-T) / (2m+1) downwards (the reference interpolates
// every order; the two agree to the table's own accuracy, 1e-14).
template <int L>
__device__ __forceinline__ void add_seeds(const double *__restrict__ tab, double p, double r2, double scale,
                                          double (&S)[L + 1]) {
  const double T = p * r2;
  double F[L + 1];
  if (T >= 36.0) {
    const double inv = 1.0 / T;
    double Fm = 0.88622692545275801365 * sqrt(inv);
    F[0] = Fm;
#pragma unroll
    for (int i = 1; i <= L; ++i) {
      Fm = Fm * (static_cast<double>(i) - 0.5) * inv;
      F[i] = Fm;
    }
  } else {
    const double xod = T * 7.0; // 1 / Delta, Delta = 36 / 252
    const int seg = static_cast<int>(xod);
    const double xd = xod - static_cast<double>(seg) - 0.5;
    const double2 *c = reinterpret_cast<const double2 *>(tab + (static_cast<size_t>(seg) * 16 + L) * 8);
    const double2 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
    double r = c3.y;
    r = r * xd + c3.x;
    r = r * xd + c2.y;
    r = r * xd + c2.x;
    r = r * xd + c1.y;
    r = r * xd + c1.x;
    r = r * xd + c0.y;
    r = r * xd + c0.x;
    F[L] = r;
    if (L > 0) {
      const double e = exp(-T), T2 = 2.0 * T;
#pragma unroll
      for (int m = L - 1; m >= 0; --m) F[m] = (T2 * F[m + 1] + e) * (1.0 / (2 * m + 1));
    }
  }
  double pw = scale;
  const double m2p = -2.0 * p;
#pragma unroll
  for (int m = 0; m <= L; ++m) {
    S[m] += pw * F[m];
    pw *= m2p;
  }
}

// Hermite Coulomb integrals R_tuv, t+u+v <= L, from the seeds (rints.h:183-219). Same recurrences as the reference
// (x for u = v = 0, y for v = 0, z otherwise), evaluated level by level in the auxiliary index m so that only two
// levels are live at a time.
template <int L>
__device__ __forceinline__ void hermite_R(const double (&S)[L + 1], double x, double y, double z,
                                          double (&R)[cx_nhermsum(L)]) {
  // every loop has a template-constant trip count (guards select the live iterations) so that the unroller resolves
  // all indices at compile time and the two level arrays stay in registers
  constexpr int NH = cx_nhermsum(L);
  double A[NH];
  A[0] = S[L];
#pragma unroll
  for (int m = L - 1; m >= 0; --m) {
    double B[NH];
    B[0] = S[m];
#pragma unroll
    for (int N = 1; N <= L; ++N) {
      if (N <= L - m) {
#pragma unroll
        for (int uv = 0; uv <= L; ++uv) {
          if (uv <= N) {
#pragma unroll
            for (int v = 0; v <= L; ++v) {
              if (v <= uv) {
                const int u = uv - v, t = N - uv;
                double val;
                if (v > 0) {
                  val = z * A[cx_hidx(t, u, v - 1)];
                  if (v > 1) val += (v - 1) * A[cx_hidx(t, u, v - 2)];
                } else if (u > 0) {
                  val = y * A[cx_hidx(t, u - 1, 0)];
                  if (u > 1) val += (u - 1) * A[cx_hidx(t, u - 2, 0)];
                } else {
                  val = x * A[cx_hidx(t - 1, 0, 0)];
                  if (t > 1) val += (t - 1) * A[cx_hidx(t - 2, 0, 0)];
                }
                B[cx_hidx(t, u, v)] = val;
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NH; ++i)
      if (i < cx_nhermsum(L - m)) A[i] = B[i];
  }
#pragma unroll
  for (int i = 0; i < NH; ++i) R[i] = A[i];
}

// V_ab = sum_t Ex[ax,bx,t] sum_u Ey[ay,by,u] sum_v Ez[az,bz,v] R_tuv, folded straight into the two G rows
// (kernels.h:256-290 builds the same products into the dense E-matrix; Cartesian order lx descending, then ly).
template <int LA, int LB>
__device__ __forceinline__ void contract(const double *__restrict__ Ex, const double *__restrict__ Ey,
                                         const double *__restrict__ Ez, const double (&R)[cx_nhermsum(LA + LB)],
                                         const double (&Fp)[cx_ncart(LA)], const double (&Fq)[cx_ncart(LB)],
                                         double (&Gp)[cx_ncart(LA)], double (&Gq)[cx_ncart(LB)], bool offdiag) {
  // template-constant trip counts + guards: see hermite_R
  constexpr int L = LA + LB, L1 = L + 1;
#pragma unroll
  for (int ax = LA; ax >= 0; --ax) {
#pragma unroll
    for (int ay = LA; ay >= 0; --ay) {
      if (ay <= LA - ax) {
        const int az = LA - ax - ay;
        const int ia = (LA - ax) * (LA - ax + 1) / 2 + (LA - ax - ay); // CCA Cartesian index
        double gp = 0.0;
#pragma unroll
        for (int bx = LB; bx >= 0; --bx) {
#pragma unroll
          for (int by = LB; by >= 0; --by) {
            if (by <= LB - bx) {
              const int bz = LB - bx - by;
              const int jb = (LB - bx) * (LB - bx + 1) / 2 + (LB - bx - by);
              const double *ex = Ex + (ax * (LB + 1) + bx) * L1;
              const double *ey = Ey + (ay * (LB + 1) + by) * L1;
              const double *ez = Ez + (az * (LB + 1) + bz) * L1;
              double V = 0.0;
#pragma unroll
              for (int t = 0; t <= L; ++t) {
                if (t <= ax + bx) {
                  double sy = 0.0;
#pragma unroll
                  for (int u = 0; u <= L; ++u) {
                    if (u <= ay + by) {
                      double sz = 0.0;
#pragma unroll
                      for (int v = 0; v <= L; ++v)
                        if (v <= az + bz) sz += ez[v] * R[cx_hidx(t, u, v)];
                      sy += ey[u] * sz;
                    }
                  }
                  V += ex[t] * sy;
                }
              }
              Gq[jb] += Fp[ia] * V;
              gp += Fq[jb] * V;
            }
          }
        }
        if (offdiag) Gp[ia] += gp;
      }
    }
  }
}

template <int LA, int LB>
__global__ void __launch_bounds__(TILE_P) k_cosx_esp(CosxParams P, int pair0, int npairs, int chunk, int cls) {
  constexpr int L = LA + LB, NE = (LA + 1) * (LB + 1) * (L + 1), STRIDE = COSX_PRIM_HDR + 3 * NE;
  constexpr int NCA = cx_ncart(LA), NCB = cx_ncart(LB), NH = cx_nhermsum(L);
  constexpr int STAGE = (1024 / STRIDE) > 0 ? (1024 / STRIDE) : 1;
  __shared__ double sm[STAGE * STRIDE];
  const int tile = P.tile0 + blockIdx.x;
  if (P.tile_skip[tile]) return;
  const int tid = threadIdx.x;
  const int k0 = pair0 + blockIdx.y * chunk;
  const int k1 = min(k0 + chunk, pair0 + npairs);
  const double tcx = P.tcx[tile], tcy = P.tcy[tile], tcz = P.tcz[tile], lim = P.trad[tile] + P.margin;
  const size_t pt = static_cast<size_t>(tile) * TILE_P + tid;
  const double X = P.px[pt], Y = P.py[pt], Z = P.pz[pt], w = P.pw[pt];
  const size_t slab = static_cast<size_t>(blockIdx.x) * P.nbp * TILE_P + tid;
  const double *Fcol = P.F + slab;
  double *Gcol = P.G + slab;
  const bool pure = P.pure != 0;
  const double w2 = P.omega * P.omega;

  double Fp[NCA], Gp[NCA];
  int cur_p = -1, cur_bfp = 0;
  bool gp_dirty = false;
  for (int k = k0; k < k1; ++k) {
    if (!P.pair_active[k]) continue; // seminumerical_exchange.cpp:758-764
    const CosxPair pr = P.pairs[k];
    const double dx = pr.cx - tcx, dy = pr.cy - tcy, dz = pr.cz - tcz;
    if (sqrt(dx * dx + dy * dy + dz * dz) - pr.extent >= lim) continue; // :744-748
    if (pr.p != cur_p) {
      if (gp_dirty) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
      cur_p = pr.p;
      cur_bfp = pr.bfp;
      gp_dirty = false;
      load_F<LA>(Fp, Fcol, pr.bfp, w, pure, P.c2s);
#pragma unroll
      for (int i = 0; i < NCA; ++i) Gp[i] = 0.0;
    }
    double Fq[NCB], Gq[NCB];
    load_F<LB>(Fq, Fcol, pr.bfq, w, pure, P.c2s);
#pragma unroll
    for (int j = 0; j < NCB; ++j) Gq[j] = 0.0;
    const bool offdiag = pr.p != pr.q;
    gp_dirty = gp_dirty || offdiag;
    for (int i0 = 0; i0 < pr.nprim; i0 += STAGE) {
      const int ns = min(STAGE, pr.nprim - i0);
      __syncthreads();
      const double *src = P.pool + pr.off + static_cast<size_t>(i0) * STRIDE;
      for (int i = tid; i < ns * STRIDE; i += TILE_P) sm[i] = __ldg(src + i);
      __syncthreads();
      for (int s = 0; s < ns; ++s) {
        const double *rec = sm + s * STRIDE;
        const double pe = rec[0];
        const double PCx = rec[1] - X, PCy = rec[2] - Y, PCz = rec[3] - Z; // rints_specialized.h:1457-1460: P - C
        const double r2 = PCx * PCx + PCy * PCy + PCz * PCz;
        double S[L + 1];
#pragma unroll
        for (int m = 0; m <= L; ++m) S[m] = 0.0;
        if (P.w_full != 0.0) add_seeds<L>(P.boys, pe, r2, P.w_full, S);
        if (P.w_lr != 0.0) { // rints_specialized.h:1462-1500: p_eff = p w^2 / (p + w^2), scale sqrt(w^2 / (p + w^2))
          const double sc = w2 / (pe + w2);
          add_seeds<L>(P.boys, pe * sc, r2, P.w_lr * sqrt(sc), S);
        }
        double R[NH];
        hermite_R<L>(S, PCx, PCy, PCz, R);
        contract<LA, LB>(rec + COSX_PRIM_HDR, rec + COSX_PRIM_HDR + NE, rec + COSX_PRIM_HDR + 2 * NE, R, Fp, Fq, Gp, Gq,
                         offdiag);
      }
    }
    flush_G<LB>(Gq, Gcol, pr.bfq, pure, P.c2s);
    if (P.stats && tid == 0) {
      atomicAdd(P.stats + 2 * cls, 1ULL);
      atomicAdd(P.stats + 2 * cls + 1, static_cast<unsigned long long>(pr.nprim) * P.counts[tile]);
    }
  }
  if (gp_dirty) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
}

template <int LA, int LB>
void launch_class(const CosxParams &P, int pair0, int npairs, int ntiles, cudaStream_t st) {
  constexpr int L = LA + LB;
  int chunk = L == 0 ? 64 : (L == 1 ? 32 : (L == 2 ? 16 : (L == 3 ? 8 : 4)));
  while ((npairs + chunk - 1) / chunk > 65535) chunk *= 2;
  dim3 grid(ntiles, (npairs + chunk - 1) / chunk);
  k_cosx_esp<LA, LB><<<grid, TILE_P, 0, st>>>(P, pair0, npairs, chunk, LA * (COSX_MAX_L + 1) + LB);
}

} // namespace xcb
