// k_cosx_esp.cuh — the ESP + contraction kernel of the COSX build, one instantiation per shell-pair class (la, lb);
// the instantiations are spread over k_cosx_la{0..3}.cu so that they compile in parallel. See k_cosx.cu for the design.
#pragma once
#include "cosx_internal.h"
#include <cuda_runtime.h>

#ifndef COSX_EXP_TABLE
#define COSX_EXP_TABLE 1
#endif

namespace xcb {

__host__ __device__ constexpr int cx_ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int cx_nhermsum(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }
__host__ __device__ constexpr int cx_hidx(int t, int u, int v) {
  // rints.h:24-29
  return cx_nhermsum(t + u + v - 1) + (u + v) * (u + v + 1) / 2 + v;
}

// fully normalised cartesian -> spherical tables of gto.cpp:289-400 in global memory: l = 1 (the (y, z, x)
// permutation) at 0, l = 2 at 9, l = 3 at 39, l = 4 at 109
template <int L> __device__ __forceinline__ const double *c2s_of(const double *tab) {
  return tab + (L == 1 ? 0 : (L == 2 ? 9 : (L == 3 ? 39 : 109)));
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
    // F_m = (2m-1)!! / 2^(m+1) sqrt(pi / T^(2m+1)) (boys.h:17-64).  One reciprocal square root (MUFU.RSQ64H + Newton,
    // 2 ulp) instead of the division and the square root of the reference's form: those two cost more than the
    // table path they stand in for, and most s-s / p-s evaluations of a large molecule take this branch
    const double rs = rsqrt(T), inv = rs * rs;
    double Fm = 0.88622692545275801365 * rs;
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
    // Estrin evaluation of the degree-7 polynomial (dependency depth 4 instead of 8; boys.h:101-106 uses Horner)
    const double x2 = xd * xd, x4 = x2 * x2;
    auto row = [&](int m) {
      const double2 *c = reinterpret_cast<const double2 *>(tab + (static_cast<size_t>(m) * COSX_BOYS_SEGMENTS + seg) * 8);
      const double2 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
      const double q0 = fma(c0.y, xd, c0.x), q1 = fma(c1.y, xd, c1.x), q2 = fma(c2.y, xd, c2.x), q3 = fma(c3.y, xd, c3.x);
      return fma(fma(q3, x2, q2), x4, fma(q1, x2, q0));
    };
    F[L] = row(L);
    if (L > 0) { // (a second table row for L = 1 instead of exp + recursion was measured: +48 ms on valinomycin, L1-bound)
#if COSX_EXP_TABLE
      // e^-T = e^-(seg + 1/2)/7 . e^(-xd/7): the segment's midpoint value from a 2 KB table behind the Boys rows, the
      // rest (|xd / 7| <= 1/14) by a degree-9 Taylor polynomial (remainder 1e-18) — about half the instructions of exp()
      const double y = xd * (-1.0 / 7.0);
      double ey = fma(y, 1.0 / 362880.0, 1.0 / 40320.0);
      ey = fma(ey, y, 1.0 / 5040.0);
      ey = fma(ey, y, 1.0 / 720.0);
      ey = fma(ey, y, 1.0 / 120.0);
      ey = fma(ey, y, 1.0 / 24.0);
      ey = fma(ey, y, 1.0 / 6.0);
      ey = fma(ey, y, 0.5);
      ey = fma(ey, y, 1.0);
      ey = fma(ey, y, 1.0);
      const double e = __ldg(tab + COSX_BOYS_EXP_OFFSET + seg) * ey, T2 = 2.0 * T;
#else
      const double e = exp(-T), T2 = 2.0 * T;
#endif
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

// Sliced form of hermite_R + contract for L = la + lb >= 4, where the 35-84 R_tuv no longer fit in registers next to F and
// G. The Hermite integrals are generated slice by slice — rows R^m_{t00} in t (two rows live), rows R^m_{tu0} in u (two rows
// live), then the 1D recursion in v for one (t, u) — and each (t, u) slice is contracted at once:
//   S[az][bz] = sum_v Ez[az,bz,v] R_tuv,   G += F . (Ex[ax,bx,t] Ey[ay,by,u]) S[az][bz]
// Same recurrences and the same products as the reference (rints.h:183-219, kernels.h:256-290), other summation order.
template <int LA, int LB>
__device__ __forceinline__ void prim_sliced(const double *__restrict__ Ex, const double *__restrict__ Ey,
                                            const double *__restrict__ Ez, const double (&S)[LA + LB + 1], double x,
                                            double y, double z, const double (&Fp)[cx_ncart(LA)],
                                            const double (&Fq)[cx_ncart(LB)], double (&Gp)[cx_ncart(LA)],
                                            double (&Gq)[cx_ncart(LB)], bool offdiag) {
  constexpr int L = LA + LB, L1 = L + 1;
  double Xc[L1], Xp[L1]; // R^m_{t00} and R^m_{t-1,00}
#pragma unroll
  for (int m = 0; m <= L; ++m) {
    Xc[m] = S[m];
    Xp[m] = 0.0;
  }
  double gp[cx_ncart(LA)];
#pragma unroll
  for (int i = 0; i < cx_ncart(LA); ++i) gp[i] = 0.0;
#pragma unroll
  for (int t = 0; t <= L; ++t) {
    // rows in u for this t: Yc[m] = R^m_{t,u,0}, Yp[m] = R^m_{t,u-1,0}
    double Yc[L1], Yp[L1];
#pragma unroll
    for (int m = 0; m <= L; ++m) {
      Yc[m] = (m <= L - t) ? Xc[m] : 0.0;
      Yp[m] = 0.0;
    }
#pragma unroll
    for (int u = 0; u <= L; ++u) {
      if (u <= L - t) {
        const int k = L - t - u; // v runs 0..k
        // 1D recursion in v: c[m] = R^m_{t,u,v}
        double Rv[L1], c[L1], cp[L1];
#pragma unroll
        for (int m = 0; m <= L; ++m) {
          c[m] = (m <= k) ? Yc[m] : 0.0;
          cp[m] = 0.0;
        }
        Rv[0] = c[0];
#pragma unroll
        for (int v = 0; v < L; ++v) {
          if (v < k) {
            double cn[L1];
#pragma unroll
            for (int m = 0; m <= L; ++m) {
              cn[m] = 0.0;
              if (m <= k - v - 1) {
                cn[m] = z * c[m + 1];
                if (v > 0) cn[m] += v * cp[m + 1];
              }
            }
#pragma unroll
            for (int m = 0; m <= L; ++m) {
              cp[m] = c[m];
              c[m] = cn[m];
            }
            Rv[v + 1] = c[0];
          }
        }
        // S[az][bz]
        double Sz[LA + 1][LB + 1];
#pragma unroll
        for (int az = 0; az <= LA; ++az) {
#pragma unroll
          for (int bz = 0; bz <= LB; ++bz) {
            double s = 0.0;
            if (t + u <= L - az - bz) {
              const double *ez = Ez + (az * (LB + 1) + bz) * L1;
#pragma unroll
              for (int v = 0; v <= L; ++v)
                if (v <= az + bz && v <= k) s += ez[v] * Rv[v];
            }
            Sz[az][bz] = s;
          }
        }
#pragma unroll
        for (int ax = 0; ax <= LA; ++ax) {
#pragma unroll
          for (int bx = 0; bx <= LB; ++bx) {
            if (ax + bx >= t) {
              const double ext = Ex[(ax * (LB + 1) + bx) * L1 + t];
#pragma unroll
              for (int ay = 0; ay <= LA; ++ay) {
#pragma unroll
                for (int by = 0; by <= LB; ++by) {
                  if (ay <= LA - ax && by <= LB - bx && ay + by >= u) {
                    const int az = LA - ax - ay, bz = LB - bx - by;
                    const int ia = (LA - ax) * (LA - ax + 1) / 2 + (LA - ax - ay);
                    const int jb = (LB - bx) * (LB - bx + 1) / 2 + (LB - bx - by);
                    const double V = (ext * Ey[(ay * (LB + 1) + by) * L1 + u]) * Sz[az][bz];
                    Gq[jb] += Fp[ia] * V;
                    gp[ia] += Fq[jb] * V;
                  }
                }
              }
            }
          }
        }
        // next row in u: R^m_{t,u+1,0} = y R^{m+1}_{t,u,0} + u R^{m+1}_{t,u-1,0}
        if (u < L - t) {
          double Yn[L1];
#pragma unroll
          for (int m = 0; m <= L; ++m) {
            Yn[m] = 0.0;
            if (m <= L - t - u - 1) {
              Yn[m] = y * Yc[m + 1];
              if (u > 0) Yn[m] += u * Yp[m + 1];
            }
          }
#pragma unroll
          for (int m = 0; m <= L; ++m) {
            Yp[m] = Yc[m];
            Yc[m] = Yn[m];
          }
        }
      }
    }
    // next row in t: R^m_{t+1,00} = x R^{m+1}_{t00} + t R^{m+1}_{t-1,00}
    if (t < L) {
      double Xn[L1];
#pragma unroll
      for (int m = 0; m <= L; ++m) {
        Xn[m] = 0.0;
        if (m <= L - t - 1) {
          Xn[m] = x * Xc[m + 1];
          if (t > 0) Xn[m] += t * Xp[m + 1];
        }
      }
#pragma unroll
      for (int m = 0; m <= L; ++m) {
        Xp[m] = Xc[m];
        Xc[m] = Xn[m];
      }
    }
  }
  if (offdiag) {
#pragma unroll
    for (int i = 0; i < cx_ncart(LA); ++i) Gp[i] += gp[i];
  }
}

// one primitive pair of class (LA, LB): Hermite integrals from the seeds S, contraction with the E coefficients of the
// record, F . V into the two G rows.  s-s and p-s use the pre-combined coefficients of the record (cosx_stride).
template <int LA, int LB>
__device__ __forceinline__ void prim_eval(const double *__restrict__ rec, const double (&S)[LA + LB + 1], double x, double y,
                                          double z, const double (&Fp)[cx_ncart(LA)], const double (&Fq)[cx_ncart(LB)],
                                          double (&Gp)[cx_ncart(LA)], double (&Gq)[cx_ncart(LB)], bool offdiag) {
  constexpr int L = LA + LB, NE = (LA + 1) * (LB + 1) * (L + 1);
  if constexpr (L == 0) {
    const double V = rec[4] * S[0];
    Gq[0] += Fp[0] * V;
    if (offdiag) Gp[0] += Fq[0] * V;
  } else if constexpr (LA == 1 && LB == 0) {
    const double r[3] = {x * S[1], y * S[1], z * S[1]}; // R_100, R_010, R_001 (rints.h:183-219 with L = 1)
    double gq = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double V = fma(rec[COSX_PRIM_HDR + 3 + a], r[a], rec[COSX_PRIM_HDR + a] * S[0]);
      gq = fma(Fp[a], V, gq);
      if (offdiag) Gp[a] = fma(Fq[0], V, Gp[a]);
    }
    Gq[0] += gq;
  } else if constexpr (cosx_dense(LA, LB)) { // dense pre-combined coefficients, every index a compile-time constant
    double R[cx_nhermsum(L)];
    hermite_R<L>(S, x, y, z, R);
    const double *c = rec + COSX_PRIM_HDR;
    int k = 0;
#pragma unroll
    for (int ax = LA; ax >= 0; --ax) {
#pragma unroll
      for (int ay = LA; ay >= 0; --ay) {
        if (ay <= LA - ax) {
          const int az = LA - ax - ay;
          const int ia = (LA - ax) * (LA - ax + 1) / 2 + (LA - ax - ay);
          double gp = 0.0;
#pragma unroll
          for (int bx = LB; bx >= 0; --bx) {
#pragma unroll
            for (int by = LB; by >= 0; --by) {
              if (by <= LB - bx) {
                const int bz = LB - bx - by;
                const int jb = (LB - bx) * (LB - bx + 1) / 2 + (LB - bx - by);
                double V = 0.0;
#pragma unroll
                for (int t = 0; t <= L; ++t)
#pragma unroll
                  for (int u = 0; u <= L; ++u)
#pragma unroll
                    for (int v = 0; v <= L; ++v)
                      if (t <= ax + bx && u <= ay + by && v <= az + bz) V = fma(c[k++], R[cx_hidx(t, u, v)], V);
                Gq[jb] = fma(Fp[ia], V, Gq[jb]);
                gp = fma(Fq[jb], V, gp);
              }
            }
          }
          if (offdiag) Gp[ia] += gp;
        }
      }
    }
  } else if constexpr (L >= 4) {
    prim_sliced<LA, LB>(rec + COSX_PRIM_HDR, rec + COSX_PRIM_HDR + NE, rec + COSX_PRIM_HDR + 2 * NE, S, x, y, z, Fp, Fq, Gp, Gq,
                        offdiag);
  } else {
    double R[cx_nhermsum(L)];
    hermite_R<L>(S, x, y, z, R);
    contract<LA, LB>(rec + COSX_PRIM_HDR, rec + COSX_PRIM_HDR + NE, rec + COSX_PRIM_HDR + 2 * NE, R, Fp, Fq, Gp, Gq, offdiag);
  }
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// read the staged F columns of a shell: F~[c] = w sum_m T[m][c] F[m] (spherical) or w F[c]
template <int L>
__device__ __forceinline__ void take_F(double (&Ft)[cx_ncart(L)], const double *sF /* [col][128] + tid */, double w,
                                       bool pure, const double *__restrict__ c2s) {
  constexpr int NC = cx_ncart(L), NS = 2 * L + 1;
  if (!pure || L == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) Ft[c] = w * sF[c * TILE_P];
  } else {
    const double *T = c2s_of<L>(c2s);
#pragma unroll
    for (int c = 0; c < NC; ++c) Ft[c] = 0.0;
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      const double f = w * sF[m * TILE_P];
#pragma unroll
      for (int c = 0; c < NC; ++c) Ft[c] += __ldg(T + m * NC + c) * f;
    }
  }
}

// minimum resident CTAs per SM asked of ptxas per class (a register cap; 0 = none). The low-l classes are latency-bound at
// 64-250 registers; capping them (with at most 136 B of spill) bought about 90 ms on valinomycin: (0,0) 64 -> 56 registers
// (161 -> 149 ms), (1,0) 124 -> 72 (275 -> 235 ms), (1,1) 120 -> 80, (2,0) 128 -> 96 (80 was slower again), (2,1) 162 -> 128,
// (3,0) 166 -> 128, (3,1) 248 -> 168.
#ifndef COSX_M00
#define COSX_M00 9
#endif
#ifndef COSX_M10
#define COSX_M10 7
#endif
#ifndef COSX_M11
#define COSX_M11 6
#endif
#ifndef COSX_M20
#define COSX_M20 5
#endif
#ifndef COSX_M21
#define COSX_M21 4
#endif
#ifndef COSX_M22
#define COSX_M22 0
#endif
#ifndef COSX_M30
#define COSX_M30 4
#endif
#ifndef COSX_M31
#define COSX_M31 3
#endif
#ifndef COSX_M32
#define COSX_M32 0
#endif
#ifndef COSX_M33
#define COSX_M33 0
#endif
__host__ __device__ constexpr int cosx_min_ctas(int la, int lb) {
  constexpr int t[5][5] = {{COSX_M00, 0, 0, 0, 0},
                           {COSX_M10, COSX_M11, 0, 0, 0},
                           {COSX_M20, COSX_M21, COSX_M22, 0, 0},
                           {COSX_M30, COSX_M31, COSX_M32, COSX_M33, 0},
                           {0, 0, 0, 0, 0}}; // g shells: no cap
  return t[la][lb];
}

// which classes run the warp-private kernel (k_cosx_esp_w): those with la + lb <= COSX_WARP_MAXL.  Measured on valinomycin
// (whole compute_K_mo call, profiles/r02_experiments.md): s-s 1406 -> 1394 ms; every class with L >= 1 is SLOWER there
// (L = 1 +22 ms, L = 2 +52 ms, all classes 1534 ms) because its primitive records then come from global memory in the
// inner loop, so only s-s uses it.  OCC_COSX_DEBUG bit 8 forces the staged kernel everywhere (A/B timing).
#ifndef COSX_WARP_MAXL
#define COSX_WARP_MAXL 0
#endif

struct PairLite {
  int p, q, bfp, bfq, nprim, off;
};

// One CTA = one grid batch x one run of `chunk` shell pairs of class (LA, LB).
//   round:    128 candidate pairs are screened in parallel (density flag, geometric test) and the survivors compacted,
//             in order, into shared memory — one global-load latency per 128 pairs instead of one per pair;
//   pipeline: the primitive records and the F columns of the surviving pairs stream through an NBUF-deep cp.async ring,
//             so the loads of the next pairs are in flight while the current one is evaluated.
template <int LA, int LB>
__global__ void __launch_bounds__(TILE_P, cosx_min_ctas(LA, LB)) k_cosx_esp(CosxParams P, int pair0, int npairs, int chunk, int cls, int ntiles,
                                                     int nruns) {
  constexpr int L = LA + LB, NE = (LA + 1) * (LB + 1) * (L + 1), STRIDE = cosx_stride(LA, LB);
  constexpr int NCA = cx_ncart(LA), NCB = cx_ncart(LB), NH = cx_nhermsum(L);
#ifndef COSX_GROUPED
#define COSX_GROUPED 1
#endif
  // classes whose q side is an s shell have so little work per pair (often one or two primitives) that the two CTA
  // barriers of a pipeline step weigh as much as the integrals: there a step carries up to U consecutive pairs
  constexpr bool GROUPED = COSX_GROUPED && LB == 0;
  constexpr int U = GROUPED ? 4 : 1;
  constexpr int NBUF = GROUPED ? (L == 1 ? 3 : 2) : (L <= 1 ? 4 : 2);
  // s-s runs 7 CTAs per SM: with 32 KB of shared memory each only ~30 KB of the unified L1 would be left for the Boys
  // slice and the F / G columns, so its pool is halved (a pair with more than 32 primitives continues in the next step)
  constexpr int POOLCAP = L == 0 ? 256 : (L == 1 ? 512 : 1024);
  constexpr int STAGE = (POOLCAP / STRIDE) > 0 ? (POOLCAP / STRIDE) : 1;
  __shared__ double s_pool[NBUF][STAGE * STRIDE];
  __shared__ double s_F[NBUF][U][NCB][TILE_P];
  __shared__ PairLite s_pairs[TILE_P]; // survivors of a screening round, without the screening-only fields
  __shared__ int s_wcnt[TILE_P / 32];
  // 1D grid, group-major: the CTAs of COSX_TILE_GROUP batches x all pair runs are adjacent, so that the F and G slabs of
  // a group (2 x 8 x 128 x nbp doubles) stay L2-resident while every pair run reads / accumulates into them
  const int tl = blockIdx.x % COSX_TILE_GROUP;
  const int yrun = (blockIdx.x / COSX_TILE_GROUP) % nruns;
  const int ltile = (blockIdx.x / (COSX_TILE_GROUP * nruns)) * COSX_TILE_GROUP + tl; // chunk-relative batch
  if (ltile >= ntiles) return;
  const int tile = P.tile0 + ltile;
  if (P.tile_skip[tile]) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k0 = pair0 + yrun * chunk;
  const int k1 = min(k0 + chunk, pair0 + npairs);
  const double tcx = P.tcx[tile], tcy = P.tcy[tile], tcz = P.tcz[tile], lim = P.trad[tile] + P.margin;
  const size_t pt = static_cast<size_t>(tile) * TILE_P + tid;
  const double X = P.px[pt], Y = P.py[pt], Z = P.pz[pt], w = P.pw[pt];
  const size_t slab = static_cast<size_t>(ltile) * P.nbp * TILE_P + tid;
  const double *Fcol = P.F + slab;
  double *Gcol = P.G + slab;
  const bool pure = P.pure != 0;
  const int nq_cols = (pure && LB > 0) ? 2 * LB + 1 : NCB;
  const double w2 = P.omega * P.omega;

  double Fp[NCA], Gp[NCA], Fq[NCB], Gq[NCB];
  int cur_p = -1, cur_bfp = 0;
  bool gp_dirty = false;
  for (int kr = k0; kr < k1; kr += TILE_P) {
    // ---- screen 128 candidates (seminumerical_exchange.cpp:744-748 geometric, :758-764 density) ----
    const int k = kr + tid;
    bool keep = false;
    CosxPair cand;
    if (k < k1 && P.pair_active[k]) {
      cand = P.pairs[k];
      const double dx = cand.cx - tcx, dy = cand.cy - tcy, dz = cand.cz - tcz;
      keep = !(sqrt(dx * dx + dy * dy + dz * dz) - cand.extent >= lim);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncthreads(); // previous round fully consumed (s_pairs, s_wcnt reuse)
    if (lane == 0) s_wcnt[wid] = __popc(bal);
    __syncthreads();
    int base = 0, nsurv = 0;
#pragma unroll
    for (int i = 0; i < TILE_P / 32; ++i) {
      if (i < wid) base += s_wcnt[i];
      nsurv += s_wcnt[i];
    }
    if (keep) s_pairs[base + __popc(bal & ((1u << lane) - 1u))] = PairLite{cand.p, cand.q, cand.bfp, cand.bfq, cand.nprim, cand.off};
    __syncthreads();
    if (nsurv == 0) continue;

    if constexpr (GROUPED) {
      // ---- cp.async ring over steps of up to U consecutive (pair, primitive-range) segments ----
      // producer and consumer walk the compacted pair list with the same deterministic packing rule, so no segment
      // descriptors need to be stored: a step takes segments while fewer than U are in it and its pool has room; a pair
      // whose primitives do not fit continues in the next step (Gq stays in registers across steps)
      int li_n = 0, i0_n = 0, li = 0, i0 = 0, issued = 0, consumed = 0;
      auto issue_step = [&](int buf) {
        int used = 0;
#pragma unroll 1
        for (int j = 0; j < U && li_n < nsurv && used < STAGE; ++j) {
          const PairLite &pr = s_pairs[li_n];
          const int ns = min(pr.nprim - i0_n, STAGE - used);
          const double *src = P.pool + pr.off + static_cast<size_t>(i0_n) * STRIDE;
          double *dst = &s_pool[buf][used * STRIDE];
          for (int i = tid; i < ns * STRIDE; i += TILE_P) cp_async8(dst + i, src + i);
          if (i0_n == 0) {
            const double *f = Fcol + static_cast<size_t>(pr.bfq) * TILE_P;
            for (int c = 0; c < nq_cols; ++c) cp_async8(&s_F[buf][j][c][tid], f + static_cast<size_t>(c) * TILE_P);
          }
          used += ns;
          i0_n += ns;
          if (i0_n < pr.nprim) break; // pool full: the pair continues in the next step
          i0_n = 0;
          ++li_n;
        }
      };
#pragma unroll
      for (int j = 0; j < NBUF - 1; ++j) {
        if (li_n < nsurv) issue_step(issued % NBUF);
        cp_async_commit();
        ++issued;
      }
      while (li < nsurv) {
        if (li_n < nsurv) issue_step(issued % NBUF);
        cp_async_commit();
        ++issued;
        cp_async_wait<NBUF - 1>();
        __syncthreads();
        const int buf = consumed % NBUF;
        int used = 0;
#pragma unroll 1
        for (int j = 0; j < U && li < nsurv && used < STAGE; ++j) {
          const PairLite &pr = s_pairs[li];
          const bool offdiag = pr.p != pr.q;
          const int ns = min(pr.nprim - i0, STAGE - used);
          if (i0 == 0) {
            if (pr.p != cur_p) {
              if (gp_dirty) if (!(P.debug & 1)) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
              cur_p = pr.p;
              cur_bfp = pr.bfp;
              gp_dirty = false;
              load_F<LA>(Fp, Fcol, pr.bfp, w, pure, P.c2s);
#pragma unroll
              for (int i = 0; i < NCA; ++i) Gp[i] = 0.0;
            }
            take_F<LB>(Fq, &s_F[buf][j][0][tid], w, pure, P.c2s);
#pragma unroll
            for (int jj = 0; jj < NCB; ++jj) Gq[jj] = 0.0;
            gp_dirty = gp_dirty || offdiag;
            if (P.stats && tid == 0) {
              atomicAdd(P.stats + 2 * cls, 1ULL);
              atomicAdd(P.stats + 2 * cls + 1, static_cast<unsigned long long>(pr.nprim) * P.counts[tile]);
            }
          }
#pragma unroll(L == 0 ? 4 : (L <= 2 ? 2 : 1))
          for (int sidx = 0; sidx < ((P.debug & 2) ? 0 : ns); ++sidx) {
            const double *rec = s_pool[buf] + (used + sidx) * STRIDE;
            const double pe = rec[0];
            const double PCx = rec[1] - X, PCy = rec[2] - Y, PCz = rec[3] - Z; // rints_specialized.h:1457-1460: P - C
            const double r2 = PCx * PCx + PCy * PCy + PCz * PCz;
            double S[L + 1];
#pragma unroll
            for (int m = 0; m <= L; ++m) S[m] = 0.0;
            if (P.w_full != 0.0) add_seeds<L>(P.boys, pe, r2, P.w_full, S);
            if (P.w_lr != 0.0) { // rints_specialized.h:1462-1500
              const double sc = w2 / (pe + w2);
              add_seeds<L>(P.boys, pe * sc, r2, P.w_lr * sqrt(sc), S);
            }
            prim_eval<LA, LB>(rec, S, PCx, PCy, PCz, Fp, Fq, Gp, Gq, offdiag);
          }
          used += ns;
          i0 += ns;
          if (i0 < pr.nprim) break;
          if (!(P.debug & 1)) flush_G<LB>(Gq, Gcol, pr.bfq, pure, P.c2s);
          i0 = 0;
          ++li;
        }
        ++consumed;
        __syncthreads(); // the buffer just read is the target of the next issue
      }
      cp_async_wait<0>();
      continue;
    }
    // ---- cp.async ring over (pair, primitive-chunk) units ----
    auto issue = [&](int li, int i0, int buf) {
      const PairLite &pr = s_pairs[li];
      const int ns = min(STAGE, pr.nprim - i0);
      const double *src = P.pool + pr.off + static_cast<size_t>(i0) * STRIDE;
      for (int i = tid; i < ns * STRIDE; i += TILE_P) cp_async8(&s_pool[buf][i], src + i);
      if (i0 == 0) {
        const double *f = Fcol + static_cast<size_t>(pr.bfq) * TILE_P;
        for (int c = 0; c < nq_cols; ++c) cp_async8(&s_F[buf][0][c][tid], f + static_cast<size_t>(c) * TILE_P);
      }
    };
    int li_n = 0, i0_n = 0, issued = 0, consumed = 0;
    auto advance_n = [&]() {
      i0_n += STAGE;
      if (i0_n >= s_pairs[li_n].nprim) {
        i0_n = 0;
        ++li_n;
      }
    };
#pragma unroll
    for (int j = 0; j < NBUF - 1; ++j) {
      if (li_n < nsurv) {
        issue(li_n, i0_n, issued % NBUF);
        advance_n();
      }
      cp_async_commit();
      ++issued;
    }
    int li = 0, i0 = 0;
    while (li < nsurv) {
      if (li_n < nsurv) {
        issue(li_n, i0_n, issued % NBUF);
        advance_n();
      }
      cp_async_commit();
      ++issued;
      cp_async_wait<NBUF - 1>();
      __syncthreads();
      const int buf = consumed % NBUF;
      const PairLite &pr = s_pairs[li];
      const bool offdiag = pr.p != pr.q;
      if (i0 == 0) {
        if (pr.p != cur_p) {
          if (gp_dirty) if (!(P.debug & 1)) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
          cur_p = pr.p;
          cur_bfp = pr.bfp;
          gp_dirty = false;
          load_F<LA>(Fp, Fcol, pr.bfp, w, pure, P.c2s);
#pragma unroll
          for (int i = 0; i < NCA; ++i) Gp[i] = 0.0;
        }
        take_F<LB>(Fq, &s_F[buf][0][0][tid], w, pure, P.c2s);
#pragma unroll
        for (int j = 0; j < NCB; ++j) Gq[j] = 0.0;
        gp_dirty = gp_dirty || offdiag;
        if (P.stats && tid == 0) {
          atomicAdd(P.stats + 2 * cls, 1ULL);
          atomicAdd(P.stats + 2 * cls + 1, static_cast<unsigned long long>(pr.nprim) * P.counts[tile]);
        }
      }
      const int ns = min(STAGE, pr.nprim - i0);
#pragma unroll(L == 0 ? 4 : (L <= 2 ? 2 : 1))
      for (int s = 0; s < ((P.debug & 2) ? 0 : ns); ++s) {
        const double *rec = s_pool[buf] + s * STRIDE;
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
        prim_eval<LA, LB>(rec, S, PCx, PCy, PCz, Fp, Fq, Gp, Gq, offdiag);
      }
      i0 += STAGE;
      if (i0 >= pr.nprim) {
        if (!(P.debug & 1)) flush_G<LB>(Gq, Gcol, pr.bfq, pure, P.c2s);
        i0 = 0;
        ++li;
      }
      ++consumed;
      __syncthreads(); // the buffer just read is the target of the next issue
    }
    cp_async_wait<0>();
  }
  if (gp_dirty) if (!(P.debug & 1)) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
}

// Warp-private variant (COSX_WARP classes): after the CTA-wide screening round each WARP walks the survivor list on its
// own — 32 grid points, no CTA barrier and no shared-memory staging inside the pair loop.  The primitive records are read
// straight from global memory with warp-uniform (broadcast) loads, the F column values of the NEXT pair are fetched into
// registers while the current pair is evaluated.  Measured motive (valinomycin, OCC_COSX_DEBUG=2): with the primitive
// loop removed the staged pipeline still takes 734 of 1146 ms — two CTA barriers, the cp.async address arithmetic and the
// shared-memory round trip per step cost more than the integrals of the low-l classes.
template <int LA, int LB>
__global__ void __launch_bounds__(TILE_P, cosx_min_ctas(LA, LB)) k_cosx_esp_w(CosxParams P, int pair0, int npairs, int chunk, int cls,
                                                                               int ntiles, int nruns) {
  constexpr int L = LA + LB, NE = (LA + 1) * (LB + 1) * (L + 1), STRIDE = cosx_stride(LA, LB);
  constexpr int NCA = cx_ncart(LA), NCB = cx_ncart(LB), NH = cx_nhermsum(L);
  __shared__ PairLite s_pairs[TILE_P];
  __shared__ int s_wcnt[TILE_P / 32];
  const int tl = blockIdx.x % COSX_TILE_GROUP;
  const int yrun = (blockIdx.x / COSX_TILE_GROUP) % nruns;
  const int ltile = (blockIdx.x / (COSX_TILE_GROUP * nruns)) * COSX_TILE_GROUP + tl;
  if (ltile >= ntiles) return;
  const int tile = P.tile0 + ltile;
  if (P.tile_skip[tile]) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k0 = pair0 + yrun * chunk;
  const int k1 = min(k0 + chunk, pair0 + npairs);
  const double tcx = P.tcx[tile], tcy = P.tcy[tile], tcz = P.tcz[tile], lim = P.trad[tile] + P.margin;
  const size_t pt = static_cast<size_t>(tile) * TILE_P + tid;
  const double X = P.px[pt], Y = P.py[pt], Z = P.pz[pt], w = P.pw[pt];
  const size_t slab = static_cast<size_t>(ltile) * P.nbp * TILE_P + tid;
  const double *__restrict__ Fcol = P.F + slab;
  double *Gcol = P.G + slab;
  const double *__restrict__ pool = P.pool;
  const bool pure = P.pure != 0;
  const int nq_cols = (pure && LB > 0) ? 2 * LB + 1 : NCB;
  const double w2 = P.omega * P.omega;

  double Fp[NCA], Gp[NCA];
  int cur_p = -1, cur_bfp = 0;
  bool gp_dirty = false;
  for (int kr = k0; kr < k1; kr += TILE_P) {
    const int k = kr + tid;
    bool keep = false;
    CosxPair cand;
    if (k < k1 && P.pair_active[k]) {
      cand = P.pairs[k];
      const double dx = cand.cx - tcx, dy = cand.cy - tcy, dz = cand.cz - tcz;
      keep = !(sqrt(dx * dx + dy * dy + dz * dz) - cand.extent >= lim);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncthreads(); // every warp is done with the previous round's survivor list
    if (lane == 0) s_wcnt[wid] = __popc(bal);
    __syncthreads();
    int base = 0, nsurv = 0;
#pragma unroll
    for (int i = 0; i < TILE_P / 32; ++i) {
      if (i < wid) base += s_wcnt[i];
      nsurv += s_wcnt[i];
    }
    if (keep) s_pairs[base + __popc(bal & ((1u << lane) - 1u))] = PairLite{cand.p, cand.q, cand.bfp, cand.bfq, cand.nprim, cand.off};
    __syncthreads();
    if (nsurv == 0) continue;

    // ---- warp-private walk over the survivors ----
    double Fn[NCB]; // raw F columns of the next pair's q shell
    {
      const double *f = Fcol + static_cast<size_t>(s_pairs[0].bfq) * TILE_P;
#pragma unroll
      for (int c = 0; c < NCB; ++c) Fn[c] = c < nq_cols ? __ldg(f + static_cast<size_t>(c) * TILE_P) : 0.0;
    }
#pragma unroll 1
    for (int li = 0; li < nsurv; ++li) {
      const PairLite pr = s_pairs[li];
      const bool offdiag = pr.p != pr.q;
      double Fq[NCB], Gq[NCB];
      if (!pure || LB == 0) {
#pragma unroll
        for (int c = 0; c < NCB; ++c) Fq[c] = w * Fn[c];
      } else {
        const double *T = c2s_of<LB>(P.c2s);
#pragma unroll
        for (int c = 0; c < NCB; ++c) Fq[c] = 0.0;
#pragma unroll
        for (int m = 0; m < 2 * LB + 1; ++m) {
          const double f = w * Fn[m];
#pragma unroll
          for (int c = 0; c < NCB; ++c) Fq[c] += __ldg(T + m * NCB + c) * f;
        }
      }
      if (li + 1 < nsurv) { // the next pair's F values travel while this pair is evaluated
        const double *f = Fcol + static_cast<size_t>(s_pairs[li + 1].bfq) * TILE_P;
#pragma unroll
        for (int c = 0; c < NCB; ++c) Fn[c] = c < nq_cols ? __ldg(f + static_cast<size_t>(c) * TILE_P) : 0.0;
      }
      if (pr.p != cur_p) {
        if (gp_dirty && !(P.debug & 1)) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
        cur_p = pr.p;
        cur_bfp = pr.bfp;
        gp_dirty = false;
        load_F<LA>(Fp, Fcol, pr.bfp, w, pure, P.c2s);
#pragma unroll
        for (int i = 0; i < NCA; ++i) Gp[i] = 0.0;
      }
#pragma unroll
      for (int j = 0; j < NCB; ++j) Gq[j] = 0.0;
      gp_dirty = gp_dirty || offdiag;
      if (P.stats && tid == 0) {
        atomicAdd(P.stats + 2 * cls, 1ULL);
        atomicAdd(P.stats + 2 * cls + 1, static_cast<unsigned long long>(pr.nprim) * P.counts[tile]);
      }
      const double *__restrict__ rec0 = pool + pr.off;
      const int np = (P.debug & 2) ? 0 : pr.nprim;
#pragma unroll(L == 0 ? 4 : (L <= 2 ? 2 : 1))
      for (int i = 0; i < np; ++i) {
        const double *__restrict__ rec = rec0 + static_cast<size_t>(i) * STRIDE;
        const double pe = __ldg(rec);
        const double PCx = __ldg(rec + 1) - X, PCy = __ldg(rec + 2) - Y, PCz = __ldg(rec + 3) - Z;
        const double r2 = PCx * PCx + PCy * PCy + PCz * PCz;
        double S[L + 1];
#pragma unroll
        for (int m = 0; m <= L; ++m) S[m] = 0.0;
        if (P.w_full != 0.0) add_seeds<L>(P.boys, pe, r2, P.w_full, S);
        if (P.w_lr != 0.0) {
          const double sc = w2 / (pe + w2);
          add_seeds<L>(P.boys, pe * sc, r2, P.w_lr * sqrt(sc), S);
        }
        prim_eval<LA, LB>(rec, S, PCx, PCy, PCz, Fp, Fq, Gp, Gq, offdiag);
      }
      if (!(P.debug & 1)) flush_G<LB>(Gq, Gcol, pr.bfq, pure, P.c2s);
    }
  }
  if (gp_dirty && !(P.debug & 1)) flush_G<LA>(Gp, Gcol, cur_bfp, pure, P.c2s);
}

template <int LA, int LB>
void launch_class(const CosxParams &P, int pair0, int npairs, int ntiles, cudaStream_t st) {
  constexpr int L = LA + LB;
  int chunk = L == 0 ? 512 : (L == 1 ? 256 : 128);
  const int ngroups = (ntiles + COSX_TILE_GROUP - 1) / COSX_TILE_GROUP;
  while (static_cast<long long>(ngroups) * COSX_TILE_GROUP * ((npairs + chunk - 1) / chunk) > 0x7fffffffLL) chunk *= 2;
  const int nruns = (npairs + chunk - 1) / chunk;
  if constexpr (LA + LB <= COSX_WARP_MAXL) {
    if (!(P.debug & 256)) {
      k_cosx_esp_w<LA, LB><<<ngroups * COSX_TILE_GROUP * nruns, TILE_P, 0, st>>>(P, pair0, npairs, chunk,
                                                                                  LA * (COSX_MAX_L + 1) + LB, ntiles, nruns);
      return;
    }
  }
  k_cosx_esp<LA, LB><<<ngroups * COSX_TILE_GROUP * nruns, TILE_P, 0, st>>>(P, pair0, npairs, chunk,
                                                                              LA * (COSX_MAX_L + 1) + LB, ntiles, nruns);
}

} // namespace xcb
