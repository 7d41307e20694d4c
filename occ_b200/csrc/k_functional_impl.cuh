// k_functional_impl.cuh — the functional kernels of the XC build, one instantiation per (family, spin) case in its own
// translation unit (k_functional_f{0,1,2}s{1,2}.cu): the functional bodies with their dual-number derivatives are by far the
// slowest code of the library to compile (minutes in one unit), and the six cases are independent.
#pragma once
#include "xc_functionals.cuh"
#include "xc_internal.h"
#include <cuda_runtime.h>

namespace xcb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One out-of-line instance per (family, spin) case: the functional bodies are large, and both the build kernel
// and the standalone evaluation kernel call the same code.
template <int FAM> __device__ __noinline__ void eval_unpol(int nfunc, const int *ids, double rho, double sigma,
                                                           double tau, XCPoint *o) {
  xc_eval_unpolarized<FAM>(nfunc, ids, rho, sigma, tau, *o);
}
template <int FAM> __device__ __noinline__ void eval_pol(int nfunc, const int *ids, double ra, double rb, double saa,
                                                         double sab, double sbb, double ta, double tb, XCPoint *o) {
  xc_eval_polarized<FAM>(nfunc, ids, ra, rb, saa, sab, sbb, ta, tb, *o);
}

template <int FAM, int NSPIN>
__global__ void __launch_bounds__(TILE_P) k_functional(BuildParams P, int leaf0) {
  const int leaf = leaf0 + blockIdx.x;
  const LeafDesc lf = P.leaves[leaf];
  const int p = threadIdx.x, lane = p & 31, warp = p >> 5;
  __shared__ double red[4][4];

  // pass 1: electron counts and the leaf maximum of rho (a leaf is 1 tile, the last leaf up to 2)
  double na = 0.0, nbeta = 0.0, mx = -1e300;
#pragma unroll 1
  for (int k = 0; k < lf.ntiles; ++k) {
    const int tile = lf.tile0 + k;
    const int count = P.tiles[tile].count;
    const double w = P.pw[static_cast<size_t>(tile) * TILE_P + p];
    const double *a = P.acc + static_cast<size_t>(tile) * NSPIN * NACC * TILE_P + p;
    const double ra = a[0];
    na += ra * w;
    if (p < count) mx = fmax(mx, ra);
    if (NSPIN == 2) {
      const double rb = a[NACC * TILE_P];
      nbeta += rb * w;
      if (p < count) mx = fmax(mx, rb);
    }
  }
  mx = warp_max(mx);
  na = warp_sum(na);
  nbeta = warp_sum(nbeta);
  if (lane == 0) {
    red[0][warp] = mx;
    red[1][warp] = na;
    red[2][warp] = nbeta;
  }
  __syncthreads();
  mx = fmax(fmax(red[0][0], red[0][1]), fmax(red[0][2], red[0][3]));
  const bool skip = mx < P.density_threshold;

  // pass 2: functionals and Vxc coefficients
  double e = 0.0;
  if (!skip) {
#pragma unroll 1
    for (int k = 0; k < lf.ntiles; ++k) {
      const int tile = lf.tile0 + k;
      const double wk = P.pw[static_cast<size_t>(tile) * TILE_P + p];
      double rho[NSPIN], g[NSPIN][3], tau[NSPIN];
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) {
        const double *a = P.acc + (static_cast<size_t>(tile) * NSPIN + s) * NACC * TILE_P + p;
        rho[s] = a[0];
        g[s][0] = g[s][1] = g[s][2] = 0.0;
        tau[s] = 0.0;
        if (FAM >= 1) {
          g[s][0] = 2.0 * a[TILE_P];
          g[s][1] = 2.0 * a[2 * TILE_P];
          g[s][2] = 2.0 * a[3 * TILE_P];
        }
        if (FAM >= 2) tau[s] = 0.5 * a[4 * TILE_P];
      }
      XCPoint o;
      if (NSPIN == 1) {
        const double sigma = g[0][0] * g[0][0] + g[0][1] * g[0][1] + g[0][2] * g[0][2];
        eval_unpol<FAM>(P.nfunc, P.func_ids, rho[0], sigma, tau[0], &o);
      } else {
        constexpr int B = NSPIN - 1;
        const double saa = g[0][0] * g[0][0] + g[0][1] * g[0][1] + g[0][2] * g[0][2];
        const double sab = g[0][0] * g[B][0] + g[0][1] * g[B][1] + g[0][2] * g[B][2];
        const double sbb = g[B][0] * g[B][0] + g[B][1] * g[B][1] + g[B][2] * g[B][2];
        eval_pol<FAM>(P.nfunc, P.func_ids, rho[0], rho[B], saa, sab, sbb, tau[0], tau[B], &o);
      }
      double rt = rho[0];
      if (NSPIN == 2) rt += rho[NSPIN - 1];
      e += rt * (o.exc * wk);
#pragma unroll
      for (int s = 0; s < NSPIN; ++s) {
        double *cf = P.coef + (static_cast<size_t>(tile) * NSPIN + s) * NCOEF * TILE_P + p;
        cf[0] = 0.5 * (o.vrho[s] * wk);
        if (FAM >= 1) {
          if (NSPIN == 1) {
            const double vs = o.vsigma[0] * wk;
#pragma unroll
            for (int d = 0; d < 3; ++d) cf[(1 + d) * TILE_P] = 2.0 * (g[0][d] * vs);
          } else {
            const double vss = 2.0 * (o.vsigma[s == 0 ? 0 : 2] * wk), vab = o.vsigma[1] * wk;
            const int other = NSPIN - 1 - s;
#pragma unroll
            for (int d = 0; d < 3; ++d) cf[(1 + d) * TILE_P] = g[s][d] * vss + g[other][d] * vab;
          }
        }
        if (FAM >= 2) cf[4 * TILE_P] = 0.25 * (o.vtau[s] * wk);
      }
    }
  }
  e = warp_sum(e);
  if (lane == 0) red[3][warp] = e;
  __syncthreads();
  if (p == 0) {
    P.leaf_skip[leaf] = skip ? 1 : 0;
    double *ls = P.leaf_sums + 3 * static_cast<size_t>(leaf);
    ls[0] = (red[3][0] + red[3][1]) + (red[3][2] + red[3][3]);
    ls[1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    ls[2] = (red[2][0] + red[2][1]) + (red[2][2] + red[2][3]);
  }
}

// ---- standalone pointwise evaluation (DensityFunctional::evaluate analogue for tests / callers) ----
template <int FAM, bool POL>
__global__ void k_eval_points(int id, size_t n, const double *__restrict__ rho, const double *__restrict__ sigma,
                              const double *__restrict__ tau, double *__restrict__ exc, double *__restrict__ vrho,
                              double *__restrict__ vsigma, double *__restrict__ vtau) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XCPoint o;
  if (!POL) {
    eval_unpol<FAM>(1, &id, rho[i], FAM >= 1 ? sigma[i] : 0.0, FAM >= 2 ? tau[i] : 0.0, &o);
    exc[i] = o.exc;
    vrho[i] = o.vrho[0];
    if (FAM >= 1) vsigma[i] = o.vsigma[0];
    if (FAM >= 2) vtau[i] = o.vtau[0];
  } else {
    eval_pol<FAM>(1, &id, rho[2 * i], rho[2 * i + 1], FAM >= 1 ? sigma[3 * i] : 0.0,
                  FAM >= 1 ? sigma[3 * i + 1] : 0.0, FAM >= 1 ? sigma[3 * i + 2] : 0.0,
                  FAM >= 2 ? tau[2 * i] : 0.0, FAM >= 2 ? tau[2 * i + 1] : 0.0, &o);
    exc[i] = o.exc;
    vrho[2 * i] = o.vrho[0];
    vrho[2 * i + 1] = o.vrho[1];
    if (FAM >= 1) {
      vsigma[3 * i] = o.vsigma[0];
      vsigma[3 * i + 1] = o.vsigma[1];
      vsigma[3 * i + 2] = o.vsigma[2];
    }
    if (FAM >= 2) {
      vtau[2 * i] = o.vtau[0];
      vtau[2 * i + 1] = o.vtau[1];
    }
  }
}

template <int FAM, int NSPIN> void launch_functional_t(const BuildParams &P, int leaf0, int nleaves, cudaStream_t st) {
  k_functional<FAM, NSPIN><<<nleaves, TILE_P, 0, st>>>(P, leaf0);
}
template <int FAM, bool POL>
void launch_eval_points_t(unsigned grid, int bs, int id, size_t n, const double *rho, const double *sigma, const double *tau,
                          double *exc, double *vrho, double *vsigma, double *vtau) {
  k_eval_points<FAM, POL><<<grid, bs>>>(id, n, rho, sigma, tau, exc, vrho, vsigma, vtau);
}

} // namespace xcb
