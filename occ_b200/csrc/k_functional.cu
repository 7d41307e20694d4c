// k_functional.cu — phase 3 of the XC build, one CTA per Morton leaf, one thread per grid point:
//   * finish rho, grad rho, tau from the density-kernel accumulators (density.cpp: the 2x on gradients, 1/2 on tau)
//   * electron counts BEFORE the skip test, then the leaf-level density skip (src/dft/dft_kernels.cpp:27-40)
//   * sigma (include/occ/dft/dft_kernels.h:15-59), the functionals summed over components
//     (src/dft/functional.cpp:18-65, functional.h:438-459), weighting (functional.h:460-470)
//   * Exc contribution (xc_potential_matrix.cpp: energy += rho . (w exc)) with a warp-shuffle + block reduction
//   * the per-point coefficients of the Vxc contraction (xc_potential_matrix.cpp:18-135), stored so that
//     V = Ktmp + Ktmp^T with Ktmp = phi^T (c0 phi + cx phi_x + cy phi_y + cz phi_z) + sum_k phi_k^T (ct phi_k)
// plus the small elementwise kernels around the build (density-matrix preparation, final symmetrisation and
// deterministic reduction of the per-leaf sums) and the standalone pointwise functional evaluation.
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

void launch_functional(const BuildParams &P, int leaf0, int nleaves, void *stream) {
  if (nleaves <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int fam = P.deriv;
#define XC_LAUNCH(F, S) k_functional<F, S><<<nleaves, TILE_P, 0, st>>>(P, leaf0)
  if (P.nspin == 1) {
    if (fam == 0) XC_LAUNCH(0, 1);
    else if (fam == 1) XC_LAUNCH(1, 1);
    else XC_LAUNCH(2, 1);
  } else {
    if (fam == 0) XC_LAUNCH(0, 2);
    else if (fam == 1) XC_LAUNCH(1, 2);
    else XC_LAUNCH(2, 2);
  }
#undef XC_LAUNCH
}

// ---- density-matrix preparation: restricted builds read only the upper triangle of D
// (D.selfadjointView<Upper>(), density.cpp:14,38,78); unrestricted builds use D as given ----
__global__ void k_symmetrize_upper(const double *__restrict__ D, double *__restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n) return;
  const int r = min(i, j), c = max(i, j);
  out[static_cast<size_t>(j) * n + i] = D[static_cast<size_t>(c) * n + r];
}
void launch_prepare_density(const double *D2, double *Dsym, int nbf, int nspin, void *stream) {
  if (nspin != 1) return;
  dim3 grid((nbf + 255) / 256, nbf);
  k_symmetrize_upper<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(D2, Dsym, nbf);
}

// ---- finalise: K = Ktmp + Ktmp^T per spin block, and the deterministic (fixed-order) reduction of the
// per-leaf sums (the reference's serial reduction over thread copies, dft.h:375-380) ----
__global__ void k_finalize_K(BuildParams P, double *__restrict__ K) {
  __shared__ double tl[32][33];
  const int bi = blockIdx.x * 32, bj = blockIdx.y * 32, s = blockIdx.z;
  const double *Kt = P.Ktmp[s];
  const int n = P.nbf, ld = P.ldD;
  const int tx = threadIdx.x, ty = threadIdx.y; // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bj + tx, j = bi + r; // transposed block element Kt(i, j) with i from the j-block
    tl[r][tx] = (i < n && j < n) ? Kt[static_cast<size_t>(j) * ld + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + tx, j = bj + r;
    if (i < n && j < n) {
      const size_t o = static_cast<size_t>(j) * ld + i;
      K[o + static_cast<size_t>(s) * n] = Kt[o] + tl[tx][r];
    }
  }
}
__global__ void k_reduce_scalars(const double *__restrict__ leaf_sums, int nleaves, double *__restrict__ out) {
  __shared__ double sh[3][256];
  const int t = threadIdx.x;
  double a[3] = {0.0, 0.0, 0.0};
  for (int l = t; l < nleaves; l += 256)
    for (int k = 0; k < 3; ++k) a[k] += leaf_sums[3 * static_cast<size_t>(l) + k];
  for (int k = 0; k < 3; ++k) sh[k][t] = a[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o)
      for (int k = 0; k < 3; ++k) sh[k][t] += sh[k][t + o];
    __syncthreads();
  }
  if (t < 3) out[t] = sh[t][0];
}
void launch_finalize(const BuildParams &P, int nleaves, double *K, double *scalars, void *stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((P.nbf + 31) / 32, (P.nbf + 31) / 32, P.nspin);
  k_finalize_K<<<grid, dim3(32, 8), 0, st>>>(P, K);
  k_reduce_scalars<<<1, 256, 0, st>>>(P.leaf_sums, nleaves, scalars);
}

// ---- density on the grid (occ::density::evaluate_density_on_grid, include/occ/gto/density.h:67-107): turn the
// density-kernel accumulators into the reference's rho matrix, (nspin * npts) x {1, 4, 6} column-major with the
// alpha rows first; columns rho, rho_x, rho_y, rho_z, laplacian, tau ----
__global__ void __launch_bounds__(TILE_P) k_gather_rho(BuildParams P, const size_t *__restrict__ tile_src, size_t npts,
                                                       double *__restrict__ out) {
  const int tile = blockIdx.x, p = threadIdx.x;
  if (p >= P.tiles[tile].count) return;
  const size_t si = tile_src[tile] + p;
  const size_t ld = static_cast<size_t>(P.nspin) * npts;
  for (int s = 0; s < P.nspin; ++s) {
    const double *a = P.acc + (static_cast<size_t>(tile) * P.nspin + s) * NACC * TILE_P + p;
    double *o = out + static_cast<size_t>(s) * npts + si;
    o[0] = a[0];
    if (P.deriv >= 1) {
      o[ld] = 2.0 * a[TILE_P];
      o[2 * ld] = 2.0 * a[2 * TILE_P];
      o[3 * ld] = 2.0 * a[3 * TILE_P];
    }
    if (P.deriv >= 2) {
      // laplacian = 2 rowsum(lap phi o D phi) + 2 sum_d rowsum(phi_d o (phi_d D)) (density.cpp:84-100): the second term
      // here, the first one is added by k_gather_laplacian after the second density pass
      o[4 * ld] = 2.0 * a[4 * TILE_P];
      o[5 * ld] = 0.5 * a[4 * TILE_P];
    }
  }
}
// second pass of a deriv-2 density-on-grid call: the AO slabs hold (lap phi, phi, ...), so accumulator row 1 of the
// ordinary density items is rowsum(phi o (lap phi . D))
__global__ void __launch_bounds__(TILE_P) k_gather_laplacian(BuildParams P, const size_t *__restrict__ tile_src,
                                                             size_t npts, double *__restrict__ out) {
  const int tile = blockIdx.x, p = threadIdx.x;
  if (p >= P.tiles[tile].count) return;
  const size_t si = tile_src[tile] + p;
  const size_t ld = static_cast<size_t>(P.nspin) * npts;
  for (int s = 0; s < P.nspin; ++s) {
    const double *a = P.acc + (static_cast<size_t>(tile) * P.nspin + s) * NACC * TILE_P + p;
    out[static_cast<size_t>(s) * npts + si + 4 * ld] += 2.0 * a[TILE_P];
  }
}
void launch_gather_laplacian(const BuildParams &P, const size_t *tile_src, size_t npts, int ntiles, double *out, void *stream) {
  if (ntiles > 0) k_gather_laplacian<<<ntiles, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(P, tile_src, npts, out);
}
void launch_gather_rho(const BuildParams &P, const size_t *tile_src, size_t npts, int ntiles, double *out, void *stream) {
  if (ntiles > 0) k_gather_rho<<<ntiles, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(P, tile_src, npts, out);
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

int eval_functional_device(int id, int polarized, size_t n, const double *rho, const double *sigma,
                           const double *tau, double *exc, double *vrho, double *vsigma, double *vtau) {
  const int fam = xc_family_of(id);
  if (fam < 0) return 1;
  if (n == 0) return 0;
  const int s = polarized ? 2 : 1, ns = polarized ? 3 : 1;
  double *d_rho = nullptr, *d_sigma = nullptr, *d_tau = nullptr, *d_exc = nullptr, *d_vrho = nullptr,
         *d_vsigma = nullptr, *d_vtau = nullptr;
  cudaError_t err = cudaSuccess;
  auto up = [&](double **d, const double *h, size_t cnt) {
    if (err != cudaSuccess) return;
    err = cudaMalloc(d, cnt * sizeof(double));
    if (err == cudaSuccess && h) err = cudaMemcpy(*d, h, cnt * sizeof(double), cudaMemcpyHostToDevice);
  };
  up(&d_rho, rho, n * s);
  if (fam >= 1) up(&d_sigma, sigma, n * ns);
  if (fam >= 2) up(&d_tau, tau, n * s);
  up(&d_exc, nullptr, n);
  up(&d_vrho, nullptr, n * s);
  if (fam >= 1) up(&d_vsigma, nullptr, n * ns);
  if (fam >= 2) up(&d_vtau, nullptr, n * s);
  if (err == cudaSuccess) {
    const int bs = 128;
    const unsigned grid = static_cast<unsigned>((n + bs - 1) / bs);
#define XC_EV(F, Pp) k_eval_points<F, Pp><<<grid, bs>>>(id, n, d_rho, d_sigma, d_tau, d_exc, d_vrho, d_vsigma, d_vtau)
    if (!polarized) {
      if (fam == 0) XC_EV(0, false);
      else if (fam == 1) XC_EV(1, false);
      else XC_EV(2, false);
    } else {
      if (fam == 0) XC_EV(0, true);
      else if (fam == 1) XC_EV(1, true);
      else XC_EV(2, true);
    }
#undef XC_EV
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
  }
  auto down = [&](double *h, const double *d, size_t cnt) {
    if (err == cudaSuccess && h && d) err = cudaMemcpy(h, d, cnt * sizeof(double), cudaMemcpyDeviceToHost);
  };
  down(exc, d_exc, n);
  down(vrho, d_vrho, n * s);
  if (fam >= 1) down(vsigma, d_vsigma, n * ns);
  if (fam >= 2) down(vtau, d_vtau, n * s);
  cudaFree(d_rho); cudaFree(d_sigma); cudaFree(d_tau); cudaFree(d_exc); cudaFree(d_vrho); cudaFree(d_vsigma);
  cudaFree(d_vtau);
  return err == cudaSuccess ? 0 : 2;
}

} // namespace xcb
