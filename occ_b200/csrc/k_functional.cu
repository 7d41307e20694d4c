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

// the functional kernels live in k_functional_impl.cuh, one (family, spin) instantiation per translation unit
template <int FAM, int NSPIN> void launch_functional_t(const BuildParams &P, int leaf0, int nleaves, cudaStream_t st);
template <int FAM, bool POL>
void launch_eval_points_t(unsigned grid, int bs, int id, size_t n, const double *rho, const double *sigma, const double *tau,
                          double *exc, double *vrho, double *vsigma, double *vtau);
#define XC_DECL(F) \
  extern template void launch_functional_t<F, 1>(const BuildParams &, int, int, cudaStream_t); \
  extern template void launch_functional_t<F, 2>(const BuildParams &, int, int, cudaStream_t); \
  extern template void launch_eval_points_t<F, false>(unsigned, int, int, size_t, const double *, const double *, const double *, double *, double *, double *, double *); \
  extern template void launch_eval_points_t<F, true>(unsigned, int, int, size_t, const double *, const double *, const double *, double *, double *, double *, double *);
XC_DECL(0)
XC_DECL(1)
XC_DECL(2)
#undef XC_DECL

void launch_functional(const BuildParams &P, int leaf0, int nleaves, void *stream) {
  if (nleaves <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int fam = P.deriv;
#define XC_LAUNCH(F, S) launch_functional_t<F, S>(P, leaf0, nleaves, st)
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
#define XC_EV(F, Pp) launch_eval_points_t<F, Pp>(grid, bs, id, n, d_rho, d_sigma, d_tau, d_exc, d_vrho, d_vsigma, d_vtau)
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
