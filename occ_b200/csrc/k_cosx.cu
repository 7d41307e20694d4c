// k_cosx.cu — kernels of the seminumerical exchange (COSX) build, SURVEY 8(f) rank 2.
//
// Reference: seminumerical_exchange_sweep (src/qm/seminumerical_exchange.cpp:660-815): for every grid batch and every
// significant shell pair (p >= q) that survives the geometric and the density screen, the ESP integrals
//     V_ab(g) = sum_prim prefactor * sum_tuv E^ab_tuv R_tuv(p, P - r_g)              (src/ints/esp.cpp:102-125)
// are contracted with F = (w phi) D2q:   G[g, q-block] += F[g, p-block] . V(g),  G[g, p-block] += F[g, q-block] . V(g)^T.
// The reference forms the dense product R (npt x nherm) . E^T (nherm x nab) per primitive pair; E is the outer product
// of three 1D McMurchie-Davidson tables, so here the contraction runs over that product structure directly
// (sum_t E^x_t sum_u E^y_u sum_v E^z_v R_tuv), 4-7x fewer flops than the dense form for d/f pairs.
//
// B200 mapping: FP64 CUDA-core work (B200's DFMA rate, 33.9 TFLOP/s measured, is within 10 % of its FP64 tensor rate, so
// the sparse form on the FMA pipe beats the dense form on DMMA). One thread per grid point, one CTA per (batch, run of
// 128-512 shell pairs of one (la >= lb) class; the higher-l shell is the register-resident side). Per round the CTA
// screens 128 candidate pairs in parallel and compacts the survivors in order; their primitive records (exponent, centre,
// 1D E tables) and F columns stream through a cp.async ring in shared memory. Boys values come from the reference's
// Chebyshev table for the highest order and the stable downward recursion for the lower ones; for L = la + lb <= 3 the
// Hermite integrals R_tuv live in registers, for L >= 4 they are produced and contracted slice by slice (k_cosx_esp.cuh).
// Every index is a compile-time constant (one instantiation per class, spread over k_cosx_la*.cu so that they compile in
// parallel). F and G move as coalesced 1 KiB columns of the batch slabs, G through red.global.add.f64.
#include "k_cosx_esp.cuh"

namespace xcb {

extern template void launch_class<0, 0>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<1, 0>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<1, 1>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<2, 0>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<2, 1>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<2, 2>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<3, 0>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<3, 1>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<3, 2>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<3, 3>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<4, 0>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<4, 1>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<4, 2>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<4, 3>(const CosxParams &, int, int, int, cudaStream_t);
extern template void launch_class<4, 4>(const CosxParams &, int, int, int, cudaStream_t);

void launch_cosx_esp(const CosxParams &P, int la, int lb, int pair0, int npairs, int ntiles, void *stream) {
  if (npairs <= 0 || ntiles <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define COSX_CASE(A, B)                                                                                      \
  case (A) * 8 + (B):                                                                                        \
    launch_class<A, B>(P, pair0, npairs, ntiles, st);                                                        \
    break;
  switch (la * 8 + lb) {
    COSX_CASE(0, 0)   
    COSX_CASE(1, 0) COSX_CASE(1, 1)  
    COSX_CASE(2, 0) COSX_CASE(2, 1) COSX_CASE(2, 2) 
    COSX_CASE(3, 0) COSX_CASE(3, 1) COSX_CASE(3, 2) COSX_CASE(3, 3)
    COSX_CASE(4, 0) COSX_CASE(4, 1) COSX_CASE(4, 2) COSX_CASE(4, 3) COSX_CASE(4, 4)
  default: break;
  }
#undef COSX_CASE
}

// seminumerical_exchange.cpp:729-730: a batch is skipped when max(phi) < precision (signed maximum)
__global__ void __launch_bounds__(256) k_cosx_tile_flags(const double *__restrict__ phi, const int *__restrict__ counts,
                                                         int nbf, int nbp, double precision, int *__restrict__ skip) {
  const int tile = blockIdx.x;
  const double *s = phi + static_cast<size_t>(tile) * nbp * TILE_P;
  const int cnt = counts[tile];
  double m = -1.0e300;
  for (int i = threadIdx.x; i < nbf * TILE_P; i += 256)
    if ((i & (TILE_P - 1)) < cnt) m = fmax(m, s[i]);
  __shared__ double red[256];
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) skip[tile] = red[0] < precision ? 1 : 0;
}
void launch_cosx_tile_flags(const double *phi, const int *counts, int ntiles, int nbf, int nbp, double precision,
                            int *tile_skip, void *stream) {
  if (ntiles <= 0) return;
  k_cosx_tile_flags<<<ntiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(phi, counts, nbf, nbp, precision, tile_skip);
}

// seminumerical_exchange.cpp:758-764: max |D2q| over the (p,q) block and, for p != q, the (q,p) block
__global__ void k_cosx_pair_flags(const CosxPair *__restrict__ pairs, const int *__restrict__ shell_size, int npairs,
                                  const double *__restrict__ D, int nbf, double thr, unsigned char *__restrict__ act) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= npairs) return;
  const CosxPair pr = pairs[k];
  const int sp = shell_size[pr.p], sq = shell_size[pr.q];
  double m = 0.0;
  for (int j = 0; j < sq; ++j)
    for (int i = 0; i < sp; ++i) {
      m = fmax(m, fabs(D[(pr.bfp + i) + static_cast<size_t>(pr.bfq + j) * nbf]));
      if (pr.p != pr.q) m = fmax(m, fabs(D[(pr.bfq + j) + static_cast<size_t>(pr.bfp + i) * nbf]));
    }
  act[k] = (m < thr) ? 0 : 1;
}
void launch_cosx_pair_flags(const CosxPair *pairs, const int *shell_size, int npairs, const double *D2q, int nbf,
                            double threshold, unsigned char *active, void *stream) {
  if (npairs <= 0) return;
  k_cosx_pair_flags<<<(npairs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(pairs, shell_size, npairs, D2q,
                                                                                       nbf, threshold, active);
}

// K = factor * (A + B^T): A = B = K' for the spin-diagonal blocks (seminumerical_exchange.cpp:529, 645-646), A = K'_ab,
// B = K'_ba for the off-diagonal blocks of a general spinorbital build (:640-645 of compute_K_general)
__global__ void k_cosx_symmetrize(const double *__restrict__ Kraw, const double *__restrict__ KrawT, int nbf,
                                  double factor, double *__restrict__ K, int ldK) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 32 + threadIdx.y;
  __shared__ double t[32][33];
  const int ti = blockIdx.y * 32 + threadIdx.x, tj = blockIdx.x * 32 + threadIdx.y;
  t[threadIdx.y][threadIdx.x] = (ti < nbf && tj < nbf) ? KrawT[ti + static_cast<size_t>(tj) * nbf] : 0.0;
  __syncthreads();
  if (i < nbf && j < nbf)
    K[i + static_cast<size_t>(j) * ldK] = factor * (Kraw[i + static_cast<size_t>(j) * nbf] + t[threadIdx.x][threadIdx.y]);
}
void launch_cosx_symmetrize(const double *Kraw, const double *KrawT, int nbf, double factor, double *K, int ldK,
                            void *stream) {
  dim3 grid((nbf + 31) / 32, (nbf + 31) / 32), block(32, 32);
  k_cosx_symmetrize<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(Kraw, KrawT, nbf, factor, K, ldK);
}

// (w phi) for the numerical overlap (seminumerical_exchange.cpp:366-369)
__global__ void k_cosx_scale_weights(double *__restrict__ out, const double *__restrict__ phi,
                                     const double *__restrict__ pw, int nbf, int nbp) {
  const int tile = blockIdx.x;
  const size_t base = static_cast<size_t>(tile) * nbp * TILE_P;
  const double w = pw[static_cast<size_t>(tile) * TILE_P + threadIdx.x];
  for (int c = blockIdx.y; c < nbf; c += gridDim.y)
    out[base + static_cast<size_t>(c) * TILE_P + threadIdx.x] = w * phi[base + static_cast<size_t>(c) * TILE_P + threadIdx.x];
}
void launch_cosx_scale_weights(double *phi_w, const double *phi, const double *pw, int ntiles, int nbf, int nbp,
                               void *stream) {
  if (ntiles <= 0) return;
  dim3 grid(ntiles, nbf < 64 ? nbf : 64);
  k_cosx_scale_weights<<<grid, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(phi_w, phi, pw, nbf, nbp);
}

} // namespace xcb
