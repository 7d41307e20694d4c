// k_vv10.cu — post-SCF VV10 non-local correlation on the GPU.
// Reference: vv10_kernel, src/dft/nonlocal_correlation.cpp:38-118 and NonLocalCorrelationFunctional::vv10, :367-420.
//   * rho / grad rho on the NLC grid come from the same collocation + DMMA density kernels as the XC build
//   * the O(N^2) pair sum  F_i, U_i, W_i = sum_{j<i} ...  is a tiled all-pairs kernel: thread = point i, the j side is
//     staged through shared memory 128 points at a time; ONE reciprocal per pair (1/(g_i g_j (g_i+g_j)), the other two
//     quotients of the reference are products with it).  "j < i" is taken in the order the REFERENCE enumerates the
//     points (its concatenated atom grids): the per-point arrays are therefore laid out in that original order and the
//     triangular structure is exploited (work items (i-block, chunk of j-blocks) with j-block <= i-block).
//   * the resulting vrho, vsigma become the per-point coefficients of the same resident-B Vxc contraction.
// FP64 CUDA-core bound (about 30 DFMA-class instructions per pair); no tensor-core shape in this sum.
#include "xc_internal.h"
#include <cuda_runtime.h>

namespace xcb {

constexpr int VV_JCHUNK = 32; // j-blocks (of 128 points) per work item

// every point starts inactive: points of leaves that keep no shell (dropped at create, dft.h:309-321) have no tile
__global__ void k_vv10_init(VV10Params V) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= V.n) return;
  V.X[i] = V.Y[i] = V.Z[i] = 0.0;
  V.rw2[i] = 0.0;
  V.om[i] = 1.0;
  V.kap[i] = -1.0;
  V.F[i] = V.U[i] = V.W[i] = 0.0;
}

__global__ void __launch_bounds__(TILE_P) k_vv10_prepare(BuildParams P, VV10Params V) {
  const int tile = blockIdx.x, p = threadIdx.x;
  const TileDesc t = P.tiles[tile];
  if (p >= t.count) return;
  const size_t si = V.tile_src[tile] + p;
  const size_t oi = V.orig ? V.orig[si] : si;
  const size_t slot = static_cast<size_t>(tile) * TILE_P + p;
  const double *a = P.acc + static_cast<size_t>(tile) * NACC * TILE_P + p;
  const double rho = a[0];
  const double gx = 2.0 * a[TILE_P], gy = 2.0 * a[2 * TILE_P], gz = 2.0 * a[3 * TILE_P];
  const double g = gx * gx + gy * gy + gz * gz;
  V.X[oi] = P.px[slot];
  V.Y[oi] = P.py[slot];
  V.Z[oi] = P.pz[slot];
  V.F[oi] = 0.0;
  V.U[oi] = 0.0;
  V.W[oi] = 0.0;
  if (rho < 1e-8) { // neither receives nor contributes (nonlocal_correlation.cpp:71-72,81-82)
    V.rw2[oi] = 0.0;
    V.om[oi] = 1.0;
    V.kap[oi] = -1.0; // marks "inactive" for the i side; j side is neutralised by rw2 = 0 and uses |kap|
    return;
  }
  double tq = g / (rho * rho);
  tq = V.C * tq * tq;
  V.rw2[oi] = 2.0 * (rho * P.pw[slot]);
  V.om[oi] = sqrt(tq + (4.0 * 3.14159265358979323846 / 3.0) * rho);
  V.kap[oi] = V.kappa_factor * pow(rho, 1.0 / 6.0);
}

// work item = (i-block bi, j-chunk jc): j-blocks [jc*VV_JCHUNK, min((jc+1)*VV_JCHUNK, bi+1))
__global__ void __launch_bounds__(TILE_P) k_vv10_pairs(VV10Params V, const int2 *__restrict__ items) {
  __shared__ double sx[TILE_P], sy[TILE_P], sz[TILE_P], srw[TILE_P], som[TILE_P], skp[TILE_P];
  const int2 it = items[blockIdx.x];
  const int bi = it.x, tid = threadIdx.x;
  const size_t i = static_cast<size_t>(bi) * TILE_P + tid;
  const bool ivalid = i < V.n;
  double xi = 0, yi = 0, zi = 0, omi = 1, kpi = -1;
  if (ivalid) {
    xi = V.X[i];
    yi = V.Y[i];
    zi = V.Z[i];
    omi = V.om[i];
    kpi = V.kap[i];
  }
  const bool active = kpi > 0.0;
  if (!active) kpi = 1.0;
  double f = 0.0, u = 0.0, w = 0.0;
  const int jb0 = it.y * VV_JCHUNK, jb1 = min(jb0 + VV_JCHUNK, bi + 1);
  for (int jb = jb0; jb < jb1; ++jb) {
    const size_t j = static_cast<size_t>(jb) * TILE_P + tid;
    __syncthreads();
    if (j < V.n) {
      sx[tid] = V.X[j];
      sy[tid] = V.Y[j];
      sz[tid] = V.Z[j];
      srw[tid] = V.rw2[j];
      som[tid] = V.om[j];
      skp[tid] = fabs(V.kap[j]);
    } else {
      sx[tid] = sy[tid] = sz[tid] = 0.0;
      srw[tid] = 0.0;
      som[tid] = skp[tid] = 1.0;
    }
    __syncthreads();
    const int jmax = jb == bi ? tid : TILE_P; // strictly j < i on the diagonal block
#pragma unroll 4
    for (int jj = 0; jj < jmax; ++jj) {
      const double dx = xi - sx[jj], dy = yi - sy[jj], dz = zi - sz[jj];
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double gi = fma(r2, omi, kpi), gj = fma(r2, som[jj], skp[jj]);
      const double s = gi + gj;
      const double inv = 1.0 / (gi * gj * s);
      const double t = srw[jj] * inv;             // 2 rho_j w_j / (gi gj (gi + gj))
      f += t;
      const double t2 = t * (gj * (s + gi) * inv); // t * (1/gi + 1/(gi + gj))
      u += t2;
      w = fma(t2, r2, w);
    }
  }
  if (ivalid && active) {
    if (it.y == 0) { // self term seeded once (nonlocal_correlation.cpp:75-77)
      const double rwi = 0.5 * V.rw2[i];
      const double f0 = rwi / (2.0 * kpi * kpi * kpi);
      f += f0;
      u += f0 * (1.0 / kpi + 1.0 / (2.0 * kpi));
    }
    atomicAdd(V.F + i, f);
    atomicAdd(V.U + i, u);
    atomicAdd(V.W + i, w);
  }
}

__device__ __forceinline__ double vv_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one CTA per leaf: exc, vrho, vsigma (weighted), the energy sum and the Vxc coefficients of xc_potential_matrix's
// restricted GGA form (c0 = 1/2 vrho, c_d = 2 vsigma grad_d rho), nonlocal_correlation.cpp:105-117,408-419
__global__ void __launch_bounds__(TILE_P) k_vv10_coef(BuildParams P, VV10Params V) {
  const int leaf = blockIdx.x, p = threadIdx.x, lane = p & 31, warp = p >> 5;
  const LeafDesc lf = P.leaves[leaf];
  __shared__ double red[2][4];
  double e = 0.0, ne = 0.0;
  for (int k = 0; k < lf.ntiles; ++k) {
    const int tile = lf.tile0 + k;
    const int count = P.tiles[tile].count;
    double *cf = P.coef + static_cast<size_t>(tile) * NCOEF * TILE_P + p;
    double c0 = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    if (p < count) {
      const size_t si = V.tile_src[tile] + p;
      const size_t oi = V.orig ? V.orig[si] : si;
      const double wk = P.pw[static_cast<size_t>(tile) * TILE_P + p];
      const double *a = P.acc + static_cast<size_t>(tile) * NACC * TILE_P + p;
      const double rho = a[0];
      const double gx = 2.0 * a[TILE_P], gy = 2.0 * a[2 * TILE_P], gz = 2.0 * a[3 * TILE_P];
      double vrho = V.beta * wk, vsig = 0.0, exc = V.beta * wk;
      if (rho >= 1e-8) {
        const double g = gx * gx + gy * gy + gz * gz;
        const double F = -1.5 * V.F[oi], U = V.U[oi], W = V.W[oi];
        double tq = g / (rho * rho);
        tq = V.C * tq * tq;
        const double pi43 = 4.0 * 3.14159265358979323846 / 3.0;
        const double om = sqrt(tq + pi43 * rho);
        const double dom_dr = (0.5 * pi43 * rho - 2.0 * tq) / om;
        const double dkap_dr = V.kappa_factor * pow(rho, 1.0 / 6.0) / 6.0;
        exc = (V.beta + 0.5 * F) * wk;
        vrho = (V.beta + F + 1.5 * (U * dkap_dr + W * dom_dr)) * wk;
        vsig = W == 0.0 ? 0.0 : (1.5 * W * (tq * rho / (g * om))) * wk;
      }
      e += exc * rho;
      ne += rho * wk;
      c0 = 0.5 * vrho;
      cx = 2.0 * (gx * vsig);
      cy = 2.0 * (gy * vsig);
      cz = 2.0 * (gz * vsig);
    }
    cf[0] = c0;
    cf[TILE_P] = cx;
    cf[2 * TILE_P] = cy;
    cf[3 * TILE_P] = cz;
    cf[4 * TILE_P] = 0.0;
  }
  e = vv_warp_sum(e);
  ne = vv_warp_sum(ne);
  if (lane == 0) {
    red[0][warp] = e;
    red[1][warp] = ne;
  }
  __syncthreads();
  if (p == 0) {
    P.leaf_skip[leaf] = 0;
    double *ls = P.leaf_sums + 3 * static_cast<size_t>(leaf);
    ls[0] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    ls[1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    ls[2] = 0.0;
  }
}

void launch_vv10_prepare(const BuildParams &P, const VV10Params &V, int ntiles, void *stream) {
  k_vv10_init<<<static_cast<unsigned>((V.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(V);
  if (ntiles > 0) k_vv10_prepare<<<ntiles, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(P, V);
}
void launch_vv10_pairs(const VV10Params &V, const int2 *items, int nitems, void *stream) {
  if (nitems > 0) k_vv10_pairs<<<nitems, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(V, items);
}
void launch_vv10_coef(const BuildParams &P, const VV10Params &V, int nleaves, void *stream) {
  if (nleaves > 0) k_vv10_coef<<<nleaves, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(P, V);
}
int vv10_jchunk() { return VV_JCHUNK; }

} // namespace xcb
