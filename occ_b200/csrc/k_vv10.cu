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
// `unowned` is the kappa written for them: -1 (inactive) on a handle that holds the whole grid, 0 ("not this rank's point":
// resolved by k_vv10_merge) on a shard of a multi-GPU call
__global__ void k_vv10_init(VV10Params V, double unowned) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= V.n) return;
  V.X[i] = V.Y[i] = V.Z[i] = 0.0;
  V.rw2[i] = 0.0;
  V.om[i] = 1.0;
  V.kap[i] = unowned;
  V.F[i] = V.U[i] = V.W[i] = 0.0;
}

// multi-GPU: every point is owned by at most one shard (the one whose kappa is non-zero); gather its six values
__global__ void k_vv10_merge(VV10Peers src, VV10Params out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= out.n) return;
  double x = 0.0, y = 0.0, z = 0.0, rw = 0.0, om = 1.0, kp = -1.0;
  for (int r = 0; r < src.n; ++r) {
    const double *b = src.base[r]; // X Y Z rw2 om kap, each of length n
    const double k = b[5 * out.n + i];
    if (k != 0.0) {
      x = b[i];
      y = b[out.n + i];
      z = b[2 * out.n + i];
      rw = b[3 * out.n + i];
      om = b[4 * out.n + i];
      kp = k;
      break;
    }
  }
  out.X[i] = x;
  out.Y[i] = y;
  out.Z[i] = z;
  out.rw2[i] = rw;
  out.om[i] = om;
  out.kap[i] = kp;
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

// ------------------------------------------------------------------------------------------------
// VV10 nuclear gradient, grid-displacement term (compute_vv10_grid_displacement_gradient,
// nonlocal_correlation.cpp:193-328).  The reference visits every ordered pair (i, j != i) within 15 Bohr and does
//   grad[atom(i)] -= f_ij,  grad[atom(j)] += f_ij,   f_ij = 2 dPhi/dr^2 (r_i - r_j);
// f_ji = -f_ij, so the total is  grad[a] = -2 sum_{i in a} sum_j f_ij  and pairs on the same atom cancel (skipped
// here).  Thread = point i, the j side staged through shared memory, one reciprocal per pair:
//   f_ij = 3 (rho_i w_i) * [ (2 rho_j w_j) dD / (g_i g_j (g_i + g_j))^2 ] (r_i - r_j),
//   dD   = 2 om_i (g_j s + g_i g_j) + 2 om_j (g_i s + g_i g_j),  s = g_i + g_j.
// "sqrt(r2) > 15" is `r2 > 225` exactly (only r2 == 225 rounds to 15).  The |denom| < 1e-20 guard of the reference
// cannot fire (g >= kappa >= kappa(1e-8) > 0.1 for active points) and is not evaluated.  j-blocks whose bounding box is
// farther than the cut-off from the i-block's are skipped as a whole.
// ------------------------------------------------------------------------------------------------
__global__ void k_vv10_point_atoms(size_t n, int natoms, const size_t *__restrict__ atom_offsets, int *__restrict__ pt_atom) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = natoms; // largest a with atom_offsets[a] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (atom_offsets[mid] <= i) lo = mid;
    else hi = mid;
  }
  pt_atom[i] = lo;
}

// bounding box of the ACTIVE points of each block of 128 (empty box: lo = +inf)
__global__ void __launch_bounds__(TILE_P) k_vv10_bbox(VV10Params V) {
  const size_t i = static_cast<size_t>(blockIdx.x) * TILE_P + threadIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  if (i < V.n && V.kap[i] > 0.0) {
    lo[0] = hi[0] = V.X[i];
    lo[1] = hi[1] = V.Y[i];
    lo[2] = hi[2] = V.Z[i];
  }
  __shared__ double red[6][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) {
      red[d][warp] = lo[d];
      red[3 + d][warp] = hi[d];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    V.bbox[6 * static_cast<size_t>(blockIdx.x) + d] = fmin(fmin(red[d][0], red[d][1]), fmin(red[d][2], red[d][3]));
    V.bbox[6 * static_cast<size_t>(blockIdx.x) + 3 + d] =
        fmax(fmax(red[3 + d][0], red[3 + d][1]), fmax(red[3 + d][2], red[3 + d][3]));
  }
}

// work item = (i-block bi, j-chunk): j-blocks [jc*VV_JCHUNK, min((jc+1)*VV_JCHUNK, nblk)) — the full square
__global__ void __launch_bounds__(TILE_P) k_vv10_grid_pairs(VV10Params V, const int2 *__restrict__ items, int nblk) {
  __shared__ double sx[TILE_P], sy[TILE_P], sz[TILE_P], srw[TILE_P], som[TILE_P], skp[TILE_P];
  __shared__ int sat[TILE_P];
  const int2 it = items[blockIdx.x];
  const int bi = it.x, tid = threadIdx.x;
  const size_t i = static_cast<size_t>(bi) * TILE_P + tid;
  const bool ivalid = i < V.n;
  double xi = 0, yi = 0, zi = 0, omi = 1, kpi = -1;
  int ati = -1;
  if (ivalid) {
    xi = V.X[i];
    yi = V.Y[i];
    zi = V.Z[i];
    omi = V.om[i];
    kpi = V.kap[i];
    ati = V.pt_atom[i];
  }
  const bool active = kpi > 0.0;
  if (!active) kpi = 1.0;
  double bl[3], bh[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    bl[d] = V.bbox[6 * static_cast<size_t>(bi) + d];
    bh[d] = V.bbox[6 * static_cast<size_t>(bi) + 3 + d];
  }
  if (!(bl[0] <= bh[0])) return; // no active point in this block
  double fx = 0.0, fy = 0.0, fz = 0.0;
  const int jb0 = it.y * VV_JCHUNK, jb1 = min(jb0 + VV_JCHUNK, nblk);
  for (int jb = jb0; jb < jb1; ++jb) {
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double jl = V.bbox[6 * static_cast<size_t>(jb) + d], jh = V.bbox[6 * static_cast<size_t>(jb) + 3 + d];
      const double gap = fmax(0.0, fmax(bl[d] - jh, jl - bh[d]));
      d2 += gap * gap; // +inf for an empty j box
    }
    if (!(d2 <= 225.0)) continue; // every pair of the two blocks is beyond the cut-off (uniform over the CTA)
    const size_t j = static_cast<size_t>(jb) * TILE_P + tid;
    __syncthreads();
    if (j < V.n) {
      sx[tid] = V.X[j];
      sy[tid] = V.Y[j];
      sz[tid] = V.Z[j];
      srw[tid] = V.rw2[j];
      som[tid] = V.om[j];
      skp[tid] = fabs(V.kap[j]);
      sat[tid] = V.pt_atom[j];
    } else {
      sx[tid] = sy[tid] = sz[tid] = 0.0;
      srw[tid] = 0.0;
      som[tid] = skp[tid] = 1.0;
      sat[tid] = -2;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < TILE_P; ++jj) {
      const double dx = xi - sx[jj], dy = yi - sy[jj], dz = zi - sz[jj];
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double gi = fma(r2, omi, kpi), gj = fma(r2, som[jj], skp[jj]);
      const double s = gi + gj, gg = gi * gj;
      const double inv = 1.0 / (gg * s);
      const double dD = 2.0 * omi * fma(gj, s, gg) + 2.0 * som[jj] * fma(gi, s, gg);
      double t = srw[jj] * dD * (inv * inv);
      if (r2 > 225.0 || sat[jj] == ati) t = 0.0; // beyond the cut-off / same atom (cancels, also covers j == i)
      fx = fma(t, dx, fx);
      fy = fma(t, dy, fy);
      fz = fma(t, dz, fz);
    }
  }
  if (ivalid && active) {
    const double c = 1.5 * V.rw2[i]; // 3 rho_i w_i
    atomicAdd(V.GX + i, c * fx);
    atomicAdd(V.GY + i, c * fy);
    atomicAdd(V.GZ + i, c * fz);
  }
}

// grad[atom(i)] -= 2 f_i
__global__ void k_vv10_grid_reduce(VV10Params V, double *__restrict__ grad) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= V.n) return;
  const double gx = V.GX[i], gy = V.GY[i], gz = V.GZ[i];
  if (gx == 0.0 && gy == 0.0 && gz == 0.0) return;
  const int a = V.pt_atom[i];
  atomicAdd(grad + 3 * a, -2.0 * gx);
  atomicAdd(grad + 3 * a + 1, -2.0 * gy);
  atomicAdd(grad + 3 * a + 2, -2.0 * gz);
}

void launch_vv10_point_atoms(const VV10Params &V, int natoms, const size_t *atom_offsets_dev, void *stream) {
  k_vv10_point_atoms<<<static_cast<unsigned>((V.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      V.n, natoms, atom_offsets_dev, V.pt_atom);
}
void launch_vv10_grid_pairs(const VV10Params &V, const int2 *items, int nitems, void *stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>((V.n + TILE_P - 1) / TILE_P);
  k_vv10_bbox<<<nblk, TILE_P, 0, st>>>(V);
  if (nitems > 0) k_vv10_grid_pairs<<<nitems, TILE_P, 0, st>>>(V, items, nblk);
}
void launch_vv10_grid_reduce(const VV10Params &V, double *grad, void *stream) {
  k_vv10_grid_reduce<<<static_cast<unsigned>((V.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(V, grad);
}
void launch_vv10_merge(const VV10Peers &src, const VV10Params &out, void *stream) {
  k_vv10_merge<<<static_cast<unsigned>((out.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, out);
}

void launch_vv10_prepare(const BuildParams &P, const VV10Params &V, int ntiles, double unowned, void *stream) {
  k_vv10_init<<<static_cast<unsigned>((V.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(V, unowned);
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
