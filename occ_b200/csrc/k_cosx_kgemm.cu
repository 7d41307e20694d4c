// k_cosx_kgemm.cu — K' += Phi_t^T . G_t of the seminumerical exchange sweep (seminumerical_exchange.cpp:797,
// `K.noalias() -= phi^T * Gg` per batch) on the FP64 tensor pipe, one launch per chunk of batches.
//
// Same resident-B design as k_vxc_resident (k_gemm.cu): a CTA owns (batch, 64 output columns); its B operand, the
// 128-point x 64-column block of the batch's G slab, is copied ONCE into shared memory, the A operand (Phi_t^T, all
// nbf rows) streams per warp through a private cp.async ring (warp w owns rows [16 w, 16 w + 16) of every 128-row
// block), and the 16 x 64 warp tiles are added into K' with red.global.add.f64. The product has k = 128 only, so a
// library GEMM per batch re-reads and re-writes the whole K' (90 MB for valinomycin) 2652 times and spends its time
// in pipeline fill; here the pipeline runs through all row blocks of a CTA and K' is touched by reductions only.
//
// G is sparse by construction: a column of G_t is written only when a shell pair containing that shell passed the
// geometric and the density screen for batch t (:746-764). A CTA whose B block is identically zero contributes
// exactly nothing and exits after the B copy — the result equals the dense product bit for bit.
#include "cosx_internal.h"
#include "gemm_common.cuh"

namespace xcb {

namespace {
constexpr int RK = 16, RS = 2;
constexpr int R_LDA = RK + 4, R_LDB = TILE_P + 4; // = 4 (mod 16) doubles: conflict-free f64 MMA fragment loads
constexpr int R_WROWS = GM / 8;
constexpr int R_ASZ = R_WROWS * R_LDA, R_BSZ = GN * R_LDB;
constexpr int R_SMEM_DOUBLES = R_BSZ + 8 * RS * R_ASZ;
} // namespace

__global__ void __launch_bounds__(NTHREADS, 2)
    k_cosx_kgemm(const double *__restrict__ phi, const double *__restrict__ G, const int *__restrict__ tile_skip, int nbf,
                 int nbp, int njb, double *__restrict__ K, unsigned long long *__restrict__ stats) {
  extern __shared__ __align__(16) double smem[];
  const int tile = blockIdx.x / njb, jb = blockIdx.x - tile * njb;
  if (tile_skip[tile]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const size_t slab = static_cast<size_t>(nbp) * TILE_P;
  const double *Ag = phi + tile * slab; // Phi_t[i*128 + p]
  const double *Bg = G + tile * slab;   // G_t[j*128 + p]
  const int n0 = jb * GN;
  const int nib = (nbf + GM - 1) / GM;
  constexpr int NKC = TILE_P / RK;
  const int nq = nib * NKC; // flattened (row block, k chunk)
  const int ni_valid = min(8, max(0, (nbf - n0 + 7) >> 3));
  double *Bres = smem;                             // [64][R_LDB]
  double *Aw = smem + R_BSZ + warp * (RS * R_ASZ); // this warp's ring: RS x [16][R_LDA]

  // B block: 64 columns x 64 16-byte segments
  bool nz = false;
  {
    constexpr int SEGS = GN * (TILE_P / 2);
#pragma unroll 4
    for (int s = tid; s < SEGS; s += NTHREADS) {
      const int j = s / (TILE_P / 2), p2 = (s % (TILE_P / 2)) * 2;
      const bool ok = n0 + j < nbf;
      cp_async16(Bres + j * R_LDB + p2, Bg + static_cast<size_t>(ok ? n0 + j : 0) * TILE_P + p2, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
#pragma unroll 4
    for (int s = tid; s < SEGS; s += NTHREADS) {
      const int j = s / (TILE_P / 2), p2 = (s % (TILE_P / 2)) * 2;
      const double2 v = *reinterpret_cast<const double2 *>(Bres + j * R_LDB + p2);
      nz |= (v.x != 0.0) | (v.y != 0.0);
    }
  }
  if (!__syncthreads_or(nz)) return; // also the only CTA-wide barrier: B block complete
  if (stats && tid == 0) atomicAdd(stats, 1ull);

  auto load_stage = [&](int q, int stage) {
    double *As = Aw + stage * R_ASZ;
    const int ib = q / NKC, kc = q % NKC;
    constexpr int SPR = RK / 2; // 16-byte segments per row
#pragma unroll
    for (int r = 0; r < R_WROWS * SPR / 32; ++r) {
      const int seg = lane + 32 * r;
      const int row = seg / SPR, s2 = (seg % SPR) * 2;
      const int i = ib * GM + warp * R_WROWS + row;
      cp_async16(As + row * R_LDA + s2, Ag + static_cast<size_t>(i < nbf ? i : 0) * TILE_P + kc * RK + s2, i < nbf);
    }
  };
#pragma unroll
  for (int s = 0; s < RS - 1; ++s) {
    if (s < nq) load_stage(s, s);
    cp_async_commit();
  }

  double acc[2][8][2];
  int ib = 0, kc = 0;
  bool full = false;
  int mi_valid = 0;
  const double *Bw = Bres + gid * R_LDB + tig;
  for (int q = 0; q < nq; ++q) {
    if (kc == 0) {
      mi_valid = min(2, max(0, (nbf - (ib * GM + warp * R_WROWS) + 7) >> 3));
      full = mi_valid == 2 && ni_valid == 8;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
    cp_async_wait<RS - 2>();
    __syncwarp(); // this warp's stage q has landed for all its lanes; its stage q-1 is free again
    const double *As = Aw + (q % RS) * R_ASZ + gid * R_LDA + tig;
    const double *Bs = Bw + kc * RK;
    if (full) {
#pragma unroll
      for (int k4 = 0; k4 < RK / 4; ++k4) {
        double a[2], b[8];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * R_LDA + k4 * 4];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) b[ni] = Bs[ni * 8 * R_LDB + k4 * 4];
        if (k4 == 0) { // refill as soon as the first fragments are in flight
          if (q + RS - 1 < nq) load_stage(q + RS - 1, (q + RS - 1) % RS);
          cp_async_commit();
        }
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    } else {
      if (q + RS - 1 < nq) load_stage(q + RS - 1, (q + RS - 1) % RS);
      cp_async_commit();
      if (mi_valid > 0 && ni_valid > 0) {
#pragma unroll
        for (int k4 = 0; k4 < RK / 4; ++k4) {
          double a[2];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * R_LDA + k4 * 4];
#pragma unroll
          for (int ni = 0; ni < 8; ++ni)
            if (ni < ni_valid) {
              const double b = Bs[ni * 8 * R_LDB + k4 * 4];
#pragma unroll
              for (int mi = 0; mi < 2; ++mi)
                if (mi < mi_valid) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b);
            }
        }
      }
    }
    if (kc == NKC - 1) {
      // accumulator layout of the f64 MMA: row = gid, columns 2 tig and 2 tig + 1 of each 8 x 8 tile
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int row = ib * GM + warp * R_WROWS + mi * 8 + gid;
        if (row < nbf) {
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) {
            const int col = n0 + ni * 8 + 2 * tig;
            if (col < nbf) red_add_f64(K + static_cast<size_t>(col) * nbf + row, acc[mi][ni][0]);
            if (col + 1 < nbf) red_add_f64(K + static_cast<size_t>(col + 1) * nbf + row, acc[mi][ni][1]);
          }
        }
      }
    }
    if (++kc == NKC) {
      kc = 0;
      ++ib;
    }
  }
  cp_async_wait<0>();
}

void launch_cosx_kgemm(const double *phi, const double *G, const int *tile_skip, int ntiles, int nbf, int nbp, double *K,
                       unsigned long long *stats, void *stream) {
  if (ntiles <= 0) return;
  const int smem_bytes = R_SMEM_DOUBLES * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(k_cosx_kgemm, smem_bytes);
  const int njb = (nbf + GN - 1) / GN;
  k_cosx_kgemm<<<ntiles * njb, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(phi, G, tile_skip, nbf, nbp, njb,
                                                                                          K, stats);
}

} // namespace xcb
