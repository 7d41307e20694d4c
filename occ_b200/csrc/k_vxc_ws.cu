// k_vxc_ws.cu — warp-specialised persistent form of the Vxc contraction (restricted GGA builds; opt-in, see launch site).
//
// Same arithmetic as k_vxc_resident (k_gemm.cu): Ktmp_loc = phi^T . G with G = c0 phi + cx phi_x + cy phi_y + cz phi_z
// (xc_potential_matrix.cpp:33-39 of the reference), a (tile, 64 output columns) item per step, 16 x 64 warp tiles on
// DMMA.8x8x4, red.global.add.f64 scatter.  What changes is who does the data movement.  In k_vxc_resident every warp
// builds the B block, then multiplies and refills its own A ring, so the tensor pipe idles while a CTA is in its build
// (measured: 4.5 % of the kernel on valinomycin, 13-18 % at nb ~ 150) and the refills take issue slots from the DMMA
// warps.  Here ONE CTA per SM stays resident and walks a strided list of items with three roles:
//   warps 0-15  consumers: LDS + DMMA + scatter only; warp w owns rows [16 (w & 7), +16) of every 128-row block and the
//               32-column half (w >> 3) of the item's 64 columns (16 x 32 warp tiles: 96 registers, so that 16 of them fit
//               next to the movers); the two warps of a row group share one A ring; they wait on mbarriers for their A
//               stages and for the item's B block
//   warps 16-17 A producers: cp.async the 16 x 16 A slices into the 8 rings (warp 16 feeds row groups 0-3, warp 17
//               feeds 4-7), completion signalled through cp.async.mbarrier.arrive
//   warps 18-19 B builders: build the NEXT item's B block into the other of two shared-memory buffers while the
//               consumers multiply the current one
// Shared memory: 2 x 66 KiB (B) + 8 x 3 x 2.5 KiB (A rings) = 192 KiB.
#include "gemm_common.cuh"
#include <cstdint>

namespace xcb {

namespace {
#ifndef WS_RS_
#define WS_RS_ 3
#endif
#ifndef WS_RK_
#define WS_RK_ 16
#endif
constexpr int WS_RS = WS_RS_, WS_RK = WS_RK_;
constexpr int WS_LDA = WS_RK + 4, WS_LDB = TILE_P + 4;
constexpr int WS_ASZ = 16 * WS_LDA, WS_BSZ = GN * WS_LDB;
constexpr int WS_NRG = 8;    // row groups of a 128-row block = A rings
constexpr int WS_NCONS = 16; // consumer warps: row group x column half
constexpr int WS_THREADS = 32 * (WS_NCONS + 4);
constexpr int WS_NKC = TILE_P / WS_RK;
constexpr int WS_SMEM_BYTES =
    (2 * WS_BSZ + WS_NRG * WS_RS * WS_ASZ) * 8 + 2 * GN * 4 + (2 * WS_NRG * WS_RS + 4) * 8;

__device__ __forceinline__ unsigned sm_addr(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(sm_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(sm_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WS_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WS_DONE;\n"
               "bra WS_WAIT;\n"
               "WS_DONE:\n"
               "}\n" ::"r"(sm_addr(b)),
               "r"(parity)
               : "memory");
}
// the executing thread's arrive on `b` happens when all its earlier cp.async have landed (count not incremented: the
// barrier is initialised with one expected arrival per producer lane)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(sm_addr(b)) : "memory");
}
} // namespace

__global__ void __launch_bounds__(WS_THREADS, 1) k_vxc_ws(BuildParams P, const VxcItem *__restrict__ items, int nitems) {
  extern __shared__ __align__(16) double smem[];
  double *Bbuf = smem;
  double *Aring = smem + 2 * WS_BSZ;
  int *scol = reinterpret_cast<int *>(Aring + WS_NRG * WS_RS * WS_ASZ);
  uint64_t *A_full = reinterpret_cast<uint64_t *>(scol + 2 * GN); // [8][RS]
  uint64_t *A_empty = A_full + WS_NRG * WS_RS;                     // [8][RS]
  uint64_t *B_full = A_empty + WS_NRG * WS_RS;                     // [2]
  uint64_t *B_empty = B_full + 2;                                  // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < WS_NRG * WS_RS; ++i) {
      mbar_init(A_full + i, 32); // one deferred arrival per producer lane
      mbar_init(A_empty + i, 2); // lane 0 of the two consumer warps of the row group
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(B_full + i, 2);          // one arrival per builder warp
      mbar_init(B_empty + i, WS_NCONS);  // one per consumer warp
    }
  }
  __syncthreads();

  if (warp < WS_NCONS) {
    // ------------------------------------------------------------------------------------------ consumers
    const int gid = lane >> 2, tig = lane & 3;
    const int rg = warp & 7, ch = warp >> 3; // row group, column half
    unsigned c = 0; // A stages consumed so far (ring position and phase)
    int kk = 0;     // items processed so far (B buffer and phase)
    for (int idx = blockIdx.x; idx < nitems; idx += gridDim.x) {
      const VxcItem it = items[idx];
      const TileDesc t = P.tiles[it.tile];
      if (P.leaf_skip[t.leaf]) continue;
      const int nb = t.nb;
      const int *gidx = P.bf_index + t.bf_off;
      const int n0 = it.jb * GN;
      const int nib = (nb + GM - 1) / GM;
      const int nq = nib * WS_NKC;
      const int ni_valid = min(4, max(0, (nb - (n0 + ch * 32) + 7) >> 3)); // 8-column groups of this warp's half
      const int buf = kk & 1;
      mbar_wait(B_full + buf, (kk >> 1) & 1);
      const double *Bw = Bbuf + buf * WS_BSZ + (ch * 32 + gid) * WS_LDB + tig;
      const int *sc = scol + buf * GN + ch * 32;
      double *Kt = P.Ktmp[it.spin];

      double acc[2][4][2];
      int ib = 0, kc = 0, mi_valid = 0;
      bool full = false;
      int grow[2] = {0, 0};
      for (int q = 0; q < nq; ++q) {
        if (kc == 0) {
          mi_valid = min(2, max(0, (nb - (ib * GM + rg * 16) + 7) >> 3));
          full = mi_valid == 2 && ni_valid == 4;
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const int i = ib * GM + rg * 16 + mi * 8 + gid;
            grow[mi] = i < nb ? __ldg(gidx + i) : -1;
          }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        }
        const unsigned slot = c % WS_RS;
        mbar_wait(A_full + rg * WS_RS + slot, (c / WS_RS) & 1);
        const double *As = Aring + (rg * WS_RS + slot) * WS_ASZ + gid * WS_LDA + tig;
        const double *Bs = Bw + kc * WS_RK;
        if (full) {
#pragma unroll
          for (int k4 = 0; k4 < WS_RK / 4; ++k4) {
            double a[2], b[4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * WS_LDA + k4 * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * WS_LDB + k4 * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
          }
        } else if (mi_valid > 0 && ni_valid > 0) {
#pragma unroll
          for (int k4 = 0; k4 < WS_RK / 4; ++k4) {
            double a[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * WS_LDA + k4 * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
              if (ni < ni_valid) {
                const double b = Bs[ni * 8 * WS_LDB + k4 * 4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
                  if (mi < mi_valid) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b);
              }
          }
        }
        __syncwarp(); // every lane has read its fragments of this stage
        if (lane == 0) mbar_arrive(A_empty + rg * WS_RS + slot);
        ++c;
        if (kc == WS_NKC - 1 && !(P.debug & 1)) {
          // scatter-add this warp's 16 x 32 block of Ktmp_loc into the full matrix (dft.h:361-369)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int2 gc = *reinterpret_cast<const int2 *>(sc + ni * 8 + 2 * tig);
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              if (grow[mi] >= 0) {
                if (gc.x >= 0) red_add_f64(Kt + static_cast<size_t>(gc.x) * P.ldD + grow[mi], acc[mi][ni][0]);
                if (gc.y >= 0) red_add_f64(Kt + static_cast<size_t>(gc.y) * P.ldD + grow[mi], acc[mi][ni][1]);
              }
            }
          }
        }
        if (++kc == WS_NKC) {
          kc = 0;
          ++ib;
        }
      }
      __syncwarp(); // all lanes are done with this item's B block and column map
      if (lane == 0) mbar_arrive(B_empty + buf);
      ++kk;
    }
  } else if (warp < WS_NCONS + 2) { // warps 16, 17
    // ------------------------------------------------------------------------------------------ A producers
    const int cw0 = (warp - WS_NCONS) * 4;
    unsigned c = 0;
    for (int idx = blockIdx.x; idx < nitems; idx += gridDim.x) {
      const VxcItem it = items[idx];
      const TileDesc t = P.tiles[it.tile];
      if (P.leaf_skip[t.leaf]) continue;
      const int nbp = t.nbp;
      const double *Ag = P.scratch + t.scr_off; // phi[i * 128 + p]
      const int nq = ((t.nb + GM - 1) / GM) * WS_NKC;
      int ib = 0, kc = 0;
      for (int q = 0; q < nq; ++q) {
        const unsigned slot = c % WS_RS, par = ((c / WS_RS) & 1) ^ 1;
#pragma unroll
        for (int cw = 0; cw < 4; ++cw) {
          const int w = cw0 + cw;
          mbar_wait(A_empty + w * WS_RS + slot, par);
          double *As = Aring + (w * WS_RS + slot) * WS_ASZ;
#pragma unroll
          constexpr int SPR = WS_RK / 2; // 16-byte segments per row
#pragma unroll
          for (int r = 0; r < 16 * SPR / 32; ++r) {
            const int seg = lane + 32 * r;
            const int row = seg / SPR, s2 = (seg % SPR) * 2;
            const int i = ib * GM + w * 16 + row;
            cp_async16(As + row * WS_LDA + s2, Ag + static_cast<size_t>(i < nbp ? i : 0) * TILE_P + kc * WS_RK + s2, i < nbp);
          }
          cp_async_mbar_arrive(A_full + w * WS_RS + slot);
        }
        ++c;
        if (++kc == WS_NKC) {
          kc = 0;
          ++ib;
        }
      }
    }
    cp_async_wait<0>();
  } else {
    // ------------------------------------------------------------------------------------------ B builders
    const int bt = tid - 32 * (WS_NCONS + 2); // 0..63: points bt and bt + 64
    int kk = 0;
    for (int idx = blockIdx.x; idx < nitems; idx += gridDim.x) {
      const VxcItem it = items[idx];
      const TileDesc t = P.tiles[it.tile];
      if (P.leaf_skip[t.leaf]) continue;
      const int nb = t.nb;
      const int n0 = it.jb * GN;
      const int buf = kk & 1;
      mbar_wait(B_empty + buf, ((kk >> 1) & 1) ^ 1);
      const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
      const double *cf = P.coef + (static_cast<size_t>(it.tile) * P.nspin + it.spin) * NCOEF * TILE_P;
      const double *q0 = P.scratch + t.scr_off + static_cast<size_t>(n0) * TILE_P;
      double *dst = Bbuf + buf * WS_BSZ;
      double co[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int s = 0; s < 4; ++s) co[h][s] = cf[s * TILE_P + bt + 64 * h];
      const int jv = min(GN, max(0, nb - n0));
      const int jv4 = (jv + 3) & ~3; // the slabs are zero-padded to nbp (a multiple of 64)
      for (int j0 = 0; j0 < jv4; j0 += 4) {
        double v[4][2][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double *q = q0 + (j0 + jj) * TILE_P + bt + 64 * h;
            v[jj][h][0] = q[0];
            v[jj][h][1] = q[slab];
            v[jj][h][2] = q[2 * slab];
            v[jj][h][3] = q[3 * slab];
          }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            dst[(j0 + jj) * WS_LDB + bt + 64 * h] =
                co[h][0] * v[jj][h][0] + co[h][1] * v[jj][h][1] + co[h][2] * v[jj][h][2] + co[h][3] * v[jj][h][3];
      }
      for (int j = jv4; j < GN; ++j) {
        dst[j * WS_LDB + bt] = 0.0;
        dst[j * WS_LDB + bt + 64] = 0.0;
      }
      {
        const int cidx = n0 + bt;
        scol[buf * GN + bt] = cidx < nb ? __ldg(P.bf_index + t.bf_off + cidx) : -1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(B_full + buf);
      ++kk;
    }
  }
}

// items: (tile, jb) with acomp 0, one spin per item (not the spin-fused form), P.deriv >= 1
void launch_vxc_ws(const BuildParams &P, const VxcItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  static SmemOptIn optin;
  optin.ensure(k_vxc_ws, WS_SMEM_BYTES);
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = nitems < nsm ? nitems : nsm;
  k_vxc_ws<<<grid, WS_THREADS, WS_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(P, items, nitems);
}

} // namespace xcb
