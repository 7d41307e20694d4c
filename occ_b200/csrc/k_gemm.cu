// k_gemm.cu — the Vxc contraction (k_vxc_resident) and small matrix helpers.  The density contraction described
// below lives in k_density.cu (D form) and k_mo.cu (factorised form); the overview is kept here.
//
//   density  (reference: occ::density::evaluate_density, src/gto/density.cpp:8-166)
//       X = phi_a . D_loc  (128 points x nb, K = nb), immediately row-dotted with the AO slabs:
//       rho = rowsum(phi o X), d_k rho = 2 rowsum(phi_k o X), tau = 1/2 sum_k rowsum(phi_k o (phi_k D))
//       D_loc is never materialised: the B operand is gathered from the full density matrix through the
//       leaf's local->global index list while it is staged into shared memory (reference gathers Dloc
//       on the host, include/occ/dft/dft.h:340-348).
//   vxc      (reference: xc_potential_matrix, src/dft/xc_potential_matrix.cpp:8-135)
//       Ktmp_loc = phi_a^T . B  (nb x nb, K = 128 points), B = G (the weighted AO combination) or c_tau phi_k,
//       built per CTA in shared memory; scattered with red.global.add.f64 straight into the full-size
//       matrix (reference scatters Kblk on the host, dft.h:361-369).  K = Ktmp + Ktmp^T afterwards.
//
// Both run on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — the only f64 MMA shape sm_100a
// has in hardware; tcgen05/UMMA has no f64 kind).  CTA tile 128 x 64, K chunks of 16, 8 warps in a 4 x 2
// arrangement with 32 x 32 warp tiles (16 DMMA per 8 conflict-free 64-bit LDS), a 3-stage cp.async pipeline,
// 2 CTAs per SM.  Shared-memory leading dimensions are = 4 (mod 16) doubles, which makes every fragment
// load bank-conflict free for the (groupID, threadID_in_group) access pattern of the f64 MMA.
#include "gemm_common.cuh"

namespace xcb {

// ------------------------------------------------------------------------------------------------
// G slab: G_s = c0 phi + cx phi_x + cy phi_y + cz phi_z  (the bracket of xc_potential_matrix.cpp:33-39 / :86-99),
// materialised in HBM only for the XC nuclear gradient (A operand of k_xc_gradient); the Vxc build forms its G block
// in shared memory inside k_vxc_resident.
// ------------------------------------------------------------------------------------------------
#ifndef XC_DETERMINISTIC
__global__ void __launch_bounds__(TILE_P) k_gbuild(BuildParams P, const DensItem *__restrict__ items) {
  const DensItem it = items[blockIdx.x];
  const TileDesc t = P.tiles[it.tile];
  if (P.leaf_skip[t.leaf]) return;
  const int p = threadIdx.x;
  const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
  const double *phi = P.scratch + t.scr_off + static_cast<size_t>(it.jb) * GN * TILE_P + p;
  for (int s = 0; s < P.nspin; ++s) {
    const double *cf = P.coef + (static_cast<size_t>(it.tile) * P.nspin + s) * NCOEF * TILE_P + p;
    const double c0 = cf[0];
    double *G = const_cast<double *>(phi) + static_cast<size_t>(P.ncomp + s) * slab;
    if (P.deriv >= 1) {
      const double cx = cf[TILE_P], cy = cf[2 * TILE_P], cz = cf[3 * TILE_P];
#pragma unroll 4
      for (int j = 0; j < GN; ++j) {
        const double *q = phi + j * TILE_P;
        G[j * TILE_P] = c0 * q[0] + cx * q[slab] + cy * q[2 * slab] + cz * q[3 * slab];
      }
    } else {
#pragma unroll 4
      for (int j = 0; j < GN; ++j) G[j * TILE_P] = c0 * phi[j * TILE_P];
    }
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// vxc, resident-B form (the one the build launches).  A CTA owns (tile, 64 output columns): its B operand — the
// 128-point x 64-column block of G_s = c0 phi + cx phi_x + cy phi_y + cz phi_z (xc_potential_matrix.cpp:33-39 /
// :86-99), or c_tau phi_k for the meta-GGA terms (:52-56) — is the same for every row block, so it is BUILT ONCE IN
// SHARED MEMORY straight from the AO slabs (no G slab in HBM, no k_gbuild pass, no per-chunk B traffic, no per-point
// DMUL in the MMA loop).  Only the A operand (phi_a^T) streams, and it streams PER WARP: warp w owns rows
// [16 w, 16 w + 16) of every 128-row block and all 64 columns (2 x 8 DMMA tiles), stages its own 16 x RK slices
// through a private cp.async ring and synchronises with __syncwarp only — the main loop has no CTA-wide barrier.
// Shared memory: B 64 x 132 doubles (66 KiB) + 8 warps x RS stages x 16 x 20 doubles (40 KiB) -> 2 CTAs / SM.
#ifndef XC_RK
#define XC_RK 16
#endif
#ifndef XC_RS
#define XC_RS 2
#endif
constexpr int RK = XC_RK, RS = XC_RS;
#ifndef XC_BPF_AHEAD
#define XC_BPF_AHEAD 64
#endif
#ifndef XC_SWZ
#define XC_SWZ 0
#endif
// XC_SWZ = 0: padded rows (ld = 4 mod 16).  XC_SWZ = 1: dense rows, element (row, k) stored at k ^ ((row & 3) << 2): the
// same conflict-free fragment pattern without padding, which makes room for a third stage of the A rings.
constexpr int R_LDA = XC_SWZ ? RK : RK + 4;
constexpr int R_LDB = XC_SWZ ? TILE_P : TILE_P + 4;
#define R_SW(row) (XC_SWZ ? (((row) & 3) << 2) : 0)
constexpr int R_WROWS = GM / 8;        // rows of a row block owned by one warp
constexpr int R_ASZ = R_WROWS * R_LDA, R_BSZ = GN * R_LDB;
constexpr int R_SMEM_DOUBLES = R_BSZ + 8 * RS * R_ASZ + GN / 2; // + 64 ints: global column of each output column

__global__ void __launch_bounds__(NTHREADS, 2) XC_SYM(k_vxc_resident)(BuildParams P, const VxcItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  const VxcItem it = items[blockIdx.x];
  const TileDesc t = P.tiles[it.tile];
  if (P.leaf_skip[t.leaf]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nb = t.nb, nbp = t.nbp;
  const size_t slab = static_cast<size_t>(nbp) * TILE_P;
  const double *base = P.scratch + t.scr_off;
  const double *Ag = base + static_cast<size_t>(it.acomp) * slab; // phi_a[i*128 + p]
  const int *gidx = P.bf_index + t.bf_off;
  // it.spin == 2 (unrestricted builds): the CTA owns 32 basis-function columns for BOTH spins — B columns 0..31 are the
  // alpha block, 32..63 the beta block of the same functions — so that the four AO slabs of those columns are read once
  // for the two spins (the B-block build is a quarter of the kernel's warp time at nb ~ 200)
  const bool both = it.spin == 2;
  const int ncol = both ? GN / 2 : GN;
  const int n0 = it.jb * ncol;
  // Symmetric terms — phi_k^T diag(c_tau) phi_k of a meta-GGA (acomp != 0) and phi^T diag(c0) phi of an LDA build —
  // need only their upper triangle, because the build returns K = Ktmp + Ktmp^T: row blocks strictly below the CTA's
  // 64 columns are skipped, blocks strictly above count twice, blocks that straddle the diagonal are computed whole.
  const bool symmetric = it.acomp != 0 || P.deriv == 0;
  const int nib_all = (nb + GM - 1) / GM;
  const int nib = symmetric ? min(nib_all, (n0 + ncol + GM - 1) / GM) : nib_all;
  constexpr int NKC = TILE_P / RK;
  const int nq = nib * NKC; // flattened (ib, kc) chunks
  const int ni_valid = min(ncol / 8, max(0, (nb - n0 + 7) >> 3)); // valid 8-column groups (per spin when `both`)
  double *Bres = smem;                                   // [64][R_LDB]
  double *Aw = smem + R_BSZ + warp * (RS * R_ASZ);       // this warp's ring: RS x [16][R_LDA]
  int *scol = reinterpret_cast<int *>(smem + R_BSZ + 8 * RS * R_ASZ);

  auto load_stage = [&](int q, int stage) {
    double *As = Aw + stage * R_ASZ;
    const int ib = q / NKC, kc = q % NKC;
    constexpr int SPR = RK / 2; // 16-byte segments per row
#pragma unroll
    for (int r = 0; r < R_WROWS * SPR / 32; ++r) {
      const int seg = lane + 32 * r;
      const int row = seg / SPR, s2 = (seg % SPR) * 2;
      const int i = ib * GM + warp * R_WROWS + row;
      cp_async16(As + row * R_LDA + (s2 ^ R_SW(row)), Ag + static_cast<size_t>(i < nbp ? i : 0) * TILE_P + kc * RK + s2,
                 i < nbp);
    }
  };
  // the A rings start filling while the B block is being built
#pragma unroll
  for (int s = 0; s < RS - 1; ++s) {
    if (s < nq) load_stage(s, s);
    cp_async_commit();
  }

  { // resident B block: thread = (point, 32-column half)
    const int p = tid & (TILE_P - 1), half = tid >> 7;
    const int sp = both ? half : it.spin, cbase = both ? n0 : n0 + half * 32;
    const double *cf = P.coef + (static_cast<size_t>(it.tile) * P.nspin + sp) * NCOEF * TILE_P + p;
    const double *q0 = base + static_cast<size_t>(cbase) * TILE_P + p;
    double *dst = Bres + (half * 32) * R_LDB;
    if (P.debug & 16) { // timing experiment: B block without the slab loads
      for (int j = 0; j < 32; ++j) dst[j * R_LDB + (p ^ R_SW(j))] = 1e-3 * j;
    } else {
      // only the columns that exist are read (groups of 8; the rest of the block is zero): for local bases of 100-250
      // functions the last 64-column block is mostly padding
      const int jv = min(32, max(0, nb - cbase));
      if (it.acomp == 0) {
        const double c0 = cf[0];
        if (P.deriv >= 1) {
          const double cx = cf[TILE_P], cy = cf[2 * TILE_P], cz = cf[3 * TILE_P];
          for (int j0 = 0; j0 < jv; j0 += 8) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int j = j0 + jj;
              const double *q = q0 + j * TILE_P;
              dst[j * R_LDB + (p ^ R_SW(j))] = c0 * q[0] + cx * q[slab] + cy * q[2 * slab] + cz * q[3 * slab];
            }
          }
        } else {
          for (int j0 = 0; j0 < jv; j0 += 8) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) dst[(j0 + jj) * R_LDB + (p ^ R_SW(j0 + jj))] = c0 * q0[(j0 + jj) * TILE_P];
          }
        }
      } else {
        const double ct = cf[4 * TILE_P];
        const double *qk = q0 + static_cast<size_t>(it.acomp) * slab;
        for (int j0 = 0; j0 < jv; j0 += 8) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) dst[(j0 + jj) * R_LDB + (p ^ R_SW(j0 + jj))] = ct * qk[(j0 + jj) * TILE_P];
        }
      }
      for (int j = (jv + 7) & ~7; j < 32; ++j) dst[j * R_LDB + (p ^ R_SW(j))] = 0.0;
    }
    if (tid < GN) {
      const int c = both ? n0 + (tid & 31) : n0 + tid;
      scol[tid] = c < nb ? __ldg(gidx + c) : -1;
    }
  }
  __syncthreads(); // the only CTA-wide barrier: B block and column map complete

#ifndef XC_NO_BPREFETCH
  // The B-block build is the one phase of this kernel that waits on HBM with the tensor pipe idle.  While this CTA
  // multiplies, pull the slab columns of the item that starts XC_BPF_AHEAD launches later into L2, so that CTA's build
  // pays L2 latency: 4 slabs x 64 (32) columns x 8 lines of 128 bytes.
  const int ahead = (P.debug >> 8) ? static_cast<int>(P.debug >> 8) : XC_BPF_AHEAD;
  if (blockIdx.x + ahead < gridDim.x && !(P.debug & 32)) {
    const VxcItem nx = items[blockIdx.x + ahead];
    const TileDesc tn = P.tiles[nx.tile];
    const int ncn = nx.spin == 2 ? GN / 2 : GN;
    const int c0n = nx.jb * ncn;
    const int cols = min(ncn, max(0, tn.nb - c0n));
    const size_t slabn = static_cast<size_t>(tn.nbp) * TILE_P;
    const double *bn = P.scratch + tn.scr_off + static_cast<size_t>(c0n) * TILE_P;
    const int nslab = nx.acomp == 0 ? (P.deriv >= 1 ? 4 : 1) : 1;
    const double *b0 = nx.acomp == 0 ? bn : bn + static_cast<size_t>(nx.acomp) * slabn;
    const int nlines = nslab * cols * 8;
    for (int L = tid; L < nlines; L += NTHREADS) {
      const int sl = L / (cols * 8), r = L - sl * (cols * 8);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + sl * slabn + static_cast<size_t>(r) * 16));
    }
  }
#endif

  double acc[2][8][2];
  int ib = 0, kc = 0;
  bool full = false;
  int mi_valid = 0;
  int grow[2] = {0, 0};
  const double *Bw = Bres + gid * R_LDB + tig;
  const int g3 = gid & 3; // every row / column index this thread touches is = gid (mod 8): one swizzle term
  for (int q = 0; q < nq; ++q) {
    if (kc == 0) {
      mi_valid = min(2, max(0, (nb - (ib * GM + warp * R_WROWS) + 7) >> 3));
      full = mi_valid == 2 && ni_valid == ncol / 8;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int i = ib * GM + warp * R_WROWS + mi * 8 + gid;
        grow[mi] = i < nb ? __ldg(gidx + i) : -1;
      }
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
        const int ko = XC_SWZ ? ((k4 ^ g3) << 2) : k4 * 4;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * R_LDA + ko];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) b[ni] = Bs[ni * 8 * R_LDB + ko];
        if (k4 == 0) { // refill as soon as the first fragments are in flight
          if (q + RS - 1 < nq && !(P.debug & 8)) load_stage(q + RS - 1, (q + RS - 1) % RS);
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
        const int nmask = both ? 3 : 7; // position of an 8-column group inside its spin block
#pragma unroll
        for (int k4 = 0; k4 < RK / 4; ++k4) {
          double a[2];
          const int ko = XC_SWZ ? ((k4 ^ g3) << 2) : k4 * 4;
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * R_LDA + ko];
#pragma unroll
          for (int ni = 0; ni < 8; ++ni)
            if ((ni & nmask) < ni_valid) {
              const double b = Bs[ni * 8 * R_LDB + ko];
#pragma unroll
              for (int mi = 0; mi < 2; ++mi)
                if (mi < mi_valid) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b);
            }
        }
      }
    }
    if (kc == NKC - 1 && !(P.debug & 1)) {
      // scatter-add this warp's 16 x 64 block of Ktmp_loc into the full matrix (dft.h:361-369)
      if (symmetric && (ib + 1) * GM <= n0) { // strictly above the diagonal: the block stands for both halves
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            acc[mi][ni][0] *= 2.0;
            acc[mi][ni][1] *= 2.0;
          }
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        double *Kt = P.Ktmp[both ? (ni >> 2) : it.spin];
        const int2 gc = *reinterpret_cast<const int2 *>(scol + ni * 8 + 2 * tig);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          if (grow[mi] >= 0) {
            if (gc.x >= 0) XC_K_ADD(P, Kt + static_cast<size_t>(gc.x) * P.ldD + grow[mi], acc[mi][ni][0]);
            if (gc.y >= 0) XC_K_ADD(P, Kt + static_cast<size_t>(gc.y) * P.ldD + grow[mi], acc[mi][ni][1]);
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

void XC_SYM(launch_vxc_resident)(const BuildParams &P, const VxcItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = R_SMEM_DOUBLES * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(XC_SYM(k_vxc_resident), smem_bytes);
  XC_SYM(k_vxc_resident)<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}

#ifndef XC_DETERMINISTIC





__global__ void k_transpose(const double *__restrict__ D, double *__restrict__ DT, int n, int ld) {
  __shared__ double tl[32][33];
  const int bi = blockIdx.x * 32, bj = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + tx, j = bj + r;
    tl[r][tx] = (i < n && j < n) ? D[static_cast<size_t>(j) * ld + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bj + tx, j = bi + r; // DT(i, j) = D(j, i)
    if (i < n && j < n) DT[static_cast<size_t>(j) * ld + i] = tl[tx][r];
  }
}
void launch_transpose(const double *D, double *DT, int nbf, int ld, void *stream) {
  dim3 grid((nbf + 31) / 32, (nbf + 31) / 32);
  k_transpose<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(D, DT, nbf, ld);
}

void launch_gbuild(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  k_gbuild<<<nitems, TILE_P, 0, static_cast<cudaStream_t>(stream)>>>(P, items);
}
#endif

} // namespace xcb
