// k_mo.cu — the factorised ("MO") forms of the density and XC-gradient contractions (occ_xc_build_mo, occ_xc_gradient_mo).
#include "gemm_common.cuh"

namespace xcb {

// The factorised forms below run on a COPY of gemm_xd_body (k_density.cu) rather than on more MODE branches of it:
// folding them into gemm_xd_body changed nothing in its SASS instruction mix but cost the tuned D-form density kernel
// 2.5 % (101.4 -> 103.9 ms on valinomycin) through a different instruction order.
// MODE 2 / 3: the factorised ("MO") form of the density for callers that also hand over the occupied orbitals
// (D2_s = C_s C_s^T; the reference's seam receives `mo`, which carries both):
//   MODE 2   psi_a = phi_a . C_loc   (128 x nocc, K = nb; B gathered along k through the leaf's index list)
//            a = 0: stored to the tile's psi slab;  a = 1..3 (meta-GGA): epilogue sum_i psi_a,i^2 = the tau row sums
//   MODE 3   X = psi . C_loc^T       (128 x nb, K = nocc; B gathered along n), then the row-dot epilogue of MODE 0
// 4 n nb nocc flops instead of 2 n nb^2 for rho / grad rho, 6 n nb nocc instead of 6 n nb^2 for tau; the host picks
// the form per leaf and spin (nb > 2 nocc, resp. nb > nocc).  Same numbers as MODE 0 up to the rounding of C C^T.
// MODE 4 / 5: the same factorisation for the XC nuclear gradient (acc = A . D_s^T = (A C)(C^T) for the symmetric D2 = C C^T):
//   MODE 4   psi_A = A . C_loc with A = G_s or phi_e (stored to the tile's psi slab; A = phi reuses the density's psi)
//   MODE 5   X = psi_A . C_loc^T followed by the per-atom gradient epilogue of MODE 1
template <int MODE> __device__ __forceinline__ void XC_SYM(gemm_mo_body)(const BuildParams &P, const DensItem it, double *smem) {
  const TileDesc t = P.tiles[it.tile];
  if ((MODE == 1 || MODE == 4 || MODE == 5) && P.leaf_skip[t.leaf]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  constexpr bool TO_PSI = MODE == 2 || MODE == 4;   // output columns are orbitals, B = C
  constexpr bool FROM_PSI = MODE == 3 || MODE == 5; // A = psi slab, contraction over orbitals, B = C^T
  const int nbK = FROM_PSI ? P.mo_nocc[it.spin] : t.nb;   // valid extent of the contraction index
  const int nbKp = FROM_PSI ? P.mo_noccp : t.nbp;         // columns the A slab really has
  const int nb = TO_PSI ? P.mo_nocc[it.spin] : t.nb;      // valid extent of the output columns
  const int nk = (nbK + GK - 1) / GK;
  const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
  const size_t psi_slab = static_cast<size_t>(P.mo_noccp) * TILE_P;
  const int a_slab = (MODE == 0 || MODE == 2)
                         ? it.acomp
                         : (it.acomp == 0 ? 0 : (it.acomp == 1 ? P.slab_G + it.spin : it.acomp - 1));
  const long long psi_off = (TO_PSI || FROM_PSI) ? P.tile_psi_off[it.tile] : 0;
  const double *Ag = FROM_PSI ? P.scratch + psi_off + static_cast<size_t>(it.spin) * psi_slab
                              : P.scratch + t.scr_off + static_cast<size_t>(a_slab) * slab;
  const double *Dm = MODE == 0 ? (it.acomp == 0 ? P.Drho : P.Dtau)[it.spin]
                               : (MODE == 1 ? P.DT[it.spin] : (TO_PSI ? P.moC[it.spin] : P.moCT[it.spin]));
  const int ldB = TO_PSI ? P.ldC : (FROM_PSI ? P.ldCT : P.ldD);
  const int *gidx = P.bf_index + t.bf_off;
  const int n0 = it.jb * GN;

  // B loader assignment: this thread stages k = kk of every chunk for NBR columns
  constexpr int NROWT = NTHREADS / GK;   // columns covered per pass
  constexpr int NBR = GN / NROWT;        // passes
  const int kk = tid % GK, nn = tid / GK;
  size_t coloff[NBR];
  bool nvalid[NBR];
#pragma unroll
  for (int r = 0; r < NBR; ++r) {
    const int n = n0 + nn + NROWT * r;
    nvalid[r] = n < nb;
    coloff[r] = nvalid[r] ? static_cast<size_t>(TO_PSI ? n : __ldg(gidx + n)) * ldB : 0;
  }
  // 8-column groups of this warp's 32 columns that hold real basis functions (the rest is padding: no DMMA)
  const int ni_valid = min(4, max(0, (nb - (n0 + wn * 32) + 7) >> 3));

  int grow_next = 0;
  auto load_stage = [&](int kc, int stage) {
    double *As = smem + stage * D_STAGE;
    double *Bs = As + D_ASZ;
#pragma unroll
    for (int r = 0; r < GK * 64 / NTHREADS; ++r) {
      const int seg = tid + NTHREADS * r;
      const int col = seg >> 6, s2 = (seg & 63) * 2;
      cp_async16(As + col * D_LDA + s2, Ag + static_cast<size_t>(kc * GK + col) * TILE_P + s2, kc * GK + col < nbKp);
    }
    // grow_next = global row of k = kc*GK + kk, fetched one iteration ahead so the gather never waits on it
    const bool kvalid = kc * GK + kk < nbK;
    const int grow = grow_next;
#pragma unroll
    for (int r = 0; r < NBR; ++r)
      cp_async8(Bs + (nn + NROWT * r) * D_LDB + kk, Dm + coloff[r] + grow, kvalid && nvalid[r]);
  };
  auto fetch_row = [&](int kc) {
    const int k = kc * GK + kk;
    grow_next = k < nbK ? (FROM_PSI ? k : __ldg(gidx + k)) : 0;
  };

  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    fetch_row(s);
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  fetch_row(GSTAGES - 1);
  // interior warps (all 32 columns real) run the branch-free 16-DMMA step; edge warps skip padded 8-column groups
  const bool full = ni_valid == 4;
  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    const double *As = smem + (kc % GSTAGES) * D_STAGE + wm * 32 + gid;
    const double *Bs = smem + (kc % GSTAGES) * D_STAGE + D_ASZ + (wn * 32 + gid) * D_LDB;
    if (full) {
      // the refill of the stage consumed last iteration is issued in the MIDDLE of this chunk's DMMAs, so that the
      // tensor pipe is fed immediately after the barrier instead of waiting behind 8 cp.async per thread
#pragma unroll
      for (int k4 = 0; k4 < GK / 4; ++k4) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = As[(k4 * 4 + tig) * D_LDA + mi * 8];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * D_LDB + k4 * 4 + tig];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        if (k4 == GK / 8 - 1) {
          if (kc + GSTAGES - 1 < nk) load_stage(kc + GSTAGES - 1, (kc + GSTAGES - 1) % GSTAGES);
          cp_async_commit();
          fetch_row(kc + GSTAGES);
        }
      }
    } else {
      if (kc + GSTAGES - 1 < nk) load_stage(kc + GSTAGES - 1, (kc + GSTAGES - 1) % GSTAGES);
      cp_async_commit();
      fetch_row(kc + GSTAGES);
    }
    if (!full && ni_valid > 0) {
#pragma unroll
      for (int k4 = 0; k4 < GK / 4; ++k4) {
        double a[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = As[(k4 * 4 + tig) * D_LDA + mi * 8];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          if (ni < ni_valid) {
            const double b = Bs[ni * 8 * D_LDB + k4 * 4 + tig];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b);
          }
      }
    }
  }
  cp_async_wait<0>();

  if (TO_PSI) {
    if (MODE == 4 || it.acomp == 0) { // psi tile -> the tile's psi slab (padding columns carry exact zeros)
      double *ps = P.scratch + psi_off + static_cast<size_t>(it.spin) * psi_slab +
                   static_cast<size_t>(n0 + wn * 32 + 2 * tig) * TILE_P + wm * 32 + gid;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          ps[(ni * 8) * TILE_P + mi * 8] = acc[mi][ni][0];
          ps[(ni * 8 + 1) * TILE_P + mi * 8] = acc[mi][ni][1];
        }
    } else { // tau: sum_k sum_i (phi_k C)_i^2 accumulates in row 4 (the functional kernel applies the 1/2)
      double *accp = P.acc + (static_cast<size_t>(it.tile) * P.nspin + it.spin) * NACC * TILE_P + 4 * TILE_P;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        double sq = 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) sq += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][1] * acc[mi][ni][1];
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        if (tig == 0) XC_ACC_ADD(P, accp + wm * 32 + mi * 8 + gid, sq);
      }
    }
    return;
  }
  if (MODE == 1 || MODE == 5) {
    // gradient epilogue
    const double *cbase = P.scratch + t.scr_off + static_cast<size_t>(n0 + wn * 32 + 2 * tig) * TILE_P + wm * 32 + gid;
    const int *batom = P.bf_atom + t.bf_off + n0 + wn * 32 + 2 * tig;
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
      int cslab;
      if (it.acomp == 0) cslab = P.slab_H + 3 * it.spin + d;
      else if (it.acomp == 1) cslab = 1 + d;
      else {
        const int e = it.acomp - 2, lo = d < e ? d : e, hi = d < e ? e : d;
        cslab = P.slab_S + 6 * it.spin + (lo == 0 ? hi : (lo == 1 ? 2 + hi : 5)); // xx xy xz yy yz zz
      }
      const double *pc = cbase + static_cast<size_t>(cslab) * slab;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          s0 += pc[(ni * 8) * TILE_P + mi * 8] * acc[mi][ni][0];
          s1 += pc[(ni * 8 + 1) * TILE_P + mi * 8] * acc[mi][ni][1];
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (gid == 0) {
          const int col = n0 + wn * 32 + ni * 8 + 2 * tig;
          if (col < nb) red_add_f64(P.grad + 3 * __ldg(batom + ni * 8) + d, P.grad_factor * s0);
          if (col + 1 < nb) red_add_f64(P.grad + 3 * __ldg(batom + ni * 8 + 1) + d, P.grad_factor * s1);
        }
      }
    }
    return;
  }
  // epilogue: row-dot the X tile with the AO slabs of the same 64 columns.  The tile leaves the registers through shared
  // memory (the pipeline stages are free now), so that thread = (point p, 32-column half): its slab loads are coalesced
  // along p (one 1 KiB line of a column per 128 threads), independent of each other (deep batches in flight instead of
  // one 8-load round trip per (slab, mi) with the accumulator fragment live) and need no shuffles; one red.add per
  // (point, row) and CTA.  For small local bases (nb ~ 150) this epilogue moves as many bytes as the K loop.
  constexpr int X_LD = GM + 2; // (col * X_LD + row): conflict-free for the fragment stores and for the row-wise reads
  __syncthreads();             // every warp is done reading the last stages
  {
    double *Xs = smem;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        double *x = Xs + (wn * 32 + ni * 8 + 2 * tig) * X_LD + wm * 32 + mi * 8 + gid;
        x[0] = acc[mi][ni][0];
        x[X_LD] = acc[mi][ni][1];
      }
  }
  __syncthreads();
  const int p = tid & (TILE_P - 1), half = tid >> 7;
  const double *Xr = smem + (half * 32) * X_LD + p;
  double *accp = P.acc + (static_cast<size_t>(it.tile) * P.nspin + it.spin) * NACC * TILE_P;
  const double *phib = P.scratch + t.scr_off + static_cast<size_t>(n0 + half * 32) * TILE_P + p;
  double *part = smem + GN * X_LD; // [4][128] partial sums of the upper half
  if (P.debug & 2) return;
  // columns of this half that exist (in groups of 8; the slabs are zero-padded to nbp, the X columns beyond nb are zero):
  // for local bases of 100-250 functions the last 64-column block is mostly padding, and its slab loads are skipped
  const int jv = min(32, max(0, nb - (n0 + half * 32)));
  if (it.acomp == 0) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (P.ncomp > 1) {
      for (int j0 = 0; j0 < jv; j0 += 8) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = j0 + jj;
          const double x = Xr[j * X_LD];
          const double *q = phib + j * TILE_P;
          s0 = fma(q[0], x, s0);
          s1 = fma(q[slab], x, s1);
          s2 = fma(q[2 * slab], x, s2);
          s3 = fma(q[3 * slab], x, s3);
        }
      }
    } else {
      for (int j0 = 0; j0 < jv; j0 += 8) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) s0 = fma(phib[(j0 + jj) * TILE_P], Xr[(j0 + jj) * X_LD], s0);
      }
    }
    if (half) {
      part[p] = s0;
      part[TILE_P + p] = s1;
      part[2 * TILE_P + p] = s2;
      part[3 * TILE_P + p] = s3;
    }
    __syncthreads();
    if (!half) {
      XC_ACC_ADD(P, accp + p, s0 + part[p]);
      if (P.ncomp > 1) {
        XC_ACC_ADD(P, accp + TILE_P + p, s1 + part[TILE_P + p]);
        XC_ACC_ADD(P, accp + 2 * TILE_P + p, s2 + part[2 * TILE_P + p]);
        XC_ACC_ADD(P, accp + 3 * TILE_P + p, s3 + part[3 * TILE_P + p]);
      }
    }
  } else { // tau term: rowsum(phi_k o (phi_k D)) accumulates in row 4
    const double *q = phib + static_cast<size_t>(it.acomp) * slab;
    double s0 = 0.0;
    for (int j0 = 0; j0 < jv; j0 += 8) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) s0 = fma(q[(j0 + jj) * TILE_P], Xr[(j0 + jj) * X_LD], s0);
    }
    if (half) part[p] = s0;
    __syncthreads();
    if (!half) XC_ACC_ADD(P, accp + 4 * TILE_P + p, s0 + part[p]);
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) XC_SYM(k_mo_psi)(BuildParams P, const DensItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  XC_SYM(gemm_mo_body)<2>(P, items[blockIdx.x], smem);
}
__global__ void __launch_bounds__(NTHREADS, 2) XC_SYM(k_mo_density)(BuildParams P, const DensItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  XC_SYM(gemm_mo_body)<3>(P, items[blockIdx.x], smem);
}
#ifndef XC_DETERMINISTIC
__global__ void __launch_bounds__(NTHREADS, 2) k_mo_grad_psi(BuildParams P, const DensItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  gemm_mo_body<4>(P, items[blockIdx.x], smem);
}
__global__ void __launch_bounds__(NTHREADS, 2) k_mo_xc_gradient(BuildParams P, const DensItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  gemm_mo_body<5>(P, items[blockIdx.x], smem);
}
#endif

void XC_SYM(launch_mo_psi)(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = GSTAGES * D_STAGE * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(XC_SYM(k_mo_psi), smem_bytes);
  XC_SYM(k_mo_psi)<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}

void XC_SYM(launch_mo_density)(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = GSTAGES * D_STAGE * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(XC_SYM(k_mo_density), smem_bytes);
  XC_SYM(k_mo_density)<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}

#ifndef XC_DETERMINISTIC
void launch_mo_grad_psi(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = GSTAGES * D_STAGE * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(k_mo_grad_psi, smem_bytes);
  k_mo_grad_psi<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}

void launch_mo_grad_gemm(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = GSTAGES * D_STAGE * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(k_mo_xc_gradient, smem_bytes);
  k_mo_xc_gradient<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}

// rectangular transpose: out (ncol x nrow, ld_out) = in (nrow x ncol, ld_in)^T, column-major
__global__ void k_transpose_rect(const double *__restrict__ in, double *__restrict__ out, int nrow, int ncol, int ld_in,
                                 int ld_out) {
  __shared__ double tl[32][33];
  const int bi = blockIdx.x * 32, bj = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + tx, j = bj + r;
    tl[r][tx] = (i < nrow && j < ncol) ? in[static_cast<size_t>(j) * ld_in + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = bj + tx, i = bi + r; // out(j, i) = in(i, j)
    if (i < nrow && j < ncol) out[static_cast<size_t>(i) * ld_out + j] = tl[tx][r];
  }
}
void launch_transpose_rect(const double *in, double *out, int nrow, int ncol, int ld_in, int ld_out, void *stream) {
  if (nrow <= 0 || ncol <= 0) return;
  dim3 grid((nrow + 31) / 32, (ncol + 31) / 32);
  k_transpose_rect<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(in, out, nrow, ncol, ld_in, ld_out);
}
#endif

} // namespace xcb
