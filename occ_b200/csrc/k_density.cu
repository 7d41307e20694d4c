// k_density.cu — the D-form density contraction and the D-form XC-gradient contraction (gemm_xd_body<0>, <1>).
// They live in their own translation unit on purpose: ptxas' instruction order for these two tuned kernels depends on
// the virtual-register numbering of the whole unit, and sharing a unit with the Vxc / factorised kernels cost the
// density kernel 2.5 % (101.4 -> 103.9 ms on valinomycin, profiles/r01_experiments.md) with byte-identical source.
#include "gemm_common.cuh"

namespace xcb {

// ------------------------------------------------------------------------------------------------
// density
// ------------------------------------------------------------------------------------------------
// MODE 0: density (epilogue = row-dots with the AO slabs).  MODE 1: XC nuclear gradient:
// acc = A . D_s^T with A = phi (acomp 0), G_s (acomp 1) or phi_e (acomp 2+e, meta-GGA); the epilogue multiplies by
// the companion slabs (H_s,d / phi_d / S_s,de), sums over the 128 points and scatters per atom:
//   grad[A][d] += f * sum_p sum_{mu in A} companion_d[p, mu] * acc[p, mu]
// which is  -2 sum_{mu in A, nu} D_{mu nu} V_d(mu, nu)  of dft_gradient_kernels.cpp:540-592 without ever forming V_d.
template <int MODE> __device__ __forceinline__ void XC_SYM(gemm_xd_body)(const BuildParams &P, const DensItem it, double *smem) {
  const TileDesc t = P.tiles[it.tile];
  if (MODE == 1 && P.leaf_skip[t.leaf]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int nb = t.nb;
  const int nk = (nb + GK - 1) / GK;
  const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
  const int a_slab = MODE == 0 ? it.acomp : (it.acomp == 0 ? 0 : (it.acomp == 1 ? P.slab_G + it.spin : it.acomp - 1));
  const double *Ag = P.scratch + t.scr_off + static_cast<size_t>(a_slab) * slab;
  // it.spin == 2 (MODE 0, unrestricted builds): the CTA owns 32 basis-function columns for BOTH spins — X columns 0..31 are
  // phi_a D_alpha, 32..63 phi_a D_beta of the same functions — so that the AO slabs of the row-dot epilogue (for local bases
  // of 100-250 functions as many bytes as the K loop) are read once for the two spins
  const bool both = MODE == 0 && it.spin == 2;
  const int s0 = both ? 0 : it.spin;
  const double *Dsp[2];
  Dsp[0] = MODE == 0 ? (it.acomp == 0 ? P.Drho : P.Dtau)[s0] : P.DT[s0];
  Dsp[1] = both ? (it.acomp == 0 ? P.Drho : P.Dtau)[1] : Dsp[0];
  const int *gidx = P.bf_index + t.bf_off;
  const int n0 = it.jb * (both ? GN / 2 : GN);

  // B loader assignment: this thread stages k = kk of every chunk for NBR columns
  constexpr int NROWT = NTHREADS / GK;   // columns covered per pass
  constexpr int NBR = GN / NROWT;        // passes
  const int kk = tid % GK, nn = tid / GK;
  size_t coloff[NBR];
  bool nvalid[NBR];
#pragma unroll
  for (int r = 0; r < NBR; ++r) {
    const int c = nn + NROWT * r;                 // column of the 64-wide block
    const int n = n0 + (both ? (c & 31) : c);     // basis function
    nvalid[r] = n < nb;
    coloff[r] = nvalid[r] ? static_cast<size_t>(__ldg(gidx + n)) * P.ldD : 0;
  }
  // 8-column groups of this warp's 32 columns that hold real basis functions (the rest is padding: no DMMA)
  const int ni_valid = min(4, max(0, (nb - (n0 + (both ? 0 : wn * 32)) + 7) >> 3));

  int grow_next = 0;
  auto load_stage = [&](int kc, int stage) {
    double *As = smem + stage * D_STAGE;
    double *Bs = As + D_ASZ;
#pragma unroll
    for (int r = 0; r < GK * 64 / NTHREADS; ++r) {
      const int seg = tid + NTHREADS * r;
      const int col = seg >> 6, s2 = (seg & 63) * 2;
      cp_async16(As + col * D_LDA + s2, Ag + static_cast<size_t>(kc * GK + col) * TILE_P + s2, kc * GK + col < t.nbp);
    }
    // grow_next = global row of k = kc*GK + kk, fetched one iteration ahead so the gather never waits on it
    const bool kvalid = kc * GK + kk < nb;
    const int grow = grow_next;
#pragma unroll
    for (int r = 0; r < NBR; ++r)
      cp_async8(Bs + (nn + NROWT * r) * D_LDB + kk, Dsp[(NROWT * r) >> 5] + coloff[r] + grow, kvalid && nvalid[r]);
  };
  auto fetch_row = [&](int kc) { grow_next = (kc * GK + kk < nb) ? __ldg(gidx + kc * GK + kk) : 0; };

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

  if (MODE == 1) {
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
  const int cb = both ? n0 : n0 + half * 32; // first basis function of this half's 32 X columns
  double *accp = P.acc + (static_cast<size_t>(it.tile) * P.nspin + (both ? half : it.spin)) * NACC * TILE_P;
  const double *phib = P.scratch + t.scr_off + static_cast<size_t>(cb) * TILE_P + p;
  double *part = smem + GN * X_LD; // [4][128] partial sums of the upper half
  if (P.debug & 2) return;
  // columns of this half that exist (in groups of 8; the slabs are zero-padded to nbp, the X columns beyond nb are zero):
  // for local bases of 100-250 functions the last 64-column block is mostly padding, and its slab loads are skipped
  const int jv = min(32, max(0, nb - cb));
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
    if (both) { // each half holds one spin's sums: no exchange between the halves
      XC_ACC_ADD(P, accp + p, s0);
      if (P.ncomp > 1) {
        XC_ACC_ADD(P, accp + TILE_P + p, s1);
        XC_ACC_ADD(P, accp + 2 * TILE_P + p, s2);
        XC_ACC_ADD(P, accp + 3 * TILE_P + p, s3);
      }
      return;
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
    if (both) {
      XC_ACC_ADD(P, accp + 4 * TILE_P + p, s0);
      return;
    }
    if (half) part[p] = s0;
    __syncthreads();
    if (!half) XC_ACC_ADD(P, accp + 4 * TILE_P + p, s0 + part[p]);
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) XC_SYM(k_density)(BuildParams P, const DensItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  XC_SYM(gemm_xd_body)<0>(P, items[blockIdx.x], smem);
}
#ifndef XC_DETERMINISTIC
__global__ void __launch_bounds__(NTHREADS, 2) k_xc_gradient(BuildParams P, const DensItem *__restrict__ items) {
  extern __shared__ __align__(16) double smem[];
  gemm_xd_body<1>(P, items[blockIdx.x], smem);
}
#endif

void XC_SYM(launch_density)(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = GSTAGES * D_STAGE * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(XC_SYM(k_density), smem_bytes);
  XC_SYM(k_density)<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}

#ifndef XC_DETERMINISTIC
void launch_grad_gemm(const BuildParams &P, const DensItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  const int smem_bytes = GSTAGES * D_STAGE * sizeof(double);
  static SmemOptIn optin; // per device: cudaFuncSetAttribute applies to the current device only
  optin.ensure(k_xc_gradient, smem_bytes);
  k_xc_gradient<<<nitems, NTHREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P, items);
}
#endif

} // namespace xcb
