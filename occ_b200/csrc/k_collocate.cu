// k_collocate.cu — phase 1 of the XC build: AO values and gradients of the significant shells of each grid
// tile (reference: gto::evaluate_basis_subset, src/gto/gto.cpp:212-287, whose arithmetic is gau2grid's
// gg_collocation / gg_collocation_deriv1, src/3rdparty/gau2grid/gau2grid_phi.c, gau2grid_deriv1.c, followed
// by the common_fac scaling of gto.cpp:7-16).
//
// B200 mapping: one thread per grid point (128-thread CTAs, one CTA per (tile, run of 16 shells)), so the
// shell parameters are warp-uniform broadcast loads and every store of a basis-function column is a fully
// coalesced 1 KiB line burst into the batch workspace (SoA points, column-major 128 x nbp slabs).
// Radial part: S0 = sum_p c_p exp(-a_p r^2), S1 = sum_p (-2 a_p) c_p exp(-a_p r^2); angular part: Cartesian
// monomials in CCA order (lx descending, then ly), product rule for the gradient; optional real-solid-harmonic
// transform (CCA order m = -l..l) from a constant-memory table.
#include "xc_internal.h"
#include <cuda_runtime.h>

namespace xcb {

__constant__ C2STable c_c2s;

void upload_c2s_table(const C2STable &t) { cudaMemcpyToSymbol(c_c2s, &t, sizeof(C2STable)); }

__device__ __forceinline__ double common_fac_dev(int l, int pure) {
  // gto.cpp:7-16
  if (l == 0) return 0.282094791773878143;
  if (l == 1) return 0.488602511902919921;
  return pure ? sqrt((2 * l + 1) / (4.0 * 3.14159265358979323846)) : 1.0;
}

template <int L, bool DERIV, bool PURE>
__device__ __forceinline__ void shell_angular(double x, double y, double z, double S0, double S1, double fac,
                                              double *__restrict__ o0, double *__restrict__ ox,
                                              double *__restrict__ oy, double *__restrict__ oz) {
  constexpr int NC = (L + 1) * (L + 2) / 2;
  double xp[L + 1], yp[L + 1], zp[L + 1];
  xp[0] = yp[0] = zp[0] = 1.0;
#pragma unroll
  for (int i = 1; i <= L; ++i) {
    xp[i] = xp[i - 1] * x;
    yp[i] = yp[i - 1] * y;
    zp[i] = zp[i - 1] * z;
  }
  const double SX = S1 * x, SY = S1 * y, SZ = S1 * z;
  double c0[NC], cx[DERIV ? NC : 1], cy[DERIV ? NC : 1], cz[DERIV ? NC : 1];
  int idx = 0;
#pragma unroll
  for (int lx = L; lx >= 0; --lx) {
#pragma unroll
    for (int ly = L - lx; ly >= 0; --ly) {
      const int lz = L - lx - ly;
      const double A = xp[lx] * yp[ly] * zp[lz];
      c0[idx] = S0 * A;
      if (DERIV) {
        const double AX = lx ? lx * (xp[lx ? lx - 1 : 0] * yp[ly] * zp[lz]) : 0.0;
        const double AY = ly ? ly * (xp[lx] * yp[ly ? ly - 1 : 0] * zp[lz]) : 0.0;
        const double AZ = lz ? lz * (xp[lx] * yp[ly] * zp[lz ? lz - 1 : 0]) : 0.0;
        cx[idx] = SX * A + S0 * AX;
        cy[idx] = SY * A + S0 * AY;
        cz[idx] = SZ * A + S0 * AZ;
      }
      ++idx;
    }
  }
  if (!PURE || L == 0) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      o0[i * TILE_P] = c0[i] * fac;
      if (DERIV) {
        ox[i * TILE_P] = cx[i] * fac;
        oy[i * TILE_P] = cy[i] * fac;
        oz[i * TILE_P] = cz[i] * fac;
      }
    }
  } else {
    constexpr int NS = 2 * L + 1;
    const double *tab = c_c2s.coef + c_c2s.off[L];
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      double v = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const double t = tab[m * NC + i];
        v += t * c0[i];
        if (DERIV) {
          vx += t * cx[i];
          vy += t * cy[i];
          vz += t * cz[i];
        }
      }
      o0[m * TILE_P] = v * fac;
      if (DERIV) {
        ox[m * TILE_P] = vx * fac;
        oy[m * TILE_P] = vy * fac;
        oz[m * TILE_P] = vz * fac;
      }
    }
  }
}

template <bool DERIV, bool PURE>
__device__ __forceinline__ void shell_dispatch(int L, double x, double y, double z, double S0, double S1, double fac,
                                               double *o0, double *ox, double *oy, double *oz) {
  switch (L) {
  case 0: shell_angular<0, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  case 1: shell_angular<1, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  case 2: shell_angular<2, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  case 3: shell_angular<3, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  case 4: shell_angular<4, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  case 5: shell_angular<5, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  case 6: shell_angular<6, DERIV, PURE>(x, y, z, S0, S1, fac, o0, ox, oy, oz); break;
  default: break;
  }
}

template <bool DERIV, bool PURE>
__global__ void __launch_bounds__(TILE_P) k_collocate(BuildParams P, const CollocItem *__restrict__ items) {
  const CollocItem it = items[blockIdx.x];
  const TileDesc t = P.tiles[it.tile];
  const int p = threadIdx.x;
  const bool valid = p < t.count;
  const double px = P.px[t.pt0 + p], py = P.py[t.pt0 + p], pz = P.pz[t.pt0 + p];
  const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
  double *base = P.scratch + t.scr_off + p;
  const DeviceBasis &B = P.basis;
  for (int s = it.s0; s < it.s1; ++s) {
    const int sh = __ldg(P.shell_list + t.sh_off + s);
    const int col = __ldg(P.shell_col + t.sh_off + s);
    const int L = __ldg(B.l + sh);
    const int np = __ldg(B.nprim + sh);
    const int po = __ldg(B.prim_off + sh);
    const double x = px - __ldg(B.origin + 3 * sh), y = py - __ldg(B.origin + 3 * sh + 1),
                 z = pz - __ldg(B.origin + 3 * sh + 2);
    double r2 = x * x;
    r2 += y * y;
    r2 += z * z;
    double S0 = 0.0, S1 = 0.0;
    for (int k = 0; k < np; ++k) {
      const double a = __ldg(B.exps + po + k);
      const double T1 = __ldg(B.coefs + po + k) * exp(-a * r2);
      S0 += T1;
      if (DERIV) S1 += (-2.0 * a) * T1;
    }
    if (!valid) {
      S0 = 0.0;
      S1 = 0.0;
    }
    double *o0 = base + static_cast<size_t>(col) * TILE_P;
    shell_dispatch<DERIV, PURE>(L, x, y, z, S0, S1, common_fac_dev(L, PURE), o0, o0 + slab, o0 + 2 * slab,
                                o0 + 3 * slab);
  }
  if (it.last) {
    for (int col = t.nb; col < t.nbp; ++col) {
      double *o0 = base + static_cast<size_t>(col) * TILE_P;
      o0[0] = 0.0;
      if (DERIV) {
        o0[slab] = 0.0;
        o0[2 * slab] = 0.0;
        o0[3 * slab] = 0.0;
      }
    }
  }
}

// ---- AO Laplacians (occ::density::evaluate_density<2>, src/gto/density.cpp:84-100: the laplacian column of the
// density-on-grid consumers, include/occ/gto/density.h:67-107).  For phi = S(r^2) A(x, y, z), A a monomial of degree L:
//   lap phi = A (S2 r^2 + (3 + 2 L) S1) + S0 lap A,   S1 = sum (-2a) c e^(-a r^2), S2 = sum (-2a)^2 c e^(-a r^2)
// (grad S . grad A = S1 L A by Euler's theorem).  Writes lap phi to slab 0 and phi to slab 1, the layout the density
// kernel needs to form rowsum(phi o (lap phi . D)) with its ordinary acomp = 0 items (accumulator row 1).
template <int L, bool PURE>
__device__ __forceinline__ void shell_laplacian(double x, double y, double z, double r2, double S0, double S1, double S2,
                                                double fac, double *__restrict__ olap, double *__restrict__ ophi) {
  constexpr int NC = (L + 1) * (L + 2) / 2;
  double xp[L + 1], yp[L + 1], zp[L + 1];
  xp[0] = yp[0] = zp[0] = 1.0;
#pragma unroll
  for (int i = 1; i <= L; ++i) {
    xp[i] = xp[i - 1] * x;
    yp[i] = yp[i - 1] * y;
    zp[i] = zp[i - 1] * z;
  }
  const double radial = S2 * r2 + (3.0 + 2.0 * L) * S1;
  double c0[NC], cl[NC];
  int idx = 0;
#pragma unroll
  for (int lx = L; lx >= 0; --lx) {
#pragma unroll
    for (int ly = L - lx; ly >= 0; --ly) {
      const int lz = L - lx - ly;
      const double A = xp[lx] * yp[ly] * zp[lz];
      double lapA = 0.0;
      if (lx >= 2) lapA += lx * (lx - 1) * (xp[lx >= 2 ? lx - 2 : 0] * yp[ly] * zp[lz]);
      if (ly >= 2) lapA += ly * (ly - 1) * (xp[lx] * yp[ly >= 2 ? ly - 2 : 0] * zp[lz]);
      if (lz >= 2) lapA += lz * (lz - 1) * (xp[lx] * yp[ly] * zp[lz >= 2 ? lz - 2 : 0]);
      c0[idx] = S0 * A;
      cl[idx] = radial * A + S0 * lapA;
      ++idx;
    }
  }
  if (!PURE || L == 0) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      ophi[i * TILE_P] = c0[i] * fac;
      olap[i * TILE_P] = cl[i] * fac;
    }
  } else {
    constexpr int NS = 2 * L + 1;
    const double *tab = c_c2s.coef + c_c2s.off[L];
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      double v = 0.0, vl = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const double t = tab[m * NC + i];
        v += t * c0[i];
        vl += t * cl[i];
      }
      ophi[m * TILE_P] = v * fac;
      olap[m * TILE_P] = vl * fac;
    }
  }
}

template <bool PURE>
__global__ void __launch_bounds__(TILE_P) k_collocate_laplacian(BuildParams P, const CollocItem *__restrict__ items) {
  const CollocItem it = items[blockIdx.x];
  const TileDesc t = P.tiles[it.tile];
  const int p = threadIdx.x;
  const bool valid = p < t.count;
  const double px = P.px[t.pt0 + p], py = P.py[t.pt0 + p], pz = P.pz[t.pt0 + p];
  const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
  double *base = P.scratch + t.scr_off + p;
  const DeviceBasis &B = P.basis;
  for (int s = it.s0; s < it.s1; ++s) {
    const int sh = __ldg(P.shell_list + t.sh_off + s);
    const int col = __ldg(P.shell_col + t.sh_off + s);
    const int L = __ldg(B.l + sh);
    const int np = __ldg(B.nprim + sh);
    const int po = __ldg(B.prim_off + sh);
    const double x = px - __ldg(B.origin + 3 * sh), y = py - __ldg(B.origin + 3 * sh + 1),
                 z = pz - __ldg(B.origin + 3 * sh + 2);
    const double r2 = x * x + y * y + z * z;
    double S0 = 0.0, S1 = 0.0, S2 = 0.0;
    for (int k = 0; k < np; ++k) {
      const double a = __ldg(B.exps + po + k);
      const double T1 = __ldg(B.coefs + po + k) * exp(-a * r2);
      S0 += T1;
      S1 += (-2.0 * a) * T1;
      S2 += (4.0 * a * a) * T1;
    }
    if (!valid) S0 = S1 = S2 = 0.0;
    double *o = base + static_cast<size_t>(col) * TILE_P;
    const double fac = common_fac_dev(L, PURE);
    switch (L) {
    case 0: shell_laplacian<0, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    case 1: shell_laplacian<1, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    case 2: shell_laplacian<2, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    case 3: shell_laplacian<3, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    case 4: shell_laplacian<4, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    case 5: shell_laplacian<5, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    case 6: shell_laplacian<6, PURE>(x, y, z, r2, S0, S1, S2, fac, o, o + slab); break;
    default: break;
    }
  }
  if (it.last)
    for (int col = t.nb; col < t.nbp; ++col) {
      double *o = base + static_cast<size_t>(col) * TILE_P;
      o[0] = 0.0;
      o[slab] = 0.0;
    }
}

void launch_collocate_laplacian(const BuildParams &P, const CollocItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (P.basis.pure) k_collocate_laplacian<true><<<nitems, TILE_P, 0, st>>>(P, items);
  else k_collocate_laplacian<false><<<nitems, TILE_P, 0, st>>>(P, items);
}

void launch_collocate(const BuildParams &P, const CollocItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool deriv = P.ncomp > 1;
  if (deriv && P.basis.pure) k_collocate<true, true><<<nitems, TILE_P, 0, st>>>(P, items);
  else if (deriv) k_collocate<true, false><<<nitems, TILE_P, 0, st>>>(P, items);
  else if (P.basis.pure) k_collocate<false, true><<<nitems, TILE_P, 0, st>>>(P, items);
  else k_collocate<false, false><<<nitems, TILE_P, 0, st>>>(P, items);
}

} // namespace xcb

// ------------------------------------------------------------------------------------------------
// Second pass for the XC nuclear gradient (reference: accumulate_gradient_contribution_{R,U}<1|2>,
// src/dft/dft_gradient_kernels.cpp:47-157,214-470): the AO second derivatives are recomputed from the shell
// parameters and immediately folded with the per-point functional coefficients, so that only the combinations
//   H[s][d]  = c0_s phi_d + sum_e c_{s,e} phi_{d e}        (the "aow_d" of the reference, d = x, y, z)
//   S[s][de] = 2 ct_s phi_{d e}                             (meta-GGA: vtau/2 phi_{d e})
// reach HBM (3 or 9 slabs per spin) instead of the six raw second-derivative slabs.
// ------------------------------------------------------------------------------------------------
namespace xcb {

template <int L, bool PURE>
__device__ __forceinline__ void shell_hessian_combo(double x, double y, double z, double S0, double S1, double S2,
                                                    double fac, const double *__restrict__ cf /* c0,cx,cy,cz,ct */,
                                                    bool mgga, double *__restrict__ oH, size_t slab,
                                                    double *__restrict__ oS) {
  constexpr int NC = (L + 1) * (L + 2) / 2;
  constexpr int NS = PURE && L > 0 ? 2 * L + 1 : NC;
  double xp[L + 1], yp[L + 1], zp[L + 1];
  xp[0] = yp[0] = zp[0] = 1.0;
#pragma unroll
  for (int i = 1; i <= L; ++i) {
    xp[i] = xp[i - 1] * x;
    yp[i] = yp[i - 1] * y;
    zp[i] = zp[i - 1] * z;
  }
  const double q[3] = {x, y, z};
  double h[3][NC], s2[6][NC];
  int idx = 0;
#pragma unroll
  for (int lx = L; lx >= 0; --lx) {
#pragma unroll
    for (int ly = L - lx; ly >= 0; --ly) {
      const int lz = L - lx - ly;
      const int e[3] = {lx, ly, lz};
      const double A = xp[lx] * yp[ly] * zp[lz];
      // first derivatives of the monomial
      const double A1[3] = {lx ? lx * (xp[lx ? lx - 1 : 0] * yp[ly] * zp[lz]) : 0.0,
                            ly ? ly * (xp[lx] * yp[ly ? ly - 1 : 0] * zp[lz]) : 0.0,
                            lz ? lz * (xp[lx] * yp[ly] * zp[lz ? lz - 1 : 0]) : 0.0};
      double d1[3], d2[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) d1[k] = S1 * q[k] * A + S0 * A1[k];
      // second derivatives of the monomial
      const double Axx = lx > 1 ? lx * (lx - 1) * (xp[lx > 1 ? lx - 2 : 0] * yp[ly] * zp[lz]) : 0.0;
      const double Ayy = ly > 1 ? ly * (ly - 1) * (xp[lx] * yp[ly > 1 ? ly - 2 : 0] * zp[lz]) : 0.0;
      const double Azz = lz > 1 ? lz * (lz - 1) * (xp[lx] * yp[ly] * zp[lz > 1 ? lz - 2 : 0]) : 0.0;
      const double Axy = (lx && ly) ? lx * ly * (xp[lx ? lx - 1 : 0] * yp[ly ? ly - 1 : 0] * zp[lz]) : 0.0;
      const double Axz = (lx && lz) ? lx * lz * (xp[lx ? lx - 1 : 0] * yp[ly] * zp[lz ? lz - 1 : 0]) : 0.0;
      const double Ayz = (ly && lz) ? ly * lz * (xp[lx] * yp[ly ? ly - 1 : 0] * zp[lz ? lz - 1 : 0]) : 0.0;
      d2[0] = S2 * x * x * A + S1 * (2.0 * x * A1[0] + A) + S0 * Axx;
      d2[1] = S2 * x * y * A + S1 * (x * A1[1] + y * A1[0]) + S0 * Axy;
      d2[2] = S2 * x * z * A + S1 * (x * A1[2] + z * A1[0]) + S0 * Axz;
      d2[3] = S2 * y * y * A + S1 * (2.0 * y * A1[1] + A) + S0 * Ayy;
      d2[4] = S2 * y * z * A + S1 * (y * A1[2] + z * A1[1]) + S0 * Ayz;
      d2[5] = S2 * z * z * A + S1 * (2.0 * z * A1[2] + A) + S0 * Azz;
      (void)e;
      h[0][idx] = cf[0] * d1[0] + cf[1] * d2[0] + cf[2] * d2[1] + cf[3] * d2[2];
      h[1][idx] = cf[0] * d1[1] + cf[1] * d2[1] + cf[2] * d2[3] + cf[3] * d2[4];
      h[2][idx] = cf[0] * d1[2] + cf[1] * d2[2] + cf[2] * d2[4] + cf[3] * d2[5];
      if (mgga) {
#pragma unroll
        for (int k = 0; k < 6; ++k) s2[k][idx] = 2.0 * cf[4] * d2[k];
      }
      ++idx;
    }
  }
#pragma unroll
  for (int m = 0; m < NS; ++m) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double v;
      if (!PURE || L == 0) {
        v = h[d][m < NC ? m : 0];
      } else {
        const double *tab = c_c2s.coef + c_c2s.off[L] + m * NC;
        v = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) v += tab[i] * h[d][i];
      }
      oH[d * slab + m * TILE_P] = v * fac;
    }
    if (mgga) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double v;
        if (!PURE || L == 0) {
          v = s2[k][m < NC ? m : 0];
        } else {
          const double *tab = c_c2s.coef + c_c2s.off[L] + m * NC;
          v = 0.0;
#pragma unroll
          for (int i = 0; i < NC; ++i) v += tab[i] * s2[k][i];
        }
        oS[k * slab + m * TILE_P] = v * fac;
      }
    }
  }
}

template <bool PURE>
__global__ void __launch_bounds__(TILE_P) k_collocate_h(BuildParams P, const CollocItem *__restrict__ items) {
  const CollocItem it = items[blockIdx.x];
  const TileDesc t = P.tiles[it.tile];
  if (P.leaf_skip[t.leaf]) return;
  const int p = threadIdx.x;
  const bool valid = p < t.count;
  const double px = P.px[t.pt0 + p], py = P.py[t.pt0 + p], pz = P.pz[t.pt0 + p];
  const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
  double *base = P.scratch + t.scr_off + p;
  const DeviceBasis &B = P.basis;
  const bool mgga = P.deriv >= 2;
  for (int sp = 0; sp < P.nspin; ++sp) {
    double cf[5];
    const double *cg = P.coef + (static_cast<size_t>(it.tile) * P.nspin + sp) * NCOEF * TILE_P + p;
#pragma unroll
    for (int k = 0; k < 5; ++k) cf[k] = cg[k * TILE_P];
    if (!mgga) cf[4] = 0.0;
    double *oH0 = base + static_cast<size_t>(P.slab_H + 3 * sp) * slab;
    double *oS0 = base + static_cast<size_t>(P.slab_S + 6 * sp) * slab;
    for (int s = it.s0; s < it.s1; ++s) {
      const int sh = __ldg(P.shell_list + t.sh_off + s);
      const int col = __ldg(P.shell_col + t.sh_off + s);
      const int L = __ldg(B.l + sh);
      const int np = __ldg(B.nprim + sh);
      const int po = __ldg(B.prim_off + sh);
      const double x = px - __ldg(B.origin + 3 * sh), y = py - __ldg(B.origin + 3 * sh + 1),
                   z = pz - __ldg(B.origin + 3 * sh + 2);
      const double r2 = x * x + y * y + z * z;
      double S0 = 0.0, S1 = 0.0, S2 = 0.0;
      for (int k = 0; k < np; ++k) {
        const double a = __ldg(B.exps + po + k);
        const double T1 = __ldg(B.coefs + po + k) * exp(-a * r2);
        const double a2 = -2.0 * a;
        S0 += T1;
        S1 += a2 * T1;
        S2 += a2 * a2 * T1;
      }
      if (!valid) S0 = S1 = S2 = 0.0;
      const double fac = common_fac_dev(L, PURE);
      double *oH = oH0 + static_cast<size_t>(col) * TILE_P;
      double *oS = oS0 + static_cast<size_t>(col) * TILE_P;
      switch (L) {
      case 0: shell_hessian_combo<0, PURE>(x, y, z, S0, S1, S2, fac, cf, mgga, oH, slab, oS); break;
      case 1: shell_hessian_combo<1, PURE>(x, y, z, S0, S1, S2, fac, cf, mgga, oH, slab, oS); break;
      case 2: shell_hessian_combo<2, PURE>(x, y, z, S0, S1, S2, fac, cf, mgga, oH, slab, oS); break;
      case 3: shell_hessian_combo<3, PURE>(x, y, z, S0, S1, S2, fac, cf, mgga, oH, slab, oS); break;
      case 4: shell_hessian_combo<4, PURE>(x, y, z, S0, S1, S2, fac, cf, mgga, oH, slab, oS); break;
      default: break; // l > 4 is rejected at create() for gradient handles
      }
    }
    if (it.last) {
      const int nsl = mgga ? 9 : 3;
      for (int col = t.nb; col < t.nbp; ++col)
        for (int k = 0; k < nsl; ++k) {
          double *o = (k < 3 ? oH0 + k * slab : oS0 + (k - 3) * slab) + static_cast<size_t>(col) * TILE_P;
          o[0] = 0.0;
        }
    }
  }
}

void launch_collocate_h(const BuildParams &P, const CollocItem *items, int nitems, void *stream) {
  if (nitems <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (P.basis.pure) k_collocate_h<true><<<nitems, TILE_P, 0, st>>>(P, items);
  else k_collocate_h<false><<<nitems, TILE_P, 0, st>>>(P, items);
}

} // namespace xcb
