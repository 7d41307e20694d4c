// xc_internal.h — device-side data layout of the B200 XC build (shared by the host orchestration and the kernels).
//
// Work decomposition (all of it geometry-only, built once in occ_xc_create):
//   leaf   = one Morton leaf of the reference's SpatialGridHierarchy (the semantic unit: shell list, density skip)
//   tile   = <=128 points of one leaf (every leaf but the last is exactly one tile)
//   batch  = a run of consecutive leaves whose AO-value workspace fits the batch budget; the four phase
//            kernels (collocate -> density -> functional -> vxc) are launched once per batch
// AO values of a tile live in the batch workspace as NCOMP column-major (128 x nbp) slabs
//   phi_c[col * 128 + p],  c = 0 value, 1..3 d/dx,d/dy,d/dz;  nbp = nb rounded up to 64, padding columns zero.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector_types.h>

namespace xcb {

constexpr int TILE_P = 128; // points per tile (the reference's leaf size, spatial_grid_hierarchy.h:26-30)
constexpr int NB_PAD = 64;  // local basis functions padded to a multiple of this
constexpr int MAX_L = 6;
constexpr int MAX_FUNC = 4;

struct TileDesc {
  int leaf;          // local leaf index
  int pt0;           // first padded point slot (= tile * 128)
  int count;         // valid points
  int nb, nbp;       // local basis functions, padded
  int bf_off;        // offset of the leaf's local->global bf map in bf_index[]
  int sh_off, nsh;   // the leaf's significant shells in shell_list[] / shell_col[]
  long long scr_off; // offset (doubles) of phi_0 in the batch workspace; phi_c at scr_off + c * nbp * 128
};

struct LeafDesc {
  int tile0, ntiles; // tiles of this leaf
  int nb;
  int pad;
};

// one CTA of the collocation kernel: shells [s0, s1) of the tile's list; `last` also zeroes padding columns
struct CollocItem {
  int tile;
  int s0, s1;
  int last;
};
// one CTA of the density kernel: X(:, 64*jb..) = phi_acomp . D_spin, row-dotted with the AO slabs
struct DensItem {
  int tile;
  short jb;
  signed char spin, acomp;
};
// one CTA of the Vxc kernel: Ktmp[:, 64*jb..] += phi_acomp^T . B_jb
struct VxcItem {
  int tile;
  short jb;
  signed char spin, acomp;
};

struct DeviceBasis {
  const int *l;          // [nsh]
  const int *nprim;      // [nsh]
  const int *prim_off;   // [nsh]
  const int *first_bf;   // [nsh]
  const double *origin;  // [3*nsh]
  const double *exps;
  const double *coefs;
  int pure;
};

// per-point accumulators written by the density kernel, read by the functional kernel: [tile][spin][NACC][128]
constexpr int NACC = 5; // rho, 2 x (drho/dx,dy,dz)/2 ... stored raw: sum phi X, sum phi_x X, .., sum_d phi_d (phi_d D)
// per-point Vxc coefficients written by the functional kernel: [tile][spin][NCOEF][128]
constexpr int NCOEF = 5; // c0 (phi), cx, cy, cz (grad phi), ct (tau term)

struct BuildParams {
  DeviceBasis basis;
  const TileDesc *tiles;
  const LeafDesc *leaves;
  const int *shell_list; // CSR per leaf: global shell index
  const int *shell_col;  // CSR per leaf: first local column of that shell
  const int *bf_index;   // CSR per leaf: local column -> global basis function
  const double *px, *py, *pz, *pw; // [ntiles*128] padded SoA points / weights (w = 0 in padding)
  double *scratch;       // batch workspace
  double *acc;           // [ntiles][nspin][NACC][128]
  double *coef;          // [ntiles][nspin][NCOEF][128]
  int *leaf_skip;        // [nleaves] 1 = leaf failed the density threshold (dft_kernels.cpp:38-40)
  double *leaf_sums;     // [nleaves][3] exc, n_alpha, n_beta
  const double *Drho[2]; // density matrices for the phi.D product (restricted: upper-symmetrised copy)
  const double *Dtau[2]; // for the tau products (restricted: D as given, density.cpp:90-96)
  double *Ktmp[2];       // accumulation target, K = Ktmp + Ktmp^T at the end
  int ldD;               // leading dimension of D / K = nspin * nbf
  int nbf, nspin, deriv, ncomp;
  int nfunc;
  int func_ids[MAX_FUNC];
  double density_threshold;
  // ---- XC nuclear gradient (handles created with the gradient option) ----
  const int *bf_atom;    // CSR per leaf: local column -> atom
  const double *DT[2];   // D_s^T, so that (A . D_s^T) gathers contiguously like the density contraction
  double *grad;          // [3 * natoms] accumulation target
  int slab_G;            // first G slab (= ncomp), then nspin slabs
  int slab_H;            // first H slab: H[s][d] at slab_H + 3 s + d      (GGA / meta-GGA)
  int slab_S;            // first S slab: S[s][de] at slab_S + 6 s + de    (meta-GGA), de in xx,xy,xz,yy,yz,zz order
  double grad_factor;    // -2, or -4 for LDA (dft_gradient_kernels.cpp:27-45 carries no 1/2 on vrho)
  int debug; // developer switches for timing experiments (options.reserved[0]): 1 = no Vxc scatter, 2 = no density epilogue
  // ---- MO-form density (handles created with the occupied-orbital counts) ----
  const double *moC[2];  // C_s: nbf x nocc_s column-major (ld = ldC), D2_s = C_s C_s^T
  const double *moCT[2]; // C_s^T: nocc_s x nbf column-major (ld = ldCT)
  int ldC, ldCT;
  int mo_nocc[2];
  int mo_noccp;          // psi slab width: max nocc rounded up to 64
  const long long *tile_psi_off; // [ntiles] offset of the tile's psi slabs: psi_s at tile_psi_off[tile] + s * noccp * 128
                                 // (kept out of TileDesc: its layout feeds the register allocation of the tuned D-form kernel)
  // ---- deterministic accumulation (options.deterministic): distance in doubles from an accumulator to its low bin
  long long det_acc_off; // acc    -> acc_lo
  long long det_K_off;   // Ktmp   -> Ktmp_lo
};

// ---- post-SCF VV10: per-point arrays in the REFERENCE's point order ----
struct VV10Params {
  size_t n;               // points of the NLC grid
  const size_t *tile_src; // [ntiles] sorted index of the tile's first point
  const size_t *orig;     // [n] sorted index -> index in the reference's (atom-grid) order; NULL = identity
  double *X, *Y, *Z;      // coordinates
  double *rw2;            // 2 rho w (0 for points below the 1e-8 density cut)
  double *om, *kap;       // omega_0, kappa (kappa < 0 marks an inactive point)
  double *F, *U, *W;      // pair sums
  double kappa_factor, beta, C;
  // nuclear gradient, grid-displacement term (allocated on the first occ_xc_vv10_gradient call)
  double *GX, *GY, *GZ;   // per-point sum_j f_ij
  int *pt_atom;           // [n] atom whose grid the point belongs to (reference order)
  double *bbox;           // [nblk][6] bounding boxes of the active points of each block of 128
};
// multi-GPU: base pointers of the shards' (X Y Z rw2 om kap) blocks, each 6 n doubles
struct VV10Peers {
  const double *base[16];
  int n;
};

// spherical transform tables (real solid harmonics, CCA order m = -l..l) in constant memory
struct C2STable {
  // coef[off[l] + m * ncart(l) + c]
  int off[MAX_L + 2];
  double coef[1 + 9 + 30 + 70 + 135 + 231 + 364];
};

void launch_collocate(const BuildParams &P, const CollocItem *items, int nitems, void *stream);
void launch_density(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_functional(const BuildParams &P, int leaf0, int nleaves, void *stream);
void launch_vxc_resident(const BuildParams &P, const VxcItem *items, int nitems, void *stream);
void launch_vxc_ws(const BuildParams &P, const VxcItem *items, int nitems, void *stream); // k_vxc_ws.cu: restricted GGA items only
// order-independent variants (k_*_det.cu)
void launch_density_det(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_vxc_resident_det(const BuildParams &P, const VxcItem *items, int nitems, void *stream);
void launch_mo_psi_det(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_mo_density_det(const BuildParams &P, const DensItem *items, int nitems, void *stream);
// plain FP64 GEMM on the DMMA pipe and elementwise helpers (k_dgemm.cu)
void launch_dgemm(bool b_transposed, int M, int N, int K, double alpha, const double *A, int lda, long long strideA,
                  const double *B, int ldb, long long strideB, double *C, int ldc, long long strideC, int batches,
                  void *stream);
void launch_scale(const double *in, double *out, size_t n, double s, void *stream);
void launch_add_inplace(double *a, const double *b, size_t n, void *stream);
int measure_dmma_peak(double target_ms, double *tflops, double *measured_ms);
void launch_gbuild(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_collocate_h(const BuildParams &P, const CollocItem *items, int nitems, void *stream);
void launch_collocate_laplacian(const BuildParams &P, const CollocItem *items, int nitems, void *stream);
void launch_gather_laplacian(const BuildParams &P, const size_t *tile_src, size_t npts, int ntiles, double *out, void *stream);
void launch_mo_psi(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_mo_density(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_mo_grad_psi(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_mo_grad_gemm(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_transpose_rect(const double *in, double *out, int nrow, int ncol, int ld_in, int ld_out, void *stream);
void launch_grad_gemm(const BuildParams &P, const DensItem *items, int nitems, void *stream);
void launch_transpose(const double *D, double *DT, int nbf, int ld, void *stream);
void launch_prepare_density(const double *D2, double *Dsym, int nbf, int nspin, void *stream);
void launch_finalize(const BuildParams &P, int nleaves, double *K, double *scalars, void *stream);
void launch_gather_rho(const BuildParams &P, const size_t *tile_src, size_t npts, int ntiles, double *out, void *stream);
void launch_vv10_prepare(const BuildParams &P, const VV10Params &V, int ntiles, double unowned, void *stream);
void launch_vv10_point_atoms(const VV10Params &V, int natoms, const size_t *atom_offsets_dev, void *stream);
void launch_vv10_grid_pairs(const VV10Params &V, const int2 *items, int nitems, void *stream);
void launch_vv10_grid_reduce(const VV10Params &V, double *grad, void *stream);
void launch_vv10_merge(const VV10Peers &src, const VV10Params &out, void *stream);
void launch_vv10_pairs(const VV10Params &V, const int2 *items, int nitems, void *stream);
void launch_vv10_coef(const BuildParams &P, const VV10Params &V, int nleaves, void *stream);
int vv10_jchunk();
void upload_c2s_table(const C2STable &t);
int eval_functional_device(int func_id, int polarized, size_t n, const double *rho, const double *sigma,
                           const double *tau, double *exc, double *vrho, double *vsigma, double *vtau);

} // namespace xcb
