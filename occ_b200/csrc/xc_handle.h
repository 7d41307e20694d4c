// xc_handle.h — the handle behind the C ABI (include/occ_xc_b200.h), shared by the host translation units
// (xc_host.cu: create / build / gradient / VV10;  xc_multi.cu: several handles of one process, one per GPU).
#pragma once
#include "../../include/occ_xc_b200.h"
#include "xc_internal.h"
#include <cuda_runtime.h>
#include <vector>

struct Batch {
  int tile0, ntiles, leaf0, nleaves;
  size_t colloc_off, ncolloc, dens_off, ndens, g_off, ng, vxc_off, nvxc, grad_off, ngrad;
  size_t densmo_off, ndensmo, psi_off, npsi, xpsi_off, nxpsi; // MO-form density work lists
  // MO-form gradient: per A operand wh (0 phi, 1 G, 2..4 phi_e) the psi_A items and the X items; D-form remainder
  size_t gpsi_off[5], ngpsi[5], gx_off[5], ngx[5], gradd_off, ngradd;
};

struct occ_xc_handle {
  int device = 0;
  int nbf = 0, nsh = 0, nspin = 1, deriv = 0, ncomp = 1, pure = 0;
  size_t npts = 0;
  int nleaves = 0, ntiles = 0;
  double sum_n_nb2 = 0.0;
  int nb_min = 0, nb_max = 0;
  std::vector<Batch> batches;
  std::vector<xcb::TileDesc> h_tiles;
  std::vector<xcb::LeafDesc> h_leaves;
  std::vector<size_t> leaf_global; // caller's leaf index of each local leaf
  std::vector<int> h_bf_index;
  xcb::BuildParams P;
  std::vector<void *> allocs;
  size_t alloc_bytes = 0;
  xcb::CollocItem *d_colloc = nullptr;
  xcb::DensItem *d_dens = nullptr, *d_g = nullptr, *d_gradit = nullptr;
  bool gradient = false;
  int natoms = 0;
  double *d_DT = nullptr, *d_grad = nullptr;
  xcb::VxcItem *d_vxc = nullptr;
  double *d_D = nullptr, *d_Dsym = nullptr, *d_Ktmp = nullptr, *d_K = nullptr, *d_scalars = nullptr;
  cudaStream_t stream = nullptr, aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double *scratch2 = nullptr; // workspace of the second lane
  int lanes = 1;
  bool profiling = false;
  bool deterministic = false, nvtx = false;
  bool vxc_ws = false; // Vxc contraction through the warp-specialised persistent kernel (k_vxc_ws.cu)
  double density_scale = 1.0, orbital_scale = 1.0;
  double phase_ms[OCC_XC_NPHASES] = {0, 0, 0, 0, 0, 0, 0};
  size_t launches = 0;
  // MO-form density (options.mo_form with the occupied-orbital counts at create)
  bool mo = false;
  int nocc[2] = {0, 0}, noccp = 0;
  xcb::DensItem *d_densmo = nullptr, *d_psi = nullptr, *d_xpsi = nullptr;
  xcb::DensItem *d_gpsi[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *d_gx[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  xcb::DensItem *d_gradd = nullptr;
  double *d_C = nullptr, *d_CT = nullptr;
  double mo_executed_flops = 0.0; // DMMA flops one MO-form build really issues (density side)
  double vxc_executed_flops = 0.0; // DMMA flops of the Vxc phase (upper triangles only for the symmetric terms)
  // post-SCF VV10 (allocated on first use)
  std::vector<size_t> h_tile_src; // sorted index of the first point of each tile
  std::vector<long long> h_tile_psi_off;
  size_t npts_total = 0;          // points of the caller's whole grid (orig_index range)
  int world = 1;
  size_t *d_tile_src = nullptr;
  xcb::VV10Params vv = {};
  int2 *d_vv_items = nullptr;
  int n_vv_items = 0;
  size_t *d_vv_orig = nullptr;
  xcb::BuildParams vvP;           // build parameters of the running VV10 call (D pointers set by the density stage)
  int2 *d_vv_grid_items = nullptr; // full-square work list of the gradient's grid-displacement term
  int n_vv_grid_items = 0;
  size_t *d_vv_atom_off = nullptr;
};

// post-SCF VV10 in stages (xc_host.cu); every stage only enqueues on h->stream
int vv10_check_handle(occ_xc_handle *h, const char *who, bool gradient);
int vv10_setup(occ_xc_handle *h, int part, int nparts);
int vv10_setup_gradient(occ_xc_handle *h, int part, int nparts, const size_t *atom_offsets);
int vv10_stage_density(occ_xc_handle *h, const double *D2, const size_t *orig_index, double vv10_b, double vv10_C,
                       double unowned, bool with_transpose);
int vv10_stage_pairs(occ_xc_handle *h, const xcb::VV10Params &in);
int vv10_stage_vxc(occ_xc_handle *h, const xcb::VV10Params &sums);
int vv10_stage_gradient(occ_xc_handle *h, const xcb::VV10Params &merged, const xcb::VV10Params &sums);


// one build on device buffers (xc_host.cu); D2_dev may be NULL when C_dev is given (D2 = C C^T formed on the device)
int xc_build_device_impl(occ_xc_handle *h, const double *D2_dev, const double *C_dev, double *K_dev, double *scalars_dev,
                         void *cuda_stream);
int xc_gradient_impl(occ_xc_handle *h, const double *D2, const double *C, double *gradient); // host D2 (, C) -> host gradient
int occ_xc_set_error(const char *msg); // sets the calling thread's occ_xc_last_error() and returns 1
