/*
 * occ_cosx_b200.h — C ABI of the B200-native (sm_100a) seminumerical exchange (COSX) build, SURVEY.md 8(f) rank 2.
 *
 * Drop-in boundary for
 *     Mat occ::qm::cosx::SemiNumericalExchange::compute_K(mo, precision, Schwarz)        (ESP mode)
 *     /root/reference/src/qm/seminumerical_exchange.cpp:481-545,617-660 (dispatch), 660-828 (grid sweep)
 * as called from HartreeFock::compute_JK / compute_K (src/qm/hf.cpp:179-258) once HartreeFock::set_cosx_exchange
 * (hf.cpp:62-68) has installed the engine.  The reference has no FFI here; these entry points are what a C++ shim
 * replacing the body of compute_K binds.  What the reference's constructor builds once per geometry — the
 * Becke-partitioned atom grids and the overlap projector (seminumerical_exchange.cpp:245-262) — is handed over at
 * create(); the per-SCF-iteration work (AO values on every batch, F = (w phi) D2q, the ESP integrals of every
 * significant shell pair at every grid point of the batch, G = F.V, K = phi^T G) runs on the GPU.
 *
 * Same conventions as occ_xc_b200.h: plain pointers and sizes, column-major doubles, 0 on success, message in
 * occ_xc_last_error().  No CPU fallback: create() fails without a CUDA device.
 */
#ifndef OCC_COSX_B200_H
#define OCC_COSX_B200_H
#include "occ_xc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* occ::qm::cosx::Settings (include/occ/qm/seminumerical_exchange.h:16-21) + the constants of the sweep */
typedef struct {
  double margin;            /* Settings::margin, 1.0 Bohr: pair kept when |P_pair - c_batch| - extent < r_batch + margin */
  double density_threshold; /* 1e-6: max |D2q| over the pair's blocks below this skips the pair (:758-764) */
  double pair_threshold;    /* 1e-12: overlap-block norm of a significant shell pair (integral_engine.h:250-281) */
  size_t batch_bytes;       /* device workspace per chunk of grid batches (0 = 4 GiB) */
  int device;               /* CUDA ordinal, -1 = current */
  int collect_stats;        /* 1 = count evaluated (batch, pair) combinations (adds one atomic per combination) */
  int rank, world_size;     /* grid batches are dealt round-robin to ranks; each rank returns a partial K (sum = K) */
  int reserved[4];          /* must be zero */
} occ_cosx_options_t;

typedef struct occ_cosx_handle occ_cosx_handle;

/* SemiNumericalExchange::SemiNumericalExchange (seminumerical_exchange.cpp:245-275), minus what stays on the host:
 *   atom_offsets[natoms + 1]  delimit the Becke-partitioned atom grids (m_atom_grids) inside pts_xyz / weights;
 *                             batches are blocks of 128 consecutive points of one atom grid (:676-699)
 *   overlap_projector         m_overlap_projector = Sn.ldlt().solve(S), nbf x nbf column-major; NULL = identity
 * Shell pairs (precompute_significant_shell_pairs, :293-331), their E coefficients, centres and extents
 * (src/ints/esp.cpp:85-100, include/occ/ints/kernels.h:256-340) are built here, on the host, once. Shells up to l = 4 (g). */
int occ_cosx_create(const occ_xc_basis_t *basis, int natoms, const size_t *atom_offsets, const double *pts_xyz,
                    const double *weights, const double *overlap_projector, const occ_cosx_options_t *opt,
                    occ_cosx_handle **out);

/* SemiNumericalExchange::compute_K.  kind 0 = restricted: D is mo.D (nbf x nbf), K = 0.25 (K' + K'^T) of D2q = P.(2D);
 * kind 1 = unrestricted: D is (2 nbf x nbf) with the alpha block on top, each spin block 0.5 (K' + K'^T) of P.(2 D_s);
 * kind 2 = general: D and K are (2 nbf x 2 nbf) with blocks aa, ab / ba, bb (compute_K_general, :662-712).
 * omega = 0: full Coulomb operator; omega > 0: long-range erf(omega r)/r (set_range_separated_omega). */
int occ_cosx_compute_K(occ_cosx_handle *h, int kind, const double *D, double precision, double omega, double *K);

/* SemiNumericalExchange::compute_K_range_separated (:830-860): (alpha + beta) K[1/r] - beta K[erf(omega r)/r] in one
 * grid sweep (restricted only, like the reference). */
int occ_cosx_compute_K_range_separated(occ_cosx_handle *h, const double *D, double alpha, double beta, double omega,
                                       double precision, double *K);
/* The same with the spinorbital kind of the reference's member (:862-893): kind 0 restricted as above; kind 1
 * unrestricted — D and K are (2 nbf x nbf), one fused sweep per spin block, each block 0.5 (K' + K'^T); kind 2 (general)
 * is refused like the reference ("not implemented for general spinorbitals"). */
int occ_cosx_compute_K_range_separated_kind(occ_cosx_handle *h, int kind, const double *D, double alpha, double beta,
                                            double omega, double precision, double *K);

/* The same two builds for callers that also hand over the occupied orbitals — the reference's seam receives `mo`
 * (compute_K(const MolecularOrbitals &mo, ...), include/occ/qm/seminumerical_exchange.h:108), which holds mo.D and
 * mo.Cocc.  C holds one block per spin, block s nbf x nocc_s column-major (ld = nbf), beta after alpha, scaled so that
 * 2 D_s = C_s C_s^T (restricted: C = sqrt(2) Cocc; unrestricted: C_s = Cocc_s, because mo.D stores the halved blocks —
 * the convention of occ_xc_build_mo).  Same K up to the rounding of C C^T, but F = Phi.D2q runs in the factorised form
 * F = (Phi . (P C)) . C^T: 4 n nbf nocc flops instead of 2 n nbf^2 (valinomycin def2-TZVP, nocc 300 of 3348: 215 -> ~45
 * ms).  D is still required: the density screen of the shell pairs reads D2q = P.(2D) (:758-764).  kind 0 or 1. */
int occ_cosx_compute_K_mo(occ_cosx_handle *h, int kind, const double *D, const double *C, int nocc_alpha, int nocc_beta,
                          double precision, double omega, double *K);
int occ_cosx_compute_K_range_separated_mo(occ_cosx_handle *h, const double *D, const double *C, int nocc, double alpha,
                                          double beta, double omega, double precision, double *K);
int occ_cosx_compute_K_range_separated_mo_kind(occ_cosx_handle *h, int kind, const double *D, const double *C,
                                               int nocc_alpha, int nocc_beta, double alpha, double beta, double omega,
                                               double precision, double *K);

/* compute_K_for_density (:817-828): the raw, unsymmetrised K' of one projected density D2q (nbf x nbf); with
 * w_full / w_lr the fused two-operator form. Device-resident variant: pointers are device memory on h's device. */
int occ_cosx_K_for_density(occ_cosx_handle *h, const double *D2q, double precision, double omega, double *K);
int occ_cosx_K_for_density_device(occ_cosx_handle *h, const double *D2q_dev, double precision, double w_full,
                                  double w_lr, double omega, double *K_dev);

/* SemiNumericalExchange::compute_overlap_matrix (:336-380) on the GPU and the analytic overlap from the pair data
 * (ESPEvaluator::evaluate_overlap, esp.cpp:127-146): the two inputs of the projector, for hosts that want to build
 * it with this library (S_numeric.ldlt().solve(S_analytic)). */
int occ_cosx_numerical_overlap(occ_cosx_handle *h, double *S_numeric);
int occ_cosx_analytic_overlap(occ_cosx_handle *h, double *S_analytic);
int occ_cosx_set_projector(occ_cosx_handle *h, const double *overlap_projector);

typedef struct {
  int nbf, nsh, natoms;
  size_t npts, nbatches;
  size_t npairs;             /* significant shell pairs */
  size_t nprimitive_pairs;
  size_t device_bytes;
  /* of the last compute (collect_stats = 1): */
  size_t pair_evals;         /* (batch, pair) combinations evaluated after both screens */
  size_t skipped_batches;    /* batches with max(phi) < precision (:729-730) */
  double dense_flops;        /* 2 npt nherm nab per evaluated primitive pair: the reference's R.E^T product (esp.cpp:113-121) */
  double gemm_flops;         /* F = phi D2q and K = phi^T G: 4 npt nbf^2 (factorised F: 2 npt nbf^2 + 4 npt nbf nocc) */
  double ms_collocate, ms_fgemm, ms_esp, ms_kgemm; /* device time of the last compute, by phase */
  size_t launches;           /* kernels of this library launched by the last compute (cuBLAS GEMMs not counted) */
  size_t gemm_calls;         /* library (cuBLAS) GEMM calls of the last compute: always 0 since round 2, every product
                                runs on this library's own DMMA kernels; the field is kept for the bench line */
  size_t kgemm_blocks;       /* (batch, 64-column block)s of G that were not identically zero, i.e. executed by k_cosx_kgemm */
} occ_cosx_info_t;
int occ_cosx_get_info(const occ_cosx_handle *h, occ_cosx_info_t *info);

void occ_cosx_destroy(occ_cosx_handle *h);

/* ---- several GPUs of one box, driven from ONE host process (the layout of occ_xc_multi_*, include/occ_xc_b200.h) ----
 * SemiNumericalExchange::compute_K is called from the reference's single-process SCF (HartreeFock::compute_K,
 * src/qm/hf.cpp:179-258); its tbb loops over batches (:660-815) become: one handle per device with the grid batches dealt
 * round-robin, one host thread per device for the call, every GPU uploads D (and C) over its own PCIe link, and GPU r
 * sums slice r of the symmetrised partial K matrices through peer loads in rank order and writes it to the host matrix.
 * Needs peer access between the devices; an ordinal may repeat (several shards on one GPU).
 *   occ_cosx_multi_create     as occ_cosx_create (options.device / rank / world_size are overwritten per handle);
 *                             devices NULL = ordinals 0..ndev-1.  With batches sharded the projector has to be built from
 *                             the SUMMED numerical overlap: pass it here, or NULL and set it later from
 *                             occ_cosx_multi_numerical_overlap / _analytic_overlap (Sn.ldlt().solve(S)).
 *   occ_cosx_multi_compute_K  compute_K / compute_K_mo: kind 0 restricted, 1 unrestricted, 2 general (D form only);
 *                             C NULL = D form, else the occupied orbitals as in occ_cosx_compute_K_mo; omega > 0 = long range
 *   ..._range_separated       (alpha + beta) K(1/r) - beta K(erf(omega r)/r) in one sweep, kinds 0 and 1 */
typedef struct occ_cosx_multi occ_cosx_multi;
int occ_cosx_multi_create(const occ_xc_basis_t *basis, int natoms, const size_t *atom_offsets, const double *pts_xyz,
                          const double *weights, const double *overlap_projector, const occ_cosx_options_t *options,
                          int ndev, const int *devices, occ_cosx_multi **out);
int occ_cosx_multi_compute_K(occ_cosx_multi *m, int kind, const double *D, const double *C, int nocc_alpha, int nocc_beta,
                             double precision, double omega, double *K);
int occ_cosx_multi_compute_K_range_separated(occ_cosx_multi *m, int kind, const double *D, const double *C, int nocc_alpha,
                                             int nocc_beta, double alpha, double beta, double omega, double precision,
                                             double *K);
int occ_cosx_multi_numerical_overlap(occ_cosx_multi *m, double *S_numeric);
int occ_cosx_multi_analytic_overlap(occ_cosx_multi *m, double *S_analytic);
int occ_cosx_multi_set_projector(occ_cosx_multi *m, const double *overlap_projector);
int occ_cosx_multi_ndev(const occ_cosx_multi *m);
int occ_cosx_multi_get_info(const occ_cosx_multi *m, int rank, occ_cosx_info_t *info);
void occ_cosx_multi_destroy(occ_cosx_multi *m);

#ifdef __cplusplus
}
#endif
#endif
