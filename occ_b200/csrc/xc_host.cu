// xc_host.cu — host orchestration behind the C ABI (include/occ_xc_b200.h): what DFT::compute_K_dft
// (/root/reference include/occ/dft/dft.h:190-388) does around its leaf loop, re-designed for one B200:
//   create : shell screening per Morton leaf (dft.h:313-335), tiles, batches and the per-phase work lists —
//            all geometry-only, so built ONCE instead of every SCF iteration; everything uploaded to HBM
//   build  : D -> device, then per batch of leaves  collocate -> density (DMMA) -> functional -> G -> vxc (DMMA),
//            finally K = Ktmp + Ktmp^T and the fixed-order reduction of Exc / N_alpha / N_beta
// No CPU fallback: every phase is a CUDA kernel; create() fails without a device.
#include "../../include/occ_xc_b200.h"
#include "xc_functionals.cuh"
#include "xc_internal.h"
#include "xc_handle.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <numeric>
#include <nvtx3/nvToolsExt.h>
#include <string>
#include <vector>

using namespace xcb;

static thread_local std::string g_last_error;
static int fail(const std::string &msg) {
  g_last_error = msg;
  return 1;
}
int occ_xc_set_error(const char *msg) { return fail(msg); } // shared with grid.cu
#define CUDA_TRY(expr)                                                                                       \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return fail(std::string("CUDA error: ") + cudaGetErrorString(e__) + " in " #expr);                    \
  } while (0)

static inline int shell_size(int l, int pure) { return pure ? 2 * l + 1 : (l + 1) * (l + 2) / 2; }
static inline double common_fac_host(int l, int pure) { // gto.cpp:7-16
  if (l == 0) return 0.282094791773878143;
  if (l == 1) return 0.488602511902919921;
  return pure ? std::sqrt((2 * l + 1) / (4.0 * M_PI)) : 1.0;
}

// ---- real solid harmonics in Cartesian monomials (CCA order m = -l..l, Cartesian order lx desc then ly desc),
// the table gau2grid's GG_SPHERICAL_CCA transform applies (src/3rdparty/gau2grid/gau2grid_transform.c:44-196).
// Closed form: Helgaker, Jorgensen, Olsen, "Molecular Electronic-Structure Theory", eqs. 6.4.47-6.4.50. ----
static double binomial(int n, int k) {
  if (k < 0 || k > n) return 0.0;
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return r;
}
static double factorial(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}
static int cart_index(int l, int lx, int ly) {
  // position of x^lx y^ly z^(l-lx-ly) in the CCA ordering
  int idx = 0;
  for (int a = l; a > lx; --a) idx += l - a + 1;
  return idx + (l - lx - ly);
}
static void solid_harmonic_table(int l, double *M /* (2l+1) x ncart */) {
  const int nc = (l + 1) * (l + 2) / 2;
  std::fill(M, M + static_cast<size_t>(2 * l + 1) * nc, 0.0);
  for (int m = -l; m <= l; ++m) {
    const int am = std::abs(m);
    const double norm = std::sqrt(2.0 * factorial(l + am) * factorial(l - am) / (m == 0 ? 2.0 : 1.0)) /
                        (std::pow(2.0, am) * factorial(l));
    const int twovm = m < 0 ? 1 : 0; // 2 v_m
    for (int t = 0; 2 * t <= l - am; ++t)
      for (int u = 0; u <= t; ++u)
        for (int k = 0; 2 * k + twovm <= am; ++k) { // v = v_m + k
          const int twov = 2 * k + twovm;
          const double c = ((t + k) % 2 ? -1.0 : 1.0) * std::pow(0.25, t) * binomial(l, t) *
                           binomial(l - t, am + t) * binomial(t, u) * binomial(am, twov);
          const int lx = 2 * t + am - 2 * u - twov, ly = 2 * u + twov, lz = l - 2 * t - am;
          if (lx < 0 || ly < 0 || lz < 0) continue;
          M[static_cast<size_t>(m + l) * nc + cart_index(l, lx, ly)] += norm * c;
        }
  }
}
static const C2STable &c2s_table() {
  static const C2STable table = [] { // magic static: built once, safe when several threads create handles
    C2STable t;
    int off = 0;
    for (int l = 0; l <= MAX_L; ++l) {
      t.off[l] = off;
      solid_harmonic_table(l, t.coef + off);
      off += (2 * l + 1) * (l + 1) * (l + 2) / 2;
    }
    t.off[MAX_L + 1] = off;
    return t;
  }();
  return table;
}

namespace xcb {
void upload_default_c2s() { upload_c2s_table(c2s_table()); }
const double *c2s_racah_table(int l) { return c2s_table().coef + c2s_table().off[l]; }
} // namespace xcb

// value of every function of one shell at a point relative to its centre (host; used for the decay extents)
static void shell_values_host(int l, int pure, int np, const double *coefs, const double *exps, double x, double y,
                              double z, double *out) {
  const double r2 = x * x + y * y + z * z;
  double S0 = 0.0;
  for (int k = 0; k < np; ++k) S0 += coefs[k] * std::exp(-exps[k] * r2);
  const int nc = (l + 1) * (l + 2) / 2;
  double cart[(MAX_L + 1) * (MAX_L + 2) / 2];
  int idx = 0;
  for (int lx = l; lx >= 0; --lx)
    for (int ly = l - lx; ly >= 0; --ly) cart[idx++] = S0 * std::pow(x, lx) * std::pow(y, ly) * std::pow(z, l - lx - ly);
  if (!pure || l == 0) {
    for (int i = 0; i < nc; ++i) out[i] = cart[i];
    return;
  }
  const C2STable &t = c2s_table();
  for (int m = 0; m < 2 * l + 1; ++m) {
    double v = 0.0;
    for (int i = 0; i < nc; ++i) v += t.coef[t.off[l] + m * nc + i] * cart[i];
    out[m] = v;
  }
}

// gto::evaluate_decay_cutoff (src/gto/gto.cpp:18-67): first probe radius on the 0.1-Bohr ladder along +x where
// every component of the shell is below the threshold.  The probe points are carried through the same
// "pts += origin ... pts -= origin" shuttle as the reference, shell after shell.
static int decay_cutoff_host(const occ_xc_basis_t *b, double threshold, double *extents) {
  const double max_radius = 30.0, increment = 0.1, min_radius = 0.5;
  const int npts = static_cast<int>((max_radius - min_radius) / increment) + 1;
  std::vector<double> px(npts), py(npts, 0.0), pz(npts, 0.0);
  const double step = (max_radius - min_radius) / (npts - 1);
  for (int i = 0; i < npts; ++i) px[i] = (i == npts - 1) ? max_radius : min_radius + i * step;
  // atom_to_shell order: shells grouped by atom, atoms ascending (shells are stored atom by atom)
  std::vector<int> order(b->nsh);
  std::iota(order.begin(), order.end(), 0);
  if (b->shell_atom)
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return b->shell_atom[a] < b->shell_atom[c]; });
  double vals[2 * MAX_L + 1 > (MAX_L + 1) * (MAX_L + 2) / 2 ? 2 * MAX_L + 1 : (MAX_L + 1) * (MAX_L + 2) / 2];
  for (int s : order) {
    const double *o = b->origin + 3 * s;
    const int l = b->l[s];
    if (l > MAX_L) return fail("shell angular momentum above the supported maximum");
    for (int i = 0; i < npts; ++i) {
      px[i] += o[0];
      py[i] += o[1];
      pz[i] += o[2];
    }
    const int size = shell_size(l, b->pure);
    const double fac = common_fac_host(l, b->pure);
    for (int i = 0; i < npts; ++i) {
      extents[s] = i * increment + min_radius;
      shell_values_host(l, b->pure, b->nprim[s], b->coefs + b->prim_off[s], b->exps + b->prim_off[s], px[i] - o[0],
                        py[i] - o[1], pz[i] - o[2], vals);
      double v = 0.0;
      for (int c = 0; c < size; ++c) v = std::max(v, std::fabs(vals[c] * fac));
      if (v < threshold) break;
    }
    for (int i = 0; i < npts; ++i) {
      px[i] -= o[0];
      py[i] -= o[1];
      pz[i] -= o[2];
    }
  }
  return 0;
}

// cut[r] .. cut[r+1] = the run of (non-empty, unscreened) leaves rank r owns: prefix sums of the predicted
// cost split at multiples of total/world, so every rank gets a Morton-contiguous, spatially compact shard
static std::vector<size_t> partition_by_cost(const std::vector<double> &costs, int world) {
  double total = 0.0;
  for (double c : costs) total += c;
  std::vector<size_t> cut(world + 1, costs.size());
  cut[0] = 0;
  double run = 0.0;
  int r = 1;
  for (size_t i = 0; i < costs.size() && r < world; ++i) {
    run += costs[i];
    while (r < world && run >= total * r / world) cut[r++] = i + 1;
  }
  return cut;
}

template <class T> static int dev_alloc(occ_xc_handle *h, T **ptr, size_t count) {
  void *p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  h->alloc_bytes += bytes;
  *ptr = static_cast<T *>(p);
  return 0;
}
template <class T> static int dev_upload(occ_xc_handle *h, const T **ptr, const std::vector<T> &v) {
  T *p = nullptr;
  if (dev_alloc(h, &p, v.size())) return 1;
  if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *ptr = p;
  return 0;
}

static int ensure_tile_src(occ_xc_handle *h) {
  if (h->d_tile_src) return 0;
  if (dev_alloc(h, &h->d_tile_src, h->h_tile_src.size())) return 1;
  CUDA_TRY(cudaMemcpy(h->d_tile_src, h->h_tile_src.data(), h->h_tile_src.size() * sizeof(size_t), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" {

const char *occ_xc_last_error(void) { return g_last_error.c_str(); }
int occ_xc_abi_version(void) { return OCC_XC_B200_ABI_VERSION; }
int occ_xc_family(int id) { return xc_family_of(id); }
int occ_xc_range_separation(int id, double *omega, double *alpha, double *beta) {
  if (!omega || !alpha || !beta) return fail("occ_xc_range_separation: null argument");
  if (xc_family_of(id) < 0) return fail("functional with id = " + std::to_string(id) + " not found");
  xc_cam_of(id, omega, alpha, beta);
  return 0;
}
int occ_xc_measure_fp64_tensor_peak(int device, double target_ms, double *tflops, double *measured_ms) {
  if (!tflops) return fail("occ_xc_measure_fp64_tensor_peak: null argument");
  if (device >= 0) CUDA_TRY(cudaSetDevice(device));
  if (measure_dmma_peak(target_ms > 0.0 ? target_ms : 50.0, tflops, measured_ms))
    return fail(std::string("occ_xc_measure_fp64_tensor_peak: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}
void occ_xc_options_init(occ_xc_options_t *o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->device = -1;
}
double occ_xc_exact_exchange(int id) { return xc_exx_of(id); }

int occ_xc_decay_cutoff(const occ_xc_basis_t *basis, double threshold, double *extents) {
  if (!basis || !extents) return fail("occ_xc_decay_cutoff: null argument");
  return decay_cutoff_host(basis, threshold, extents);
}

int occ_xc_partition_leaves(const occ_xc_basis_t *basis, size_t nleaves, const occ_xc_leaf_t *leaves,
                            const double *shell_extents, double screening_threshold, int world_size,
                            int *leaf_rank, int *leaf_nb) {
  if (!basis || !leaves || !leaf_rank) return fail("occ_xc_partition_leaves: null argument");
  if (world_size < 1) return fail("occ_xc_partition_leaves: world_size must be positive");
  const int nsh = basis->nsh, pure = basis->pure ? 1 : 0;
  // the reference's environment switch for its dense path (dft.h:274-275) is honoured here too
  const bool screen = screening_threshold > 0.0 && std::getenv("OCC_XC_NO_SCREEN") == nullptr;
  std::vector<double> extents(nsh, 0.0);
  if (screen) {
    if (shell_extents) std::copy(shell_extents, shell_extents + nsh, extents.begin());
    else if (decay_cutoff_host(basis, screening_threshold, extents.data())) return 1;
  }
  std::vector<double> costs;
  std::vector<size_t> src;
  for (size_t l = 0; l < nleaves; ++l) {
    leaf_rank[l] = -1;
    int nb = 0;
    if (leaves[l].count > 0)
      for (int s = 0; s < nsh; ++s) {
        const double dx = basis->origin[3 * s] - leaves[l].center[0], dy = basis->origin[3 * s + 1] - leaves[l].center[1],
                     dz = basis->origin[3 * s + 2] - leaves[l].center[2];
        if (!screen || std::sqrt(dx * dx + dy * dy + dz * dz) - extents[s] < leaves[l].radius)
          nb += shell_size(basis->l[s], pure);
      }
    if (leaf_nb) leaf_nb[l] = nb;
    if (nb > 0) {
      costs.push_back(static_cast<double>(leaves[l].count) * nb * nb);
      src.push_back(l);
    }
  }
  const std::vector<size_t> cut = partition_by_cost(costs, world_size);
  for (int r = 0; r < world_size; ++r)
    for (size_t i = cut[r]; i < cut[r + 1]; ++i) leaf_rank[src[i]] = r;
  return 0;
}

int occ_xc_eval_functional(int func_id, int polarized, size_t n, const double *rho, const double *sigma,
                           const double *tau, double *exc, double *vrho, double *vsigma, double *vtau) {
  const int rc = eval_functional_device(func_id, polarized, n, rho, sigma, tau, exc, vrho, vsigma, vtau);
  if (rc == 1) return fail("functional with id = " + std::to_string(func_id) + " not found");
  if (rc) return fail(std::string("CUDA error in occ_xc_eval_functional: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}

void occ_xc_destroy(occ_xc_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void *p : h->allocs) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

int occ_xc_create(const occ_xc_basis_t *basis, size_t npts, const double *pts_xyz, const double *weights,
                  size_t nleaves_in, const occ_xc_leaf_t *leaves_in, const double *shell_extents, int nfunc,
                  const int *func_ids, int nspin, double density_threshold, double screening_threshold,
                  const occ_xc_options_t *opt, occ_xc_handle **out) {
  if (!basis || !pts_xyz || !weights || !leaves_in || !func_ids || !out)
    return fail("occ_xc_create: null argument");
  if (nspin != 1 && nspin != 2) return fail("occ_xc_create: nspin must be 1 (restricted) or 2 (unrestricted)");
  if (nfunc < 1 || nfunc > MAX_FUNC) return fail("occ_xc_create: between 1 and 4 functional components are supported");
  int deriv = 0;
  for (int f = 0; f < nfunc; ++f) {
    const int fam = xc_family_of(func_ids[f]);
    if (fam < 0) return fail("functional with id = " + std::to_string(func_ids[f]) + " not found");
    deriv = std::max(deriv, fam); // DFT::density_derivative (src/dft/dft.cpp:11-17)
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("occ_xc_create: no CUDA device available (this library has no CPU fallback)");
  int device = (opt && opt->device >= 0) ? opt->device : -1;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= ndev) return fail("occ_xc_create: device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  const int world = (opt && opt->world_size > 1) ? opt->world_size : 1;
  const int rank = world > 1 ? opt->rank : 0;
  if (rank < 0 || rank >= world) return fail("occ_xc_create: rank outside [0, world_size)");
  // default 16 GiB of the 180 GB: fewer, longer launches (valinomycin 171.6 ms at 4 GiB, 170.0 at 16, 169.7 at 32)
  const size_t batch_bytes = (opt && opt->batch_bytes) ? opt->batch_bytes : (size_t(16384) << 20);
  const bool want_grad = opt && opt->gradient == 1; // handle can also run occ_xc_gradient
  if (opt)
    for (int r : opt->reserved)
      if (r != 0) return fail("occ_xc_create: options.reserved must be zero (was the struct built for ABI version 1?)");
  const bool want_mo = opt && opt->mo_form == 1;
  if (want_mo && (opt->nocc_alpha < 0 || opt->nocc_beta < 0 || (opt->nocc_alpha == 0 && (nspin == 1 || opt->nocc_beta == 0))))
    return fail("occ_xc_create: mo_form needs nocc_alpha > 0 (restricted) or nocc_alpha + nocc_beta > 0 (unrestricted)");
  if (want_grad && !basis->shell_atom) return fail("occ_xc_create: the gradient option needs basis->shell_atom");

  const int nsh = basis->nsh, pure = basis->pure ? 1 : 0;
  std::vector<int> first_bf(nsh);
  int nbf = 0;
  for (int s = 0; s < nsh; ++s) {
    if (basis->l[s] < 0 || basis->l[s] > MAX_L) return fail("occ_xc_create: shell angular momentum outside 0..6");
    if (want_grad && basis->l[s] > 4) return fail("occ_xc_create: the gradient option supports l <= 4");
    first_bf[s] = nbf;
    nbf += shell_size(basis->l[s], pure);
  }
  for (size_t l = 0; l < nleaves_in; ++l)
    if (leaves_in[l].offset + leaves_in[l].count > npts) return fail("occ_xc_create: leaf exceeds the point array");

  // ---- shell extents (dft.h:293-294) ----
  // the reference's environment switch for its dense path (dft.h:274-275) is honoured here too
  const bool screen = screening_threshold > 0.0 && std::getenv("OCC_XC_NO_SCREEN") == nullptr;
  std::vector<double> extents(nsh, 0.0);
  if (screen) {
    if (shell_extents) std::copy(shell_extents, shell_extents + nsh, extents.begin());
    else if (decay_cutoff_host(basis, screening_threshold, extents.data())) return 1;
  }

  // ---- significant shells per leaf (dft.h:313-335) ----
  struct HostLeaf {
    size_t src;
    int sh_off, nshl, nb;
    double cost;
  };
  std::vector<HostLeaf> hl;
  std::vector<int> shell_list, shell_col;
  for (size_t l = 0; l < nleaves_in; ++l) {
    const occ_xc_leaf_t &lf = leaves_in[l];
    if (lf.count == 0) continue;
    HostLeaf e;
    e.src = l;
    e.sh_off = static_cast<int>(shell_list.size());
    e.nb = 0;
    for (int s = 0; s < nsh; ++s) {
      bool keep = true;
      if (screen) {
        const double dx = basis->origin[3 * s] - lf.center[0], dy = basis->origin[3 * s + 1] - lf.center[1],
                     dz = basis->origin[3 * s + 2] - lf.center[2];
        const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
        keep = dist - extents[s] < lf.radius;
      }
      if (keep) {
        shell_list.push_back(s);
        shell_col.push_back(e.nb);
        e.nb += shell_size(basis->l[s], pure);
      }
    }
    e.nshl = static_cast<int>(shell_list.size()) - e.sh_off;
    if (e.nshl == 0) {
      continue; // dft.h:320-321
    }
    e.cost = static_cast<double>(lf.count) * e.nb * e.nb;
    hl.push_back(e);
  }
  // ---- shard: contiguous Morton runs of equal predicted cost (n * nb^2; with the factorised density the density half
  // of a leaf costs n * nb * min(nb, 2 nocc) instead of n * nb^2) ----
  size_t l_lo = 0, l_hi = hl.size();
  if (world > 1) {
    std::vector<double> costs(hl.size());
    const int mo_nocc = want_mo ? std::max(opt->nocc_alpha, nspin == 2 ? opt->nocc_beta : 0) : 0;
    for (size_t i = 0; i < hl.size(); ++i) {
      costs[i] = hl[i].cost;
      if (mo_nocc > 0 && hl[i].nb > 2 * mo_nocc) costs[i] *= 0.5 * (1.0 + 2.0 * mo_nocc / hl[i].nb);
    }
    std::vector<size_t> cut = partition_by_cost(costs, world);
    l_lo = cut[rank];
    l_hi = cut[rank + 1];
  }

  occ_xc_handle *h = new occ_xc_handle();
  h->device = device;
  h->nbf = nbf;
  h->nsh = nsh;
  h->nspin = nspin;
  h->deriv = deriv;
  h->ncomp = (deriv >= 1 || want_grad) ? 4 : 1;
  h->pure = pure;
  h->gradient = want_grad;
  if (want_grad)
    for (int s = 0; s < nsh; ++s) h->natoms = std::max(h->natoms, basis->shell_atom[s] + 1);
  h->nb_min = 1 << 30;
  h->nb_max = 0;
  // AO slabs per tile: phi (, grad phi); handles that also serve occ_xc_gradient add the G_s, H_s,d and S_s,de slabs.
  // The Vxc build itself needs no G slab: k_vxc_resident forms its G block in shared memory.
  const int nslab = h->ncomp + (want_grad ? nspin : 0) + (want_grad && deriv >= 1 ? 3 * nspin : 0) +
                    (want_grad && deriv >= 2 ? 6 * nspin : 0);
  std::vector<int> bf_atom;
  if (want_mo) {
    h->mo = true;
    h->nocc[0] = opt->nocc_alpha;
    h->nocc[1] = nspin == 2 ? opt->nocc_beta : 0; // 0 is a legal beta count (hydrogen atom): that spin runs no item
    if (h->nocc[0] > nbf || h->nocc[1] > nbf) {
      delete h;
      return fail("occ_xc_create: more occupied orbitals than basis functions");
    }
    h->noccp = (std::max(h->nocc[0], h->nocc[1]) + NB_PAD - 1) / NB_PAD * NB_PAD;
  }
  const size_t psi_doubles = h->mo ? static_cast<size_t>(nspin) * h->noccp * TILE_P : 0; // per tile

  // ---- tiles, padded points, per-leaf index lists ----
  std::vector<double> px, py, pz, pw;
  std::vector<int> my_shell_list, my_shell_col;
  for (size_t i = l_lo; i < l_hi; ++i) {
    const HostLeaf &e = hl[i];
    const occ_xc_leaf_t &lf = leaves_in[e.src];
    LeafDesc ld;
    ld.tile0 = static_cast<int>(h->h_tiles.size());
    ld.ntiles = static_cast<int>((lf.count + TILE_P - 1) / TILE_P);
    ld.nb = e.nb;
    ld.pad = 0;
    if (ld.ntiles > 2) {
      delete h;
      return fail("occ_xc_create: leaf with more than 256 points (max_leaf_size, spatial_grid_hierarchy.h:26-30)");
    }
    const int sh_off = static_cast<int>(my_shell_list.size());
    const int bf_off = static_cast<int>(h->h_bf_index.size());
    for (int k = 0; k < e.nshl; ++k) {
      const int s = shell_list[e.sh_off + k];
      my_shell_list.push_back(s);
      my_shell_col.push_back(shell_col[e.sh_off + k]);
      for (int c = 0; c < shell_size(basis->l[s], pure); ++c) {
        h->h_bf_index.push_back(first_bf[s] + c);
        bf_atom.push_back(basis->shell_atom ? basis->shell_atom[s] : 0);
      }
    }
    const int nbp = (e.nb + NB_PAD - 1) / NB_PAD * NB_PAD;
    for (int k = 0; k < ld.ntiles; ++k) {
      TileDesc t;
      t.leaf = static_cast<int>(h->h_leaves.size());
      t.pt0 = static_cast<int>(h->h_tiles.size()) * TILE_P;
      t.count = static_cast<int>(std::min<size_t>(TILE_P, lf.count - static_cast<size_t>(k) * TILE_P));
      t.nb = e.nb;
      t.nbp = nbp;
      t.bf_off = bf_off;
      t.sh_off = sh_off;
      t.nsh = e.nshl;
      t.scr_off = 0;
      h->h_tile_src.push_back(lf.offset + static_cast<size_t>(k) * TILE_P);
      for (int p = 0; p < TILE_P; ++p) {
        const bool valid = p < t.count;
        const size_t src = lf.offset + static_cast<size_t>(k) * TILE_P + (valid ? p : 0);
        px.push_back(pts_xyz[3 * src]);
        py.push_back(pts_xyz[3 * src + 1]);
        pz.push_back(pts_xyz[3 * src + 2]);
        pw.push_back(valid ? weights[src] : 0.0);
      }
      h->h_tiles.push_back(t);
    }
    h->h_leaves.push_back(ld);
    h->leaf_global.push_back(e.src);
    h->npts += lf.count;
    h->sum_n_nb2 += e.cost;
    h->nb_min = std::min(h->nb_min, e.nb);
    h->nb_max = std::max(h->nb_max, e.nb);
  }
  h->nleaves = static_cast<int>(h->h_leaves.size());
  h->ntiles = static_cast<int>(h->h_tiles.size());
  h->npts_total = npts;
  h->world = world;
  if (h->nleaves == 0) h->nb_min = 0;

  h->h_tile_psi_off.assign(std::max(h->ntiles, 1), 0);
  // ---- batches of consecutive leaves under the workspace budget; per-phase work lists ----
  std::vector<CollocItem> colloc;
  // factorised density where nb > xo * nocc (rho / grad rho; flop break-even 2.0) and nb > xo/2 * nocc (tau; break-even 1.0)
  const char *xo_env = std::getenv("OCC_XC_MO_CROSSOVER"); // experiment switch
  const double mo_xo = xo_env ? std::atof(xo_env) : 2.0;
  auto mo_rho = [&](int nb, int no) { return static_cast<double>(nb) > mo_xo * no; };
  auto mo_tau = [&](int nb, int no) { return static_cast<double>(nb) > 0.5 * mo_xo * no; };
  std::vector<DensItem> dens, gitems, gradit, densmo, psi, xpsi, gpsi_w[5], gx_w[5], gradd;
  std::vector<VxcItem> vxc;
  size_t max_scratch = 0;
  {
    int leaf = 0;
    while (leaf < h->nleaves) {
      Batch b;
      b.leaf0 = leaf;
      b.tile0 = h->h_leaves[leaf].tile0;
      size_t used = 0;
      int nl = 0, nt = 0;
      while (leaf + nl < h->nleaves) {
        const LeafDesc &ld = h->h_leaves[leaf + nl];
        const size_t need =
            static_cast<size_t>(ld.ntiles) * (static_cast<size_t>(nslab) * h->h_tiles[ld.tile0].nbp * TILE_P + psi_doubles);
        if (nl > 0 && (used + need) * sizeof(double) > batch_bytes) break;
        for (int k = 0; k < ld.ntiles; ++k) {
          h->h_tiles[ld.tile0 + k].scr_off = static_cast<long long>(used);
          used += static_cast<size_t>(nslab) * h->h_tiles[ld.tile0 + k].nbp * TILE_P;
          h->h_tile_psi_off[ld.tile0 + k] = static_cast<long long>(used);
          used += psi_doubles;
        }
        nt += ld.ntiles;
        ++nl;
      }
      b.nleaves = nl;
      b.ntiles = nt;
      max_scratch = std::max(max_scratch, used);
      // tiles of the batch, most expensive first (cost of one CTA of either contraction grows with nb)
      std::vector<int> order(nt);
      std::iota(order.begin(), order.end(), b.tile0);
      std::stable_sort(order.begin(), order.end(),
                       [&](int a, int c) { return h->h_tiles[a].nb > h->h_tiles[c].nb; });
      b.colloc_off = colloc.size();
      b.dens_off = dens.size();
      b.g_off = gitems.size();
      b.vxc_off = vxc.size();
      b.grad_off = gradit.size();
      b.densmo_off = densmo.size();
      b.psi_off = psi.size();
      b.xpsi_off = xpsi.size();
      b.gradd_off = gradd.size();
      for (int wh = 0; wh < 5; ++wh) {
        b.gpsi_off[wh] = gpsi_w[wh].size();
        b.gx_off[wh] = gx_w[wh].size();
      }
      for (int ti : order) {
        const TileDesc &t = h->h_tiles[ti];
        for (int s0 = 0; s0 < t.nsh; s0 += 16) {
          CollocItem ci;
          ci.tile = ti;
          ci.s0 = s0;
          ci.s1 = std::min(t.nsh, s0 + 16);
          ci.last = ci.s1 == t.nsh;
          colloc.push_back(ci);
        }
        const int njb = t.nbp / 64;
        const int nac = deriv >= 2 ? 4 : 1;
        for (int sp = 0; sp < nspin; ++sp)
          for (int ac = 0; ac < nac; ++ac)
            for (int jb = 0; jb < njb; ++jb) {
              if (nspin == 2) continue; // unrestricted: spin-fused density / Vxc items below
              DensItem di;
              di.tile = ti;
              di.jb = static_cast<short>(jb);
              di.spin = static_cast<signed char>(sp);
              di.acomp = static_cast<signed char>(ac);
              dens.push_back(di);
              VxcItem vi;
              vi.tile = ti;
              vi.jb = static_cast<short>(jb);
              vi.spin = static_cast<signed char>(sp);
              vi.acomp = static_cast<signed char>(ac);
              vxc.push_back(vi);
              { // executed flops of this item: k_vxc_resident skips row blocks below the diagonal of symmetric terms
                const bool symmetric = ac != 0 || deriv == 0;
                const int n0 = jb * 64, nib_all = (t.nb + 127) / 128;
                const int nib = symmetric ? std::min(nib_all, (n0 + 64 + 127) / 128) : nib_all;
                const int rows = std::min(t.nb, nib * 128), cols = std::min(64, t.nb - n0);
                if (cols > 0) h->vxc_executed_flops += 2.0 * t.count * rows * cols;
              }
            }
        if (nspin == 2) // one item = 32 basis-function columns x both spins (k_density / k_vxc_resident, it.spin == 2)
          for (int ac = 0; ac < nac; ++ac)
            for (int jb = 0; jb * 32 < t.nb; ++jb) {
              DensItem di;
              di.tile = ti;
              di.jb = static_cast<short>(jb);
              di.spin = 2;
              di.acomp = static_cast<signed char>(ac);
              dens.push_back(di);
              VxcItem vi;
              vi.tile = ti;
              vi.jb = static_cast<short>(jb);
              vi.spin = 2;
              vi.acomp = static_cast<signed char>(ac);
              vxc.push_back(vi);
              const bool symmetric = ac != 0 || deriv == 0;
              const int n0 = jb * 32, nib_all = (t.nb + 127) / 128;
              const int nib = symmetric ? std::min(nib_all, (n0 + 32 + 127) / 128) : nib_all;
              const int rows = std::min(t.nb, nib * 128), cols = std::min(32, t.nb - n0);
              h->vxc_executed_flops += 2.0 * 2.0 * t.count * rows * cols;
            }
        if (h->mo) {
          // per tile and spin: rho / grad rho in the factorised form when 4 n nb nocc < 2 n nb^2, tau when nocc < nb
          const int npsib = h->noccp / 64;
          // unrestricted: where BOTH spins take the D form for a term, one spin-fused item per 32 columns serves them
          bool fused_d[4] = {false, false, false, false};
          if (nspin == 2 && h->nocc[0] > 0 && h->nocc[1] > 0)
            for (int ac = 0; ac < nac; ++ac) {
              const bool d0 = ac == 0 ? !mo_rho(t.nb, h->nocc[0]) : !mo_tau(t.nb, h->nocc[0]);
              const bool d1 = ac == 0 ? !mo_rho(t.nb, h->nocc[1]) : !mo_tau(t.nb, h->nocc[1]);
              fused_d[ac] = d0 && d1;
              if (!fused_d[ac]) continue;
              for (int jb = 0; jb * 32 < t.nb; ++jb) {
                DensItem di;
                di.tile = ti;
                di.jb = static_cast<short>(jb);
                di.spin = 2;
                di.acomp = static_cast<signed char>(ac);
                densmo.push_back(di);
              }
              h->mo_executed_flops += 2.0 * 2.0 * t.count * t.nb * t.nb;
            }
          for (int sp = 0; sp < nspin; ++sp) {
            const int no = h->nocc[sp];
            if (no == 0) continue; // empty spin (e.g. the beta spin of a hydrogen atom): its accumulators stay zero
            auto push = [&](std::vector<DensItem> &list, int nblocks, int ac) {
              for (int jb = 0; jb < nblocks; ++jb) {
                DensItem di;
                di.tile = ti;
                di.jb = static_cast<short>(jb);
                di.spin = static_cast<signed char>(sp);
                di.acomp = static_cast<signed char>(ac);
                list.push_back(di);
              }
            };
            const double n_pts = t.count;
            if (mo_rho(t.nb, no)) {
              push(psi, npsib, 0);
              push(xpsi, njb, 0);
              h->mo_executed_flops += 4.0 * n_pts * t.nb * no;
            } else if (!fused_d[0]) {
              push(densmo, njb, 0);
              h->mo_executed_flops += 2.0 * n_pts * t.nb * t.nb;
            }
            for (int ac = 1; ac < nac; ++ac) {
              if (mo_tau(t.nb, no)) {
                push(psi, npsib, ac);
                h->mo_executed_flops += 2.0 * n_pts * t.nb * no;
              } else if (!fused_d[ac]) {
                push(densmo, njb, ac);
                h->mo_executed_flops += 2.0 * n_pts * t.nb * t.nb;
              }
            }
          }
        }
        if (want_grad) {
          // acomp: 0 A = phi (companions H), 1 A = G (companions grad phi), 2..4 A = phi_e (companions S), see k_gemm.cu
          const int w0 = deriv == 0 ? 1 : 0, w1 = deriv == 0 ? 1 : (deriv == 1 ? 1 : 4);
          for (int sp = 0; sp < nspin; ++sp)
            for (int wh = w0; wh <= w1; ++wh)
              for (int jb = 0; jb < njb; ++jb) {
                DensItem gi;
                gi.tile = ti;
                gi.jb = static_cast<short>(jb);
                gi.spin = static_cast<signed char>(sp);
                gi.acomp = static_cast<signed char>(wh);
                gradit.push_back(gi);
                if (h->mo && h->nocc[sp] > 0 && !mo_rho(t.nb, h->nocc[sp])) gradd.push_back(gi);
              }
          if (h->mo)
            for (int sp = 0; sp < nspin; ++sp) {
              if (h->nocc[sp] == 0 || !mo_rho(t.nb, h->nocc[sp])) continue; // an empty spin has D_s = 0: no gradient term
              for (int wh = w0; wh <= w1; ++wh) {
                DensItem gi;
                gi.tile = ti;
                gi.spin = static_cast<signed char>(sp);
                gi.acomp = static_cast<signed char>(wh);
                if (wh != 0) // A = phi: the density phase left psi = phi C in the tile's psi slab
                  for (int jb = 0; jb < h->noccp / 64; ++jb) {
                    gi.jb = static_cast<short>(jb);
                    gpsi_w[wh].push_back(gi);
                  }
                for (int jb = 0; jb < njb; ++jb) {
                  gi.jb = static_cast<short>(jb);
                  gx_w[wh].push_back(gi);
                }
              }
            }
        }
        for (int jb = 0; jb < njb; ++jb) {
          DensItem gi;
          gi.tile = ti;
          gi.jb = static_cast<short>(jb);
          gi.spin = 0;
          gi.acomp = 0;
          gitems.push_back(gi);
        }
      }
      b.ncolloc = colloc.size() - b.colloc_off;
      b.ndens = dens.size() - b.dens_off;
      b.ng = gitems.size() - b.g_off;
      b.nvxc = vxc.size() - b.vxc_off;
      b.ngrad = gradit.size() - b.grad_off;
      b.ngradd = gradd.size() - b.gradd_off;
      for (int wh = 0; wh < 5; ++wh) {
        b.ngpsi[wh] = gpsi_w[wh].size() - b.gpsi_off[wh];
        b.ngx[wh] = gx_w[wh].size() - b.gx_off[wh];
      }
      b.ndensmo = densmo.size() - b.densmo_off;
      b.npsi = psi.size() - b.psi_off;
      b.nxpsi = xpsi.size() - b.xpsi_off;
      h->batches.push_back(b);
      leaf += nl;
    }
  }

  // ---- upload ----
  BuildParams &P = h->P;
  std::memset(&P, 0, sizeof(P));
  {
    std::vector<int> vl(basis->l, basis->l + nsh), vnp(basis->nprim, basis->nprim + nsh),
        vpo(basis->prim_off, basis->prim_off + nsh);
    size_t nprim_total = 0;
    for (int s = 0; s < nsh; ++s) nprim_total = std::max<size_t>(nprim_total, size_t(basis->prim_off[s]) + basis->nprim[s]);
    std::vector<double> vo(basis->origin, basis->origin + 3 * nsh), ve(basis->exps, basis->exps + nprim_total),
        vc(basis->coefs, basis->coefs + nprim_total);
    int rc = 0;
    rc |= dev_upload(h, &P.basis.l, vl);
    rc |= dev_upload(h, &P.basis.nprim, vnp);
    rc |= dev_upload(h, &P.basis.prim_off, vpo);
    rc |= dev_upload(h, &P.basis.first_bf, first_bf);
    rc |= dev_upload(h, &P.basis.origin, vo);
    rc |= dev_upload(h, &P.basis.exps, ve);
    rc |= dev_upload(h, &P.basis.coefs, vc);
    P.basis.pure = pure;
    rc |= dev_upload(h, &P.tiles, h->h_tiles);
    rc |= dev_upload(h, &P.leaves, h->h_leaves);
    rc |= dev_upload(h, &P.shell_list, my_shell_list);
    rc |= dev_upload(h, &P.shell_col, my_shell_col);
    rc |= dev_upload(h, &P.bf_index, h->h_bf_index);
    rc |= dev_upload(h, &P.bf_atom, bf_atom);
    rc |= dev_upload(h, &P.px, px);
    rc |= dev_upload(h, &P.py, py);
    rc |= dev_upload(h, &P.pz, pz);
    rc |= dev_upload(h, &P.pw, pw);
    const CollocItem *dc = nullptr;
    const DensItem *dd = nullptr, *dg = nullptr, *dgr = nullptr;
    const VxcItem *dv = nullptr;
    rc |= dev_upload(h, &dc, colloc);
    rc |= dev_upload(h, &dd, dens);
    rc |= dev_upload(h, &dg, gitems);
    rc |= dev_upload(h, &dv, vxc);
    rc |= dev_upload(h, &dgr, gradit);
    rc |= dev_upload(h, &P.tile_psi_off, h->h_tile_psi_off);
    if (h->mo) {
      const DensItem *d1 = nullptr, *d2 = nullptr, *d3 = nullptr;
      rc |= dev_upload(h, &d1, densmo);
      rc |= dev_upload(h, &d2, psi);
      rc |= dev_upload(h, &d3, xpsi);
      h->d_densmo = const_cast<DensItem *>(d1);
      h->d_psi = const_cast<DensItem *>(d2);
      h->d_xpsi = const_cast<DensItem *>(d3);
      if (want_grad) {
        const DensItem *dq = nullptr;
        rc |= dev_upload(h, &dq, gradd);
        h->d_gradd = const_cast<DensItem *>(dq);
        for (int wh = 0; wh < 5; ++wh) {
          const DensItem *d4 = nullptr, *d5 = nullptr;
          rc |= dev_upload(h, &d4, gpsi_w[wh]);
          rc |= dev_upload(h, &d5, gx_w[wh]);
          h->d_gpsi[wh] = const_cast<DensItem *>(d4);
          h->d_gx[wh] = const_cast<DensItem *>(d5);
        }
      }
      rc |= dev_alloc(h, &h->d_C, static_cast<size_t>(nbf) * (h->nocc[0] + h->nocc[1]));
      rc |= dev_alloc(h, &h->d_CT, static_cast<size_t>(nspin) * h->noccp * nbf);
    }
    h->d_gradit = const_cast<DensItem *>(dgr);
    h->d_colloc = const_cast<CollocItem *>(dc);
    h->d_dens = const_cast<DensItem *>(dd);
    h->d_g = const_cast<DensItem *>(dg);
    h->d_vxc = const_cast<VxcItem *>(dv);
    const size_t nt = std::max(h->ntiles, 1);
    const size_t ksz = static_cast<size_t>(nspin) * nbf * nbf;
    rc |= dev_alloc(h, &P.scratch, max_scratch);
    // second lane is opt-in (options.lanes == 2): measured on B200 it buys nothing, the block scheduler
    // does not co-schedule the second lane's kernels while the first lane fills the SMs (profiles/r01_experiments.md)
    h->lanes = (opt && opt->lanes == 2 && h->batches.size() >= 2) ? 2 : 1;
    if (h->lanes == 2) rc |= dev_alloc(h, &h->scratch2, max_scratch);
    h->deterministic = opt && opt->deterministic == 1;
    // experiment switch: restricted GGA builds (every Vxc item is a plain (tile, 64 columns) item) through k_vxc_ws
    h->vxc_ws = std::getenv("OCC_XC_VXC_WS") != nullptr && nspin == 1 && deriv == 1 && !h->deterministic;
    h->nvtx = opt && opt->nvtx == 1;
    h->density_scale = (opt && opt->density_scale != 0.0) ? opt->density_scale : 1.0;
    h->orbital_scale = (opt && opt->orbital_scale != 0.0) ? opt->orbital_scale : 1.0;
    const size_t det = h->deterministic ? 2 : 1; // low bins right behind the accumulators
    P.det_acc_off = static_cast<long long>(nt * nspin * NACC * TILE_P);
    P.det_K_off = static_cast<long long>(ksz);
    rc |= dev_alloc(h, &P.acc, det * nt * nspin * NACC * TILE_P);
    rc |= dev_alloc(h, &P.coef, nt * nspin * NCOEF * TILE_P);
    rc |= dev_alloc(h, &P.leaf_skip, std::max(h->nleaves, 1));
    rc |= dev_alloc(h, &P.leaf_sums, 3 * static_cast<size_t>(std::max(h->nleaves, 1)));
    rc |= dev_alloc(h, &h->d_D, ksz);
    rc |= dev_alloc(h, &h->d_Dsym, nspin == 1 ? ksz : 1);
    rc |= dev_alloc(h, &h->d_Ktmp, det * ksz);
    rc |= dev_alloc(h, &h->d_K, ksz);
    rc |= dev_alloc(h, &h->d_scalars, 4);
    if (want_grad) {
      rc |= dev_alloc(h, &h->d_DT, ksz);
      rc |= dev_alloc(h, &h->d_grad, 3 * static_cast<size_t>(std::max(h->natoms, 1)));
    }
    if (rc) {
      occ_xc_destroy(h);
      return 1;
    }
  }
  upload_c2s_table(c2s_table());
  P.ldD = nspin * nbf;
  P.nbf = nbf;
  P.nspin = nspin;
  P.deriv = deriv;
  P.ncomp = h->ncomp;
  P.nfunc = nfunc;
  for (int f = 0; f < nfunc; ++f) P.func_ids[f] = func_ids[f];
  P.density_threshold = density_threshold;
  P.slab_G = h->ncomp;
  P.slab_H = h->ncomp + nspin;
  P.slab_S = P.slab_H + 3 * nspin;
  P.grad = h->d_grad;
  P.grad_factor = deriv == 0 ? -4.0 : -2.0;
  P.debug = opt ? static_cast<int>(opt->debug) : 0;
  P.ldC = nbf;
  P.ldCT = h->noccp;
  P.mo_noccp = h->noccp;
  for (int sp = 0; sp < 2; ++sp) {
    P.mo_nocc[sp] = h->nocc[sp];
    P.moC[sp] = h->d_C ? h->d_C + static_cast<size_t>(sp) * nbf * h->nocc[0] : nullptr;
    P.moCT[sp] = h->d_CT ? h->d_CT + static_cast<size_t>(sp) * h->noccp * nbf : nullptr;
  }
  for (int s = 0; s < nspin; ++s) P.Ktmp[s] = h->d_Ktmp + static_cast<size_t>(s) * nbf;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    occ_xc_destroy(h);
    return fail("occ_xc_create: cudaStreamCreate failed");
  }
  h->launches = 0;
  for (const Batch &b : h->batches)
    h->launches += (b.ncolloc > 0) + (b.ndens > 0) + (b.nleaves > 0) + (b.nvxc > 0);
  h->launches += 2 + (nspin == 1 ? 1 : 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    occ_xc_destroy(h);
    return fail(std::string("occ_xc_create: ") + cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}



int occ_xc_build_device(occ_xc_handle *h, const double *D2_dev, double *K_dev, double *scalars_dev,
                        void *cuda_stream) {
  if (!h || !D2_dev || !K_dev || !scalars_dev) return fail("occ_xc_build_device: null argument");
  return xc_build_device_impl(h, D2_dev, nullptr, K_dev, scalars_dev, cuda_stream);
}

int occ_xc_build_mo_device(occ_xc_handle *h, const double *D2_dev, const double *C_dev, double *K_dev,
                           double *scalars_dev, void *cuda_stream) {
  if (!h || !C_dev || !K_dev || !scalars_dev) return fail("occ_xc_build_mo_device: null argument");
  if (!h->mo) return fail("occ_xc_build_mo_device: the handle was created without occupied-orbital counts");
  return xc_build_device_impl(h, D2_dev, C_dev, K_dev, scalars_dev, cuda_stream);
}

} // extern "C"

int xc_build_device_impl(occ_xc_handle *h, const double *D2_dev, const double *C_dev, double *K_dev, double *scalars_dev,
                         void *cuda_stream) {
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  BuildParams P = h->P;
  const size_t ksz = static_cast<size_t>(h->nspin) * h->nbf * h->nbf;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  const bool prof = h->profiling;
  const bool det = h->deterministic, nvtx = h->nvtx;
  float ms = 0.f;
  if (prof) {
    CUDA_TRY(cudaEventCreate(&ev[0]));
    CUDA_TRY(cudaEventCreate(&ev[1]));
    for (double &x : h->phase_ms) x = 0.0;
  }
  // phase brackets: CUDA events when profiling is on, NVTX ranges when options.nvtx is set
#define PHASE_BEGIN(name) do { if (nvtx) nvtxRangePushA(name); if (prof) cudaEventRecord(ev[0], st); } while (0)
#define PHASE_END(k) do { if (prof) { cudaEventRecord(ev[1], st); cudaEventSynchronize(ev[1]); \
    cudaEventElapsedTime(&ms, ev[0], ev[1]); h->phase_ms[k] += ms; } if (nvtx) nvtxRangePop(); } while (0)

  if (nvtx) nvtxRangePushA("occ_xc_build");
  PHASE_BEGIN("prepare_D");
  const bool mo = C_dev != nullptr;
  if (mo) { // C_s^T once per build: the X = psi C^T contraction gathers whole rows of it
    for (int sp = 0; sp < h->nspin; ++sp) {
      const double *Cs = C_dev + static_cast<size_t>(sp) * h->nbf * h->nocc[0];
      P.moC[sp] = Cs;
      launch_transpose_rect(Cs, h->d_CT + static_cast<size_t>(sp) * h->noccp * h->nbf, h->nbf, h->nocc[sp], h->nbf,
                            h->noccp, st);
    }
  }
  if (!D2_dev) { // D2_s = C_s C_s^T formed here (occ_xc_build_mo with D2 = NULL): only C crossed PCIe
    if (!mo) return fail("occ_xc_build: no density matrix");
    for (int sp = 0; sp < h->nspin; ++sp) {
      double *Ds = h->d_D + static_cast<size_t>(sp) * h->nbf;
      if (h->nocc[sp] == 0) {
        CUDA_TRY(cudaMemset2DAsync(Ds, static_cast<size_t>(P.ldD) * sizeof(double), 0, h->nbf * sizeof(double), h->nbf, st));
        continue;
      }
      launch_dgemm(true, h->nbf, h->nbf, h->nocc[sp], 1.0, P.moC[sp], h->nbf, 0, P.moC[sp], h->nbf, 0, Ds, P.ldD, 0, 1, st);
    }
    D2_dev = h->d_D;
  }
  if (h->nspin == 1) {
    launch_prepare_density(D2_dev, h->d_Dsym, h->nbf, 1, st);
    P.Drho[0] = h->d_Dsym;
    P.Dtau[0] = D2_dev;
  } else {
    for (int s = 0; s < 2; ++s) P.Drho[s] = P.Dtau[s] = D2_dev + static_cast<size_t>(s) * h->nbf;
  }
  CUDA_TRY(cudaMemsetAsync(h->d_Ktmp, 0, (det ? 2 : 1) * ksz * sizeof(double), st));
  PHASE_END(0);
  // Two lanes: batches alternate between the caller's stream and an auxiliary stream, each lane with its own AO
  // workspace.  The HBM-bound phases of one lane (collocation, G slab) and the tails of its DMMA kernels then overlap
  // with the tensor-bound kernels of the other lane.  Profiling runs single-lane so phase times stay attributable.
  const int lanes = prof ? 1 : h->lanes;
  if (lanes == 2) {
    CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
  }
  cudaStream_t main_st = st;
  for (size_t bi = 0; bi < h->batches.size(); ++bi) {
    const Batch &b = h->batches[bi];
    const bool second = lanes == 2 && (bi & 1);
    st = second ? h->aux_stream : main_st;
    P.scratch = second ? h->scratch2 : h->P.scratch;
    PHASE_BEGIN("collocation");
    launch_collocate(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    PHASE_END(1);
    PHASE_BEGIN("density");
    double *acc_b = P.acc + static_cast<size_t>(b.tile0) * h->nspin * NACC * TILE_P;
    const size_t acc_n = static_cast<size_t>(b.ntiles) * h->nspin * NACC * TILE_P;
    CUDA_TRY(cudaMemsetAsync(acc_b, 0, acc_n * sizeof(double), st));
    if (det) CUDA_TRY(cudaMemsetAsync(acc_b + P.det_acc_off, 0, acc_n * sizeof(double), st));
    if (mo) {
      (det ? launch_mo_psi_det : launch_mo_psi)(P, h->d_psi + b.psi_off, static_cast<int>(b.npsi), st);
      (det ? launch_density_det : launch_density)(P, h->d_densmo + b.densmo_off, static_cast<int>(b.ndensmo), st);
      (det ? launch_mo_density_det : launch_mo_density)(P, h->d_xpsi + b.xpsi_off, static_cast<int>(b.nxpsi), st);
    } else {
      (det ? launch_density_det : launch_density)(P, h->d_dens + b.dens_off, static_cast<int>(b.ndens), st);
    }
    if (det) launch_add_inplace(acc_b, acc_b + P.det_acc_off, acc_n, st); // fold the low bins: one fixed-order add
    PHASE_END(2);
    PHASE_BEGIN("functional");
    launch_functional(P, b.leaf0, b.nleaves, st);
    PHASE_END(3);
    PHASE_BEGIN("vxc");
    if (h->vxc_ws) launch_vxc_ws(P, h->d_vxc + b.vxc_off, static_cast<int>(b.nvxc), st);
    else (det ? launch_vxc_resident_det : launch_vxc_resident)(P, h->d_vxc + b.vxc_off, static_cast<int>(b.nvxc), st);
    PHASE_END(5);
  }
  st = main_st;
  if (lanes == 2) {
    CUDA_TRY(cudaEventRecord(h->ev_join, h->aux_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join, 0));
  }
  PHASE_BEGIN("finalise");
  if (det) launch_add_inplace(h->d_Ktmp, h->d_Ktmp + P.det_K_off, ksz, st);
  launch_finalize(P, h->nleaves, K_dev, scalars_dev, st);
  PHASE_END(6);
  if (nvtx) nvtxRangePop();
#undef PHASE_BEGIN
#undef PHASE_END
  if (prof) {
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" {

// host D (and C) -> device copies of the handle, with options.density_scale / orbital_scale applied on the device
static int upload_inputs(occ_xc_handle *h, const double *D2, const double *C, const double **D_dev, const double **C_dev) {
  const size_t ksz = static_cast<size_t>(h->nspin) * h->nbf * h->nbf;
  *D_dev = nullptr;
  *C_dev = nullptr;
  if (D2) {
    CUDA_TRY(cudaMemcpyAsync(h->d_D, D2, ksz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->density_scale != 1.0) launch_scale(h->d_D, h->d_D, ksz, h->density_scale, h->stream);
    *D_dev = h->d_D;
  }
  if (C) {
    const size_t csz = static_cast<size_t>(h->nbf) * (h->nocc[0] + h->nocc[1]);
    CUDA_TRY(cudaMemcpyAsync(h->d_C, C, csz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->orbital_scale != 1.0) launch_scale(h->d_C, h->d_C, csz, h->orbital_scale, h->stream);
    *C_dev = h->d_C;
  }
  return 0;
}
static int download_outputs(occ_xc_handle *h, double *K, double *exc, double *n_alpha, double *n_beta) {
  const size_t ksz = static_cast<size_t>(h->nspin) * h->nbf * h->nbf;
  CUDA_TRY(cudaMemcpyAsync(K, h->d_K, ksz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  double sc[3];
  CUDA_TRY(cudaMemcpyAsync(sc, h->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (exc) *exc = sc[0];
  if (n_alpha) *n_alpha = sc[1];
  if (n_beta) *n_beta = sc[2];
  return 0;
}

int occ_xc_build(occ_xc_handle *h, const double *D2, double *K, double *exc, double *n_alpha, double *n_beta) {
  if (!h || !D2 || !K) return fail("occ_xc_build: null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  const double *Dd, *Cd;
  if (upload_inputs(h, D2, nullptr, &Dd, &Cd)) return 1;
  if (xc_build_device_impl(h, Dd, nullptr, h->d_K, h->d_scalars, h->stream)) return 1;
  return download_outputs(h, K, exc, n_alpha, n_beta);
}

int occ_xc_density_on_grid(occ_xc_handle *h, const double *D2, double *rho) {
  if (!h || !D2 || !rho) return fail("occ_xc_density_on_grid: null argument");
  if (h->world != 1) return fail("occ_xc_density_on_grid: the handle must hold the whole grid (world_size 1)");
  CUDA_TRY(cudaSetDevice(h->device));
  if (ensure_tile_src(h)) return 1;
  cudaStream_t st = h->stream;
  BuildParams P = h->P;
  const size_t ksz = static_cast<size_t>(h->nspin) * h->nbf * h->nbf;
  const int ncol = h->deriv == 0 ? 1 : (h->deriv == 1 ? 4 : 6);
  const size_t nout = static_cast<size_t>(h->nspin) * h->npts_total * ncol;
  double *d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, nout * sizeof(double)));
  struct Free { // released on every exit path
    double *p;
    ~Free() { cudaFree(p); }
  } guard{d_out};
  CUDA_TRY(cudaMemsetAsync(d_out, 0, nout * sizeof(double), st));
  CUDA_TRY(cudaMemcpyAsync(h->d_D, D2, ksz * sizeof(double), cudaMemcpyHostToDevice, st));
  if (h->density_scale != 1.0) launch_scale(h->d_D, h->d_D, ksz, h->density_scale, st);
  if (h->nspin == 1) {
    launch_prepare_density(h->d_D, h->d_Dsym, h->nbf, 1, st);
    P.Drho[0] = h->d_Dsym;
    P.Dtau[0] = h->d_D;
  } else {
    for (int s = 0; s < 2; ++s) P.Drho[s] = P.Dtau[s] = h->d_D + static_cast<size_t>(s) * h->nbf;
  }
  // the accumulators of all tiles persist across batches, so each pass can be gathered once at its end
  auto density_pass = [&](bool laplacian) -> int {
    for (const Batch &b : h->batches) {
      if (laplacian) launch_collocate_laplacian(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
      else launch_collocate(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
      CUDA_TRY(cudaMemsetAsync(P.acc + static_cast<size_t>(b.tile0) * h->nspin * NACC * TILE_P, 0,
                               static_cast<size_t>(b.ntiles) * h->nspin * NACC * TILE_P * sizeof(double), st));
      launch_density(P, h->d_dens + b.dens_off, static_cast<int>(b.ndens), st);
    }
    return 0;
  };
  if (density_pass(false)) return 1;
  launch_gather_rho(P, h->d_tile_src, h->npts_total, h->ntiles, d_out, st);
  if (h->deriv >= 2) { // laplacian column (density.cpp:84-100): a second pass with A = lap phi, row-dotted with phi
    if (density_pass(true)) return 1;
    launch_gather_laplacian(P, h->d_tile_src, h->npts_total, h->ntiles, d_out, st);
  }
  CUDA_TRY(cudaMemcpyAsync(rho, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

} // extern "C"

// ---- post-SCF VV10 in stages (shared by the single-handle entries below and occ_xc_multi_vv10* in xc_multi.cu) ----
// first use: per-point arrays in the reference's order + this shard's slice of the triangular work list
int vv10_setup(occ_xc_handle *h, int part, int nparts) {
  if (ensure_tile_src(h)) return 1;
  if (h->vv.X) return 0;
  const size_t n = h->npts_total;
  double *buf = nullptr;
  if (dev_alloc(h, &buf, 9 * n) || dev_alloc(h, &h->d_vv_orig, n)) return 1;
  h->vv.n = n;
  h->vv.tile_src = h->d_tile_src;
  h->vv.X = buf;
  h->vv.Y = buf + n;
  h->vv.Z = buf + 2 * n;
  h->vv.rw2 = buf + 3 * n;
  h->vv.om = buf + 4 * n;
  h->vv.kap = buf + 5 * n;
  h->vv.F = buf + 6 * n;
  h->vv.U = buf + 7 * n;
  h->vv.W = buf + 8 * n;
  std::vector<int2> items;
  const int nblk = static_cast<int>((n + TILE_P - 1) / TILE_P), jc = vv10_jchunk();
  size_t k = 0;
  for (int bi = nblk - 1; bi >= 0; --bi) // longest rows first; dealt round-robin to the shards of a multi-GPU call
    for (int c = 0; c * jc <= bi; ++c, ++k)
      if (static_cast<int>(k % nparts) == part) items.push_back(make_int2(bi, c));
  h->n_vv_items = static_cast<int>(items.size());
  if (dev_alloc(h, &h->d_vv_items, items.size())) return 1;
  CUDA_TRY(cudaMemcpy(h->d_vv_items, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice));
  return 0;
}

// gradient extras: per-point force accumulators, point -> atom map, block boxes, the full-square work list
int vv10_setup_gradient(occ_xc_handle *h, int part, int nparts, const size_t *atom_offsets) {
  const size_t n = h->npts_total;
  const int nblk = static_cast<int>((n + TILE_P - 1) / TILE_P), jc = vv10_jchunk();
  if (!h->vv.GX) {
    double *buf = nullptr;
    if (dev_alloc(h, &buf, 3 * n) || dev_alloc(h, &h->vv.pt_atom, n) || dev_alloc(h, &h->vv.bbox, 6 * static_cast<size_t>(nblk)) ||
        dev_alloc(h, &h->d_vv_atom_off, static_cast<size_t>(h->natoms) + 1))
      return 1;
    h->vv.GX = buf;
    h->vv.GY = buf + n;
    h->vv.GZ = buf + 2 * n;
    std::vector<int2> items;
    size_t k = 0;
    for (int bi = 0; bi < nblk; ++bi)
      for (int c = 0; c * jc < nblk; ++c, ++k)
        if (static_cast<int>(k % nparts) == part) items.push_back(make_int2(bi, c));
    h->n_vv_grid_items = static_cast<int>(items.size());
    if (dev_alloc(h, &h->d_vv_grid_items, items.size())) return 1;
    CUDA_TRY(cudaMemcpy(h->d_vv_grid_items, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  if (atom_offsets[0] != 0 || atom_offsets[h->natoms] != n)
    return fail("occ_xc_vv10_gradient: atom_offsets must run from 0 to the number of NLC grid points");
  CUDA_TRY(cudaMemcpyAsync(h->d_vv_atom_off, atom_offsets, (static_cast<size_t>(h->natoms) + 1) * sizeof(size_t),
                           cudaMemcpyHostToDevice, h->stream));
  launch_vv10_point_atoms(h->vv, h->natoms, h->d_vv_atom_off, h->stream);
  return 0;
}

// D2 (host) -> device, rho / grad rho on this handle's tiles, per-point arrays of ITS points into h->vv (reference order)
int vv10_stage_density(occ_xc_handle *h, const double *D2, const size_t *orig_index, double vv10_b, double vv10_C,
                       double unowned, bool with_transpose) {
  cudaStream_t st = h->stream;
  BuildParams &P = h->vvP;
  P = h->P;
  const size_t n = h->npts_total;
  const size_t ksz = static_cast<size_t>(h->nbf) * h->nbf;
  VV10Params &V = h->vv;
  if (orig_index) {
    CUDA_TRY(cudaMemcpyAsync(h->d_vv_orig, orig_index, n * sizeof(size_t), cudaMemcpyHostToDevice, st));
    V.orig = h->d_vv_orig;
  } else {
    V.orig = nullptr;
  }
  const double pi = 3.14159265358979323846;
  V.kappa_factor = vv10_b * 1.5 * pi * std::pow(9.0 * pi, -1.0 / 6.0);
  V.beta = std::pow(3.0 / (vv10_b * vv10_b), 0.75) / 32.0;
  V.C = vv10_C;
  CUDA_TRY(cudaMemcpyAsync(h->d_D, D2, ksz * sizeof(double), cudaMemcpyHostToDevice, st));
  if (h->density_scale != 1.0) launch_scale(h->d_D, h->d_D, ksz, h->density_scale, st); // as every host entry point
  launch_prepare_density(h->d_D, h->d_Dsym, h->nbf, 1, st);
  P.Drho[0] = h->d_Dsym;
  P.Dtau[0] = h->d_D;
  if (with_transpose) {
    launch_transpose(h->d_D, h->d_DT, h->nbf, P.ldD, st);
    P.DT[0] = h->d_DT;
  }
  // pass 1: rho, grad rho on every tile (the accumulators persist across batches)
  for (const Batch &b : h->batches) {
    launch_collocate(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    CUDA_TRY(cudaMemsetAsync(P.acc + static_cast<size_t>(b.tile0) * NACC * TILE_P, 0,
                             static_cast<size_t>(b.ntiles) * NACC * TILE_P * sizeof(double), st));
    launch_density(P, h->d_dens + b.dens_off, static_cast<int>(b.ndens), st);
  }
  launch_vv10_prepare(P, V, h->ntiles, unowned, st);
  return 0;
}

// pair sums of this handle's work items: reads X..kap of `in`, accumulates into F, U, W of h->vv
int vv10_stage_pairs(occ_xc_handle *h, const xcb::VV10Params &in) {
  VV10Params V = in;
  V.F = h->vv.F;
  V.U = h->vv.U;
  V.W = h->vv.W;
  launch_vv10_pairs(V, h->d_vv_items, h->n_vv_items, h->stream);
  return 0;
}

// coefficients from the (summed) F, U, W of `sums`, then the Vxc contraction into h->d_K / h->d_scalars
int vv10_stage_vxc(occ_xc_handle *h, const xcb::VV10Params &sums) {
  cudaStream_t st = h->stream;
  const BuildParams &P = h->vvP;
  const size_t ksz = static_cast<size_t>(h->nbf) * h->nbf;
  VV10Params V = h->vv;
  V.F = sums.F;
  V.U = sums.U;
  V.W = sums.W;
  CUDA_TRY(cudaMemsetAsync(h->d_Ktmp, 0, ksz * sizeof(double), st));
  launch_vv10_coef(P, V, h->nleaves, st);
  // pass 2: Vxc contraction with the VV10 coefficients (AO values are recomputed: one batch of workspace)
  for (const Batch &b : h->batches) {
    if (h->batches.size() > 1) launch_collocate(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    launch_vxc_resident(P, h->d_vxc + b.vxc_off, static_cast<int>(b.nvxc), st);
  }
  launch_finalize(P, h->nleaves, h->d_K, h->d_scalars, st);
  return 0;
}

// coefficients, then the potential term of the gradient (the restricted GGA gradient matrices of
// nonlocal_correlation.cpp:462-512 with Dval = 2 mo.D = D2, i.e. HALF the weight of the XC gradient's distribution loop)
// and this handle's share of the grid-displacement term, into h->d_grad; energy into h->d_scalars[0]
int vv10_stage_gradient(occ_xc_handle *h, const xcb::VV10Params &merged, const xcb::VV10Params &sums) {
  cudaStream_t st = h->stream;
  BuildParams P = h->vvP;
  P.grad_factor = -1.0;
  VV10Params V = h->vv;
  V.F = sums.F;
  V.U = sums.U;
  V.W = sums.W;
  CUDA_TRY(cudaMemsetAsync(h->d_grad, 0, 3 * static_cast<size_t>(h->natoms) * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(h->vv.GX, 0, 3 * h->vv.n * sizeof(double), st));
  launch_vv10_coef(P, V, h->nleaves, st);
  for (const Batch &b : h->batches) {
    if (h->batches.size() > 1) launch_collocate(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    launch_gbuild(P, h->d_g + b.g_off, static_cast<int>(b.ng), st);
    launch_collocate_h(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    launch_grad_gemm(P, h->d_gradit + b.grad_off, static_cast<int>(b.ngrad), st);
  }
  launch_finalize(P, h->nleaves, h->d_K, h->d_scalars, st); // leaf energy sums -> d_scalars[0] (K is not used)
  VV10Params G = merged; // X..kap of the whole grid; force accumulators / boxes / atoms of this handle
  G.GX = h->vv.GX;
  G.GY = h->vv.GY;
  G.GZ = h->vv.GZ;
  G.pt_atom = h->vv.pt_atom;
  G.bbox = h->vv.bbox;
  launch_vv10_grid_pairs(G, h->d_vv_grid_items, h->n_vv_grid_items, st);
  launch_vv10_grid_reduce(G, h->d_grad, st);
  return 0;
}

static int vv10_check(occ_xc_handle *h, const char *who, bool gradient) {
  if (h->nspin != 1)
    return fail("Only restricted NLC functional implemented right now"); // nonlocal_correlation.cpp:370-372
  if (h->deriv != 1)
    return fail((std::string(who) + ": the handle must be created with a GGA functional list (rho and grad rho)").c_str());
  if (gradient && !h->gradient)
    return fail((std::string(who) + ": handle was created without the gradient option").c_str());
  return 0;
}
int vv10_check_handle(occ_xc_handle *h, const char *who, bool gradient) { return vv10_check(h, who, gradient); }

extern "C" {

int occ_xc_vv10(occ_xc_handle *h, const double *D2, const size_t *orig_index, double vv10_b, double vv10_C, double *K,
                double *energy) {
  if (!h || !D2 || !K || !energy) return fail("occ_xc_vv10: null argument");
  if (vv10_check(h, "occ_xc_vv10", false)) return 1;
  if (h->world != 1)
    return fail("occ_xc_vv10: the handle must hold the whole NLC grid (world_size 1; several GPUs: occ_xc_multi_vv10)");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const size_t ksz = static_cast<size_t>(h->nbf) * h->nbf;
  if (vv10_setup(h, 0, 1)) return 1;
  if (vv10_stage_density(h, D2, orig_index, vv10_b, vv10_C, -1.0, false)) return 1;
  if (vv10_stage_pairs(h, h->vv)) return 1;
  if (vv10_stage_vxc(h, h->vv)) return 1;
  CUDA_TRY(cudaMemcpyAsync(K, h->d_K, ksz * sizeof(double), cudaMemcpyDeviceToHost, st));
  double sc[3];
  CUDA_TRY(cudaMemcpyAsync(sc, h->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  *energy = sc[0];
  return 0;
}

int occ_xc_vv10_gradient(occ_xc_handle *h, const double *D2, const size_t *orig_index, const size_t *atom_offsets,
                         double vv10_b, double vv10_C, double *gradient, double *energy) {
  if (!h || !D2 || !atom_offsets || !gradient) return fail("occ_xc_vv10_gradient: null argument");
  if (vv10_check(h, "occ_xc_vv10_gradient", true)) return 1;
  if (h->world != 1)
    return fail("occ_xc_vv10_gradient: the handle must hold the whole NLC grid (world_size 1; several GPUs: "
                "occ_xc_multi_vv10_gradient)");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  if (vv10_setup(h, 0, 1) || vv10_setup_gradient(h, 0, 1, atom_offsets)) return 1;
  if (vv10_stage_density(h, D2, orig_index, vv10_b, vv10_C, -1.0, true)) return 1;
  if (vv10_stage_pairs(h, h->vv)) return 1;
  if (vv10_stage_gradient(h, h->vv, h->vv)) return 1;
  CUDA_TRY(cudaMemcpyAsync(gradient, h->d_grad, 3 * static_cast<size_t>(h->natoms) * sizeof(double),
                           cudaMemcpyDeviceToHost, st));
  double sc[3];
  CUDA_TRY(cudaMemcpyAsync(sc, h->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (energy) *energy = sc[0];
  return 0;
}

int occ_xc_build_mo(occ_xc_handle *h, const double *D2, const double *C, double *K, double *exc, double *n_alpha,
                    double *n_beta) {
  if (!h || !C || !K) return fail("occ_xc_build_mo: null argument");
  if (!h->mo) return fail("occ_xc_build_mo: the handle was created without occupied-orbital counts");
  CUDA_TRY(cudaSetDevice(h->device));
  const double *Dd, *Cd;
  if (upload_inputs(h, D2, C, &Dd, &Cd)) return 1; // D2 == NULL: formed on the device as C C^T
  if (xc_build_device_impl(h, Dd, Cd, h->d_K, h->d_scalars, h->stream)) return 1;
  return download_outputs(h, K, exc, n_alpha, n_beta);
}

int occ_xc_pinned_alloc(size_t bytes, void **out) {
  if (!out) return fail("occ_xc_pinned_alloc: null argument");
  CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return 0;
}
int occ_xc_pinned_free(void *p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return 0;
}
int occ_xc_host_register(void *p, size_t bytes) {
  if (!p) return fail("occ_xc_host_register: null argument");
  CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}
int occ_xc_host_unregister(void *p) {
  if (p) CUDA_TRY(cudaHostUnregister(p));
  return 0;
}


int occ_xc_gradient(occ_xc_handle *h, const double *D2, double *gradient) {
  if (!h || !D2 || !gradient) return fail("occ_xc_gradient: null argument");
  if (!h->gradient) return fail("occ_xc_gradient: handle was created without the gradient option");
  return xc_gradient_impl(h, D2, nullptr, gradient);
}

int occ_xc_gradient_mo(occ_xc_handle *h, const double *D2, const double *C, double *gradient) {
  if (!h || !D2 || !C || !gradient) return fail("occ_xc_gradient_mo: null argument");
  if (!h->gradient) return fail("occ_xc_gradient_mo: handle was created without the gradient option");
  if (!h->mo) return fail("occ_xc_gradient_mo: the handle was created without occupied-orbital counts");
  return xc_gradient_impl(h, D2, C, gradient);
}

} // extern "C"
int xc_gradient_impl(occ_xc_handle *h, const double *D2, const double *C, double *gradient) {
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  BuildParams P = h->P;
  const size_t ksz = static_cast<size_t>(h->nspin) * h->nbf * h->nbf;
  const bool mo = C != nullptr;
  CUDA_TRY(cudaMemcpyAsync(h->d_D, D2, ksz * sizeof(double), cudaMemcpyHostToDevice, st));
  if (h->density_scale != 1.0) launch_scale(h->d_D, h->d_D, ksz, h->density_scale, st); // as every host entry point
  if (mo) {
    const size_t csz = static_cast<size_t>(h->nbf) * (h->nocc[0] + h->nocc[1]);
    CUDA_TRY(cudaMemcpyAsync(h->d_C, C, csz * sizeof(double), cudaMemcpyHostToDevice, st));
    if (h->orbital_scale != 1.0) launch_scale(h->d_C, h->d_C, csz, h->orbital_scale, st);
    for (int sp = 0; sp < h->nspin; ++sp) {
      const double *Cs = h->d_C + static_cast<size_t>(sp) * h->nbf * h->nocc[0];
      P.moC[sp] = Cs;
      launch_transpose_rect(Cs, h->d_CT + static_cast<size_t>(sp) * h->noccp * h->nbf, h->nbf, h->nocc[sp], h->nbf,
                            h->noccp, st);
    }
  }
  if (h->nspin == 1) {
    launch_prepare_density(h->d_D, h->d_Dsym, h->nbf, 1, st);
    P.Drho[0] = h->d_Dsym;
    P.Dtau[0] = h->d_D;
  } else {
    for (int s = 0; s < 2; ++s) P.Drho[s] = P.Dtau[s] = h->d_D + static_cast<size_t>(s) * h->nbf;
  }
  for (int s = 0; s < h->nspin; ++s) {
    launch_transpose(h->d_D + static_cast<size_t>(s) * h->nbf, h->d_DT + static_cast<size_t>(s) * h->nbf, h->nbf, P.ldD, st);
    P.DT[s] = h->d_DT + static_cast<size_t>(s) * h->nbf;
  }
  CUDA_TRY(cudaMemsetAsync(h->d_grad, 0, 3 * static_cast<size_t>(h->natoms) * sizeof(double), st));
  for (const Batch &b : h->batches) {
    launch_collocate(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    CUDA_TRY(cudaMemsetAsync(P.acc + static_cast<size_t>(b.tile0) * h->nspin * NACC * TILE_P, 0,
                             static_cast<size_t>(b.ntiles) * h->nspin * NACC * TILE_P * sizeof(double), st));
    if (mo) {
      launch_mo_psi(P, h->d_psi + b.psi_off, static_cast<int>(b.npsi), st);
      launch_density(P, h->d_densmo + b.densmo_off, static_cast<int>(b.ndensmo), st);
      launch_mo_density(P, h->d_xpsi + b.xpsi_off, static_cast<int>(b.nxpsi), st);
    } else {
      launch_density(P, h->d_dens + b.dens_off, static_cast<int>(b.ndens), st);
    }
    launch_functional(P, b.leaf0, b.nleaves, st);
    launch_gbuild(P, h->d_g + b.g_off, static_cast<int>(b.ng), st);
    if (h->deriv >= 1) launch_collocate_h(P, h->d_colloc + b.colloc_off, static_cast<int>(b.ncolloc), st);
    if (mo) {
      launch_grad_gemm(P, h->d_gradd + b.gradd_off, static_cast<int>(b.ngradd), st);
      for (int wh = 0; wh < 5; ++wh) { // one psi slab per spin: the A operands take turns
        launch_mo_grad_psi(P, h->d_gpsi[wh] + b.gpsi_off[wh], static_cast<int>(b.ngpsi[wh]), st);
        launch_mo_grad_gemm(P, h->d_gx[wh] + b.gx_off[wh], static_cast<int>(b.ngx[wh]), st);
      }
    } else {
      launch_grad_gemm(P, h->d_gradit + b.grad_off, static_cast<int>(b.ngrad), st);
    }
  }
  CUDA_TRY(cudaMemcpyAsync(gradient, h->d_grad, 3 * static_cast<size_t>(h->natoms) * sizeof(double),
                           cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" {

int occ_xc_get_info(const occ_xc_handle *h, occ_xc_info_t *info) {
  if (!h || !info) return fail("occ_xc_get_info: null argument");
  info->nbf = h->nbf;
  info->nspin = h->nspin;
  info->deriv = h->deriv;
  info->nsh = h->nsh;
  info->npts = h->npts;
  info->nleaves = static_cast<size_t>(h->nleaves);
  info->ntiles = static_cast<size_t>(h->ntiles);
  info->nbatches = h->batches.size();
  info->sum_n_nb2 = h->sum_n_nb2;
  info->flops_per_build = (h->deriv >= 2 ? 16.0 : 4.0) * h->nspin * h->sum_n_nb2;
  info->nb_min = h->nb_min;
  info->nb_max = h->nb_max;
  info->workspace_bytes = h->alloc_bytes;
  info->launches_per_build = h->launches;
  info->mo_density_flops = h->mo_executed_flops;
  info->vxc_flops = h->vxc_executed_flops;
  size_t lm = 0;
  if (h->mo) {
    for (const Batch &b : h->batches)
      lm += (b.ncolloc > 0) + (b.npsi > 0) + (b.ndensmo > 0) + (b.nxpsi > 0) + (b.nleaves > 0) + (b.nvxc > 0);
    lm += 2 + (h->nspin == 1 ? 1 : 0) + h->nspin;
  }
  info->launches_per_build_mo = lm;
  return 0;
}

int occ_xc_set_profiling(occ_xc_handle *h, int on) {
  if (!h) return fail("occ_xc_set_profiling: null handle");
  h->profiling = on != 0;
  return 0;
}
int occ_xc_get_phase_ms(occ_xc_handle *h, double *ms) {
  if (!h || !ms) return fail("occ_xc_get_phase_ms: null argument");
  for (int k = 0; k < OCC_XC_NPHASES; ++k) ms[k] = h->phase_ms[k];
  return 0;
}

int occ_xc_leaf_collocation(occ_xc_handle *h, size_t leaf, int deriv, int *nb_out, double *phi, double *phi_x,
                            double *phi_y, double *phi_z, size_t capacity) {
  if (!h || !nb_out || !phi) return fail("occ_xc_leaf_collocation: null argument");
  int li = -1;
  for (size_t i = 0; i < h->leaf_global.size(); ++i)
    if (h->leaf_global[i] == leaf) li = static_cast<int>(i);
  if (li < 0) return fail("occ_xc_leaf_collocation: leaf is empty, screened out, or not on this shard");
  if (deriv > 0 && h->ncomp < 4) return fail("occ_xc_leaf_collocation: handle was created without AO gradients");
  if (deriv > 0 && (!phi_x || !phi_y || !phi_z)) return fail("occ_xc_leaf_collocation: null gradient buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  const LeafDesc &ld = h->h_leaves[li];
  size_t count = 0;
  for (int k = 0; k < ld.ntiles; ++k) count += h->h_tiles[ld.tile0 + k].count;
  *nb_out = ld.nb;
  if (capacity < count * ld.nb) return fail("occ_xc_leaf_collocation: output capacity too small");
  std::vector<CollocItem> items;
  for (int k = 0; k < ld.ntiles; ++k) {
    const TileDesc &t = h->h_tiles[ld.tile0 + k];
    for (int s0 = 0; s0 < t.nsh; s0 += 16)
      items.push_back({ld.tile0 + k, s0, std::min(t.nsh, s0 + 16), std::min(t.nsh, s0 + 16) == t.nsh});
  }
  CollocItem *d_items = nullptr;
  CUDA_TRY(cudaMalloc(&d_items, items.size() * sizeof(CollocItem)));
  CUDA_TRY(cudaMemcpy(d_items, items.data(), items.size() * sizeof(CollocItem), cudaMemcpyHostToDevice));
  launch_collocate(h->P, d_items, static_cast<int>(items.size()), h->stream);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  cudaFree(d_items);
  double *outs[4] = {phi, phi_x, phi_y, phi_z};
  std::vector<double> buf;
  size_t row0 = 0;
  for (int k = 0; k < ld.ntiles; ++k) {
    const TileDesc &t = h->h_tiles[ld.tile0 + k];
    const size_t slab = static_cast<size_t>(t.nbp) * TILE_P;
    buf.resize(slab);
    for (int c = 0; c < (deriv > 0 ? 4 : 1); ++c) {
      CUDA_TRY(cudaMemcpy(buf.data(), h->P.scratch + t.scr_off + c * slab, slab * sizeof(double),
                          cudaMemcpyDeviceToHost));
      for (int j = 0; j < t.nb; ++j)
        for (int p = 0; p < t.count; ++p) outs[c][static_cast<size_t>(j) * count + row0 + p] = buf[j * TILE_P + p];
    }
    row0 += t.count;
  }
  return 0;
}

} // extern "C"
