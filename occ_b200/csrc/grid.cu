// grid.cu — the once-per-geometry input producers of the XC build, so that the library
// can be driven stand-alone (bench, tests) with exactly the reference's quadrature:
//   * occ::numint::MolecularGrid (src/numint/molecular_grid.cpp:11-131,183-256): per-element LMG radial grids
//     (src/numint/grid_utils.cpp:398-514) with NWChem angular pruning (:234-274) on Lebedev shells (:737-749),
//     Becke fuzzy-cell weights with the Treutler-Ahlrichs radius adjustment (:125-135,149-200)
//   * occ::numint::SpatialGridHierarchy (src/numint/spatial_grid_hierarchy.cpp:40-147): 63-bit Morton order,
//     leaves of 128 points, centroid bounding spheres
// B200 mapping: the only O(N_grid * N_atoms^2) piece — the Becke partition — is a CUDA kernel (one thread per
// grid point, atom distances staged in shared memory, every cell function formed as a running product so no
// per-thread arrays are needed); radial solves, pruning and the Morton sort are small host work.
#include "../../include/occ_xc_b200.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>

int occ_xc_set_error(const char *msg); // xc_host.cu

namespace {

const int kLebedev[33] = {1,    6,    14,   26,   38,   50,   74,   86,   110,  146,  170,  194,  230,  266,  302,  350, 434,
                          590,  770,  974,  1202, 1454, 1730, 2030, 2354, 2702, 3074, 3470, 3890, 4334, 4802, 5294, 5810};

// Bragg-Slater radii (Angstrom), grid_utils.cpp:22-41
const double kBraggAngstrom[131] = {
    0.35, 1.40, 1.45, 1.05, 0.85, 0.70, 0.65, 0.60, 0.50, 1.50, 1.80, 1.50, 1.25, 1.10, 1.00, 1.00, 1.00, 1.80, 2.20,
    1.80, 1.60, 1.40, 1.35, 1.40, 1.40, 1.40, 1.35, 1.35, 1.35, 1.35, 1.30, 1.25, 1.15, 1.15, 1.15, 1.90, 2.35, 2.00,
    1.80, 1.55, 1.45, 1.45, 1.35, 1.30, 1.35, 1.40, 1.60, 1.55, 1.55, 1.45, 1.45, 1.40, 1.40, 2.10, 2.60, 2.15, 1.95,
    1.85, 1.85, 1.85, 1.85, 1.85, 1.85, 1.80, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.55, 1.45, 1.35, 1.35, 1.30,
    1.35, 1.35, 1.35, 1.50, 1.90, 1.80, 1.60, 1.90, 1.45, 2.10, 1.80, 2.15, 1.95, 1.80, 1.80, 1.75, 1.75, 1.75, 1.75,
    1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75,
    1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75};

double bragg_bohr(int Z) { // grid_utils.cpp:78-86, units.h:7-8
  const double a2b = 1.0 / 0.52917721067;
  return (Z >= 1 && Z <= 131 ? kBraggAngstrom[Z - 1] : 1.0) * a2b;
}
int lebedev_at_or_above(int n) { // grid_utils.cpp:43-53
  for (int v : kLebedev)
    if (v >= n) return v;
  return -1;
}
int lebedev_below(int n) { // grid_utils.cpp:55-69
  if (n <= kLebedev[0]) return kLebedev[0];
  for (int i = 32; i >= 0; --i)
    if (kLebedev[i] < n) return kLebedev[i];
  return kLebedev[0];
}

std::map<int, std::vector<double>> &lebedev_tables() {
  static std::map<int, std::vector<double>> t;
  return t;
}

// The reference locates r_outer and h by the same decimating walk (grid_utils.cpp:412-475): move by `step`
// in the direction given by (f > eps), shrink the step tenfold on every direction change, stop when the move
// is below 1e-14.  `up_when_above` is +1 for r_outer (f falls with r) and -1 for h (f rises with h).
template <class F> double decimating_walk(double start, double step, double eps, double up_when_above, F f) {
  double x = start, x_old = 1.0e50;
  double dir = up_when_above; // the reference seeds f = 1e50 > eps
  while (std::fabs(x_old - x) > 1e-14) {
    const double fx = f(x);
    const double prev = dir;
    dir = (fx > eps) ? up_when_above : -up_when_above;
    if (x < 0.0) dir = 1.0;
    if (dir != prev) step *= 0.1;
    x_old = x;
    x += dir * step;
  }
  return x;
}

struct Radial {
  std::vector<double> r, w;
};

// generate_lmg_radial_grid, grid_utils.cpp:478-514 (Lindh-Malmqvist-Gagliardi, TCA 106, 178 (2001))
Radial lmg_radial(int Z, double eps, double alpha_max, int l_max, const std::vector<double> &alpha_min) {
  // eq. 25 with m = 0
  double r_inner = 1.9 - std::log(1.0 / eps);
  r_inner = r_inner * 2.0 / 3.0;
  r_inner = std::sqrt(std::exp(r_inner) / (2 * alpha_max));
  double h = FLT_MAX, r_outer = 0.0;
  const double br = bragg_bohr(Z);
  for (int l = 0; l <= l_max; ++l) {
    if (!(alpha_min[l] > 0.0)) continue; // DBL_MAX entries (l the atom lacks) pass this test, as in the reference
    const int m = 2 * l;
    const double a = alpha_min[l];
    const double ro = decimating_walk(4.0 * br, 0.5, eps, 1.0, [&](double r) { // eq. 19
      return tgamma((m + 3.0) / 2.0) * std::pow(a * r * r, (m + 1.0) / 2.0) * std::exp(-a * r * r);
    });
    r_outer = std::max(r_outer, ro);
    const double guess = 0.1 * (r_outer - r_inner);
    const double hl = decimating_walk(guess, 0.1 * guess, eps, -1.0, [&](double x) { // eqs. 17, 18
      const double c0 = 4.0 * std::sqrt(2.0) * M_PI;
      const double cm = tgamma(3.0 / 2.0) / tgamma((m + 3.0) / 2.0);
      const double rd0 = c0 * (1.0 / x) * std::exp(-M_PI * M_PI / (2.0 * x));
      return cm * std::pow(M_PI / x, m / 2.0) * rd0;
    });
    h = std::min(h, hl);
  }
  const double c = r_inner / (std::exp(h) - 1.0);
  const size_t n = static_cast<int>(std::log(1.0 + (r_outer / c)) / h);
  Radial g;
  g.r.resize(n);
  g.w.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const double r = c * (std::exp((i + 1) * h) - 1.0);
    g.r[i] = r;
    g.w[i] = (r + c) * h * r * r;
  }
  return g;
}

// generate_treutler_alrichs_radial_grid (grid_utils.cpp:559-572) with the scaling of generate_atom_grid's COSX branch
// (:600-611): nodes and weights times the Bragg radius rm, then r^2 folded into the weights
Radial treutler_ahlrichs_radial(int n, double rm) {
  Radial g;
  g.r.resize(n);
  g.w.resize(n);
  const double step = M_PI / (n + 1), ln2 = 1 / std::log(2.0);
  for (int i = 1; i <= n; ++i) {
    const double x = std::cos(i * step);
    const double t1 = ln2 * std::pow(1 + x, 0.6), t2 = std::log((1 - x) / 2);
    const double r = (-t1 * t2) * rm;
    double w = step * std::sin(i * step) * t1 * (-0.6 / (1 + x) * t2 + 1 / (1 - x));
    w *= rm;
    w *= r * r;
    g.r[n - i] = r;
    g.w[n - i] = w;
  }
  return g;
}

// prune_nwchem_scheme, grid_utils.cpp:234-274: Lebedev order per radial shell
int pruned_order(int Z, int max_angular, double r) {
  if (max_angular < 50) return max_angular;
  int lv[5] = {5, 6, 6, 6, 5};
  if (max_angular != 50) {
    int i = 6;
    while (i < 33 && kLebedev[i] != max_angular) ++i;
    lv[2] = i - 1;
    lv[3] = i;
    lv[4] = i - 1;
  }
  const double a_h[4] = {0.25, 0.5, 1.0, 4.5}, a_2[4] = {0.16666667, 0.5, 0.9, 3.5}, a_x[4] = {0.1, 0.4, 0.8, 2.5};
  const double *al = Z <= 2 ? a_h : (Z <= 10 ? a_2 : a_x);
  const double scale = r / bragg_bohr(Z);
  int k = 0;
  while (k < 4 && !(al[k] > scale)) ++k;
  return kLebedev[lv[k]];
}

struct ElementGrid {
  std::vector<double> xyz, w; // centred at the origin
  int nradial = 0;
};

// ---- Becke partition on the device ----------------------------------------------------------------------
constexpr int BECKE_THREADS = 64;

// one thread per grid point of atom `owner`; gd[a][t] = |p_t - R_a| in shared memory.
// cell_k = prod_{j != k, ascending} s_kj with s_kj = (1 - v_kj)/2 for j < k, (1 + v_jk)/2 for j > k, where v_ij
// (i > j) is the thrice-iterated Becke polynomial of the TA-adjusted confocal coordinate — the same factors,
// in the same order, as calculate_atomic_grid_weights_impl (grid_utils.cpp:149-200) accumulates them.
// grid = (blocks of the largest atom grid, natoms): blockIdx.y is the owning atom, atom_off its first point
__global__ void k_becke(int natoms, const double *__restrict__ pos, const double *__restrict__ dist,
                        const double *__restrict__ aij, const size_t *__restrict__ atom_off,
                        const double *__restrict__ xyz_all, double *__restrict__ w_all) {
  extern __shared__ double gd[]; // [natoms][BECKE_THREADS]
  const int owner = blockIdx.y;
  const int npts = static_cast<int>(atom_off[owner + 1] - atom_off[owner]);
  if (static_cast<int>(blockIdx.x) * BECKE_THREADS >= npts) return;
  const double *xyz = xyz_all + 3 * atom_off[owner];
  double *w = w_all + atom_off[owner];
  const int t = threadIdx.x;
  const int p = blockIdx.x * BECKE_THREADS + t;
  const bool valid = p < npts;
  const double x = valid ? xyz[3 * p] : 0.0, y = valid ? xyz[3 * p + 1] : 0.0, z = valid ? xyz[3 * p + 2] : 0.0;
  for (int a = 0; a < natoms; ++a) {
    const double dx = x - pos[3 * a], dy = y - pos[3 * a + 1], dz = z - pos[3 * a + 2];
    gd[a * BECKE_THREADS + t] = sqrt(dx * dx + dy * dy + dz * dz);
  }
  double total = 0.0, mine = 0.0;
  for (int k = 0; k < natoms; ++k) {
    const double gk = gd[k * BECKE_THREADS + t];
    double cell = 1.0;
    for (int j = 0; j < natoms; ++j) {
      if (j == k) continue;
      const int hi = k > j ? k : j, lo = k > j ? j : k;
      const double ghi = k > j ? gk : gd[j * BECKE_THREADS + t], glo = k > j ? gd[j * BECKE_THREADS + t] : gk;
      double v = (ghi - glo) / dist[static_cast<size_t>(hi) * natoms + lo];
      const double a = aij[static_cast<size_t>(hi) * natoms + lo];
      if (a != 0.0) v += a * (1 - v * v);
      v = (3 - v * v) * v * 0.5;
      v = (3 - v * v) * v * 0.5;
      v = (3 - v * v) * v * 0.5;
      const bool saturated = fabs(1.0 - v) < 1e-14;
      if (k > j) cell = saturated ? 0.0 : cell * (0.5 * (1.0 - v));
      else if (!saturated) cell *= 0.5 * (1.0 + v);
    }
    total += cell;
    if (k == owner) mine = cell;
  }
  if (valid) w[p] *= mine / total;
}

// The same partition with every unordered atom pair visited ONCE: v_{hi,lo} feeds cell[hi] (factor (1 - v)/2) and
// cell[lo] (factor (1 + v)/2).  Walking hi ascending and lo < hi ascending hands cell[k] its factors in the order
// j = 0 .. natoms-1 of k_becke, so the products are the same; the division by the interatomic distance becomes a
// multiplication with its precomputed reciprocal (1 ulp in v, 1e-16 in the weights).  Shared memory: distances and
// cell products, 2 x natoms x THREADS doubles; THREADS is 64 or 32, larger molecules use k_becke.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_becke_pairs(int natoms, const double *__restrict__ pos,
                                                         const double *__restrict__ rdist, const double *__restrict__ aij,
                                                         const size_t *__restrict__ atom_off,
                                                         const double *__restrict__ xyz_all, double *__restrict__ w_all) {
  extern __shared__ double sm[]; // gd[natoms][THREADS], cell[natoms][THREADS]
  double *gd = sm, *cell = sm + static_cast<size_t>(natoms) * THREADS;
  const int owner = blockIdx.y;
  const int npts = static_cast<int>(atom_off[owner + 1] - atom_off[owner]);
  if (static_cast<int>(blockIdx.x) * THREADS >= npts) return;
  const double *xyz = xyz_all + 3 * atom_off[owner];
  double *w = w_all + atom_off[owner];
  const int t = threadIdx.x;
  const int p = blockIdx.x * THREADS + t;
  const bool valid = p < npts;
  const double x = valid ? xyz[3 * p] : 0.0, y = valid ? xyz[3 * p + 1] : 0.0, z = valid ? xyz[3 * p + 2] : 0.0;
  for (int a = 0; a < natoms; ++a) {
    const double dx = x - pos[3 * a], dy = y - pos[3 * a + 1], dz = z - pos[3 * a + 2];
    gd[a * THREADS + t] = sqrt(dx * dx + dy * dy + dz * dz);
    cell[a * THREADS + t] = 1.0;
  }
  for (int hi = 1; hi < natoms; ++hi) {
    const double ghi = gd[hi * THREADS + t];
    double chi = cell[hi * THREADS + t];
    const double *rd = rdist + static_cast<size_t>(hi) * natoms, *ar = aij + static_cast<size_t>(hi) * natoms;
#pragma unroll 4
    for (int lo = 0; lo < hi; ++lo) {
      double v = (ghi - gd[lo * THREADS + t]) * __ldg(rd + lo);
      const double a = __ldg(ar + lo);
      v = fma(a, 1.0 - v * v, v); // a = 0 for equal Bragg radii: v unchanged
      v = (3 - v * v) * v * 0.5;
      v = (3 - v * v) * v * 0.5;
      v = (3 - v * v) * v * 0.5;
      const bool saturated = fabs(1.0 - v) < 1e-14;
      chi = saturated ? 0.0 : chi * (0.5 * (1.0 - v));
      if (!saturated) cell[lo * THREADS + t] *= 0.5 * (1.0 + v);
    }
    cell[hi * THREADS + t] = chi;
  }
  double total = 0.0;
  for (int k = 0; k < natoms; ++k) total += cell[k * THREADS + t];
  if (valid) w[p] *= cell[owner * THREADS + t] / total;
}

// 21 bits of each coordinate interleaved x -> bit 0, y -> bit 1, z -> bit 2 (spatial_grid_hierarchy.cpp:11-24)
inline uint64_t morton63(uint64_t qx, uint64_t qy, uint64_t qz) {
  uint64_t code = 0;
  for (int b = 0; b < 21; ++b) {
    code |= ((qx >> b) & 1ULL) << (3 * b);
    code |= ((qy >> b) & 1ULL) << (3 * b + 1);
    code |= ((qz >> b) & 1ULL) << (3 * b + 2);
  }
  return code;
}

} // namespace

struct occ_xc_grid {
  std::vector<double> pts, w;   // atom-ordered molecular grid (MolecularGridPoints)
  std::vector<double> spts, sw; // Morton-sorted copies
  std::vector<size_t> perm;
  std::vector<occ_xc_leaf_t> leaves;
  std::vector<size_t> atom_npts;
  std::vector<int> atom_nradial;
};

static void build_hierarchy(occ_xc_grid *g) {
  const size_t n = g->w.size();
  g->spts.resize(3 * n);
  g->sw.resize(n);
  g->perm.resize(n);
  g->leaves.clear();
  if (n == 0) return;
  const double *P = g->pts.data();
  double lo[3] = {P[0], P[1], P[2]}, hi[3] = {P[0], P[1], P[2]};
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], P[3 * i + d]);
      hi[d] = std::max(hi[d], P[3 * i + d]);
    }
  double ext[3];
  for (int d = 0; d < 3; ++d) {
    ext[d] = hi[d] - lo[d];
    if (ext[d] < 1e-10) ext[d] = 1.0;
  }
  const uint64_t qmax = (1ULL << 21) - 1;
  std::vector<std::pair<uint64_t, size_t>> keyed(n);
  for (size_t i = 0; i < n; ++i) {
    uint64_t q[3];
    for (int d = 0; d < 3; ++d) {
      const double u = std::min(std::max((P[3 * i + d] - lo[d]) / ext[d], 0.0), 1.0);
      q[d] = static_cast<uint64_t>(u * qmax);
    }
    keyed[i] = {morton63(q[0], q[1], q[2]), i};
  }
  std::sort(keyed.begin(), keyed.end());
  for (size_t i = 0; i < n; ++i) {
    const size_t o = keyed[i].second;
    g->perm[i] = o;
    g->spts[3 * i] = P[3 * o];
    g->spts[3 * i + 1] = P[3 * o + 1];
    g->spts[3 * i + 2] = P[3 * o + 2];
    g->sw[i] = g->w[o];
  }
  // leaves: SpatialHierarchySettings{target 128, min 32, max 256} (spatial_grid_hierarchy.h:26-30, .cpp:94-124)
  size_t at = 0;
  while (at < n) {
    const size_t left = n - at;
    const size_t take = left <= 256 ? left : (left < 64 ? left / 2 : std::min<size_t>(128, left));
    occ_xc_leaf_t lf;
    lf.offset = at;
    lf.count = take;
    double c[3] = {0.0, 0.0, 0.0};
    for (size_t i = at; i < at + take; ++i)
      for (int d = 0; d < 3; ++d) c[d] += g->spts[3 * i + d];
    for (int d = 0; d < 3; ++d) lf.center[d] = c[d] / static_cast<double>(take);
    double rad = 0.0;
    for (size_t i = at; i < at + take; ++i) {
      const double dx = g->spts[3 * i] - lf.center[0], dy = g->spts[3 * i + 1] - lf.center[1],
                   dz = g->spts[3 * i + 2] - lf.center[2];
      rad = std::max(rad, std::sqrt(dx * dx + dy * dy + dz * dz));
    }
    lf.radius = rad;
    g->leaves.push_back(lf);
    at += take;
  }
}

extern "C" {

int occ_xc_register_lebedev(int npoints, const double *xyzw) {
  if (npoints <= 0 || !xyzw) return occ_xc_set_error("occ_xc_register_lebedev: bad argument");
  lebedev_tables()[npoints] = std::vector<double>(xyzw, xyzw + 4 * static_cast<size_t>(npoints));
  return 0;
}

int occ_xc_grid_from_points(size_t npts, const double *pts_xyz, const double *weights, occ_xc_grid **out) {
  if (!pts_xyz || !weights || !out) return occ_xc_set_error("occ_xc_grid_from_points: null argument");
  occ_xc_grid *g = new occ_xc_grid();
  g->pts.assign(pts_xyz, pts_xyz + 3 * npts);
  g->w.assign(weights, weights + npts);
  build_hierarchy(g);
  *out = g;
  return 0;
}

int occ_xc_grid_create(int natoms, const int *Z, const double *pos, const occ_xc_basis_t *basis,
                       const occ_xc_grid_settings_t *settings, occ_xc_grid **out) {
  if (natoms <= 0 || !Z || !pos || !basis || !settings || !out || !basis->shell_atom)
    return occ_xc_set_error("occ_xc_grid_create: null argument (shell_atom is required)");
  occ_xc_grid_settings_t s = *settings;
  // MolecularGrid::ensure_settings, molecular_grid.cpp:133-181
  const bool orca = s.angular_regions[0] > 0; // has_angular_regions(): the COSX grids of GridSettings::for_cosx
  if (!orca && s.max_angular_points < s.min_angular_points) s.max_angular_points = s.min_angular_points + 1;
  s.max_angular_points = lebedev_at_or_above(s.max_angular_points);
  s.min_angular_points = lebedev_at_or_above(s.min_angular_points);
  if (s.max_angular_points < 0 || s.min_angular_points < 0)
    return occ_xc_set_error("Requested number of angular points exceeds 5810");

  // initialize_from_basis, molecular_grid.cpp:183-256
  int lmax = 0;
  for (int sh = 0; sh < basis->nsh; ++sh) lmax = std::max(lmax, basis->l[sh]);
  std::vector<double> amax(natoms, 0.0);
  std::vector<std::vector<double>> amin(natoms, std::vector<double>(lmax + 1, DBL_MAX));
  std::vector<char> has_shells(natoms, 0);
  for (int sh = 0; sh < basis->nsh; ++sh) {
    const int a = basis->shell_atom[sh];
    if (a < 0 || a >= natoms) return occ_xc_set_error("occ_xc_grid_create: shell_atom out of range");
    has_shells[a] = 1;
    for (int k = 0; k < basis->nprim[sh]; ++k) {
      const double e = basis->exps[basis->prim_off[sh] + k];
      amin[a][basis->l[sh]] = std::min(amin[a][basis->l[sh]], e);
      amax[a] = std::max(amax[a], e);
    }
  }
  // one grid per element, from the FIRST atom of that element (molecular_grid.cpp:236-252)
  std::map<int, ElementGrid> egrids;
  for (int a = 0; a < natoms; ++a) {
    if (egrids.count(Z[a])) continue;
    Radial rad;
    int regions[5] = {0, 0, 0, 0, 0};
    if (orca) { // grid_utils.cpp:588-611: IntAcc radial count, Treutler-Ahlrichs nodes scaled by the Bragg radius
      int nr = 50;
      if (s.int_acc > 0) {
        const int z = Z[a];
        const int row = z <= 2 ? 1 : (z <= 10 ? 2 : (z <= 18 ? 3 : (z <= 36 ? 4 : (z <= 54 ? 5 : (z <= 86 ? 6 : 7)))));
        nr = static_cast<int>(std::max(1.0, std::round((15.0 * s.int_acc - 40.0) + 5.0 * row))); // grid_settings.cpp:176-181
      }
      rad = treutler_ahlrichs_radial(nr, bragg_bohr(Z[a]));
      for (int k = 0; k < 5; ++k) { // grid_utils.cpp:693-701
        regions[k] = (s.reduced_first_row_element_grid && Z[a] < 3) ? lebedev_below(s.angular_regions[k]) : s.angular_regions[k];
        if (lebedev_at_or_above(regions[k]) < 0) return occ_xc_set_error("Requested number of angular points exceeds 5810");
      }
    } else if (!has_shells[a] || amax[a] <= 0.0 || lmax <= 0) // generate_atom_grid defaults, grid_utils.cpp:617-626
      rad = lmg_radial(Z[a], s.radial_precision, 20.0, 3, std::vector<double>(4, 0.1));
    else
      rad = lmg_radial(Z[a], s.radial_precision, amax[a], lmax, amin[a]);
    int max_ang = s.max_angular_points;
    if (s.reduced_first_row_element_grid && Z[a] < 3) max_ang = lebedev_below(max_ang); // grid_utils.cpp:682-687
    ElementGrid eg;
    eg.nradial = static_cast<int>(rad.r.size());
    for (size_t i = 0; i < rad.r.size(); ++i) {
      int order;
      if (orca) { // prune_orca_scheme, grid_utils.cpp:296-335
        const double R = bragg_bohr(Z[a]), r = rad.r[i];
        const int region = r < 0.25 * R ? 0 : (r < 0.50 * R ? 1 : (r < 1.00 * R ? 2 : (r < 2.50 * R ? 3 : 4)));
        order = lebedev_at_or_above(regions[region]);
      } else if (s.pruning_scheme == 1) { // PruningScheme::None, grid_utils.cpp:719-722
        order = max_ang;
      } else {
        order = lebedev_at_or_above(pruned_order(Z[a], max_ang, rad.r[i]));
      }
      auto it = lebedev_tables().find(order);
      if (it == lebedev_tables().end())
        return occ_xc_set_error(("Lebedev table of order " + std::to_string(order) + " not registered").c_str());
      const std::vector<double> &leb = it->second;
      const double r = rad.r[i], wr = rad.w[i] * (4 * M_PI); // grid_utils.cpp:679
      for (int k = 0; k < order; ++k) {
        eg.xyz.push_back(leb[4 * k] * r);
        eg.xyz.push_back(leb[4 * k + 1] * r);
        eg.xyz.push_back(leb[4 * k + 2] * r);
        eg.w.push_back(leb[4 * k + 3] * wr);
      }
    }
    egrids[Z[a]] = std::move(eg);
  }

  // pair tables: distances (grid_utils.cpp:108-123) and the radius-adjustment coefficient a_ij for i > j
  std::vector<double> dist(static_cast<size_t>(natoms) * natoms, 0.0), aij(dist.size(), 0.0);
  for (int i = 0; i < natoms; ++i)
    for (int j = 0; j < i; ++j) {
      const double dx = pos[3 * j] - pos[3 * i], dy = pos[3 * j + 1] - pos[3 * i + 1], dz = pos[3 * j + 2] - pos[3 * i + 2];
      const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
      dist[static_cast<size_t>(i) * natoms + j] = dist[static_cast<size_t>(j) * natoms + i] = d;
      const double ri = bragg_bohr(Z[i]), rj = bragg_bohr(Z[j]);
      if (std::fabs(ri - rj) > 1e-14) {
        const double xi = s.treutler_alrichs_adjustment ? std::sqrt(ri / rj) : ri / rj; // grid_utils.cpp:125-147
        const double u = (xi - 1) / (xi + 1);
        aij[static_cast<size_t>(i) * natoms + j] = u / (u * u - 1.0);
      }
    }

  occ_xc_grid *g = new occ_xc_grid();
  size_t total = 0;
  for (int a = 0; a < natoms; ++a) total += egrids[Z[a]].w.size();
  g->pts.resize(3 * total);
  g->w.resize(total);
  size_t off = 0;
  for (int a = 0; a < natoms; ++a) { // molecular_grid.cpp:66-67, molecular_grid_points.cpp:15-50
    const ElementGrid &eg = egrids[Z[a]];
    const size_t n = eg.w.size();
    for (size_t p = 0; p < n; ++p) {
      g->pts[3 * (off + p)] = eg.xyz[3 * p] + pos[3 * a];
      g->pts[3 * (off + p) + 1] = eg.xyz[3 * p + 1] + pos[3 * a + 1];
      g->pts[3 * (off + p) + 2] = eg.xyz[3 * p + 2] + pos[3 * a + 2];
      g->w[off + p] = eg.w[p];
    }
    g->atom_npts.push_back(n);
    g->atom_nradial.push_back(eg.nradial);
    off += n;
  }

  // Becke weights on the GPU
  {
    double *d_pos = nullptr, *d_dist = nullptr, *d_aij = nullptr, *d_xyz = nullptr, *d_w = nullptr;
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    chk(cudaMalloc(&d_pos, 3 * natoms * sizeof(double)));
    chk(cudaMalloc(&d_dist, dist.size() * sizeof(double)));
    chk(cudaMalloc(&d_aij, aij.size() * sizeof(double)));
    chk(cudaMalloc(&d_xyz, std::max<size_t>(1, 3 * total) * sizeof(double)));
    chk(cudaMalloc(&d_w, std::max<size_t>(1, total) * sizeof(double)));
    if (e == cudaSuccess) {
      chk(cudaMemcpy(d_pos, pos, 3 * natoms * sizeof(double), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(d_dist, dist.data(), dist.size() * sizeof(double), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(d_aij, aij.data(), aij.size() * sizeof(double), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(d_xyz, g->pts.data(), 3 * total * sizeof(double), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(d_w, g->w.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    }
    // pair-once kernel when distances and cell products of 64 or 32 points fit shared memory (up to ~430 atoms)
    const size_t smem_pairs64 = 2 * static_cast<size_t>(natoms) * 64 * sizeof(double), smem_pairs32 = smem_pairs64 / 2;
    const int pair_threads = smem_pairs64 <= 200 * 1024 ? 64 : (smem_pairs32 <= 200 * 1024 ? 32 : 0);
    std::vector<double> rdist(dist.size(), 0.0);
    for (size_t i = 0; i < dist.size(); ++i) rdist[i] = dist[i] != 0.0 ? 1.0 / dist[i] : 0.0;
    if (e == cudaSuccess && pair_threads) chk(cudaMemcpy(d_dist, rdist.data(), rdist.size() * sizeof(double), cudaMemcpyHostToDevice));
    const size_t smem = static_cast<size_t>(natoms) * BECKE_THREADS * sizeof(double);
    if (e == cudaSuccess && !pair_threads && smem > 48 * 1024)
      chk(cudaFuncSetAttribute(k_becke, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (e == cudaSuccess && !pair_threads && smem > 220 * 1024) {
      e = cudaErrorInvalidValue;
    }
    std::vector<size_t> atom_off(natoms + 1, 0);
    size_t max_n = 0;
    for (int a = 0; a < natoms; ++a) {
      atom_off[a + 1] = atom_off[a] + g->atom_npts[a];
      max_n = std::max(max_n, g->atom_npts[a]);
    }
    size_t *d_off = nullptr;
    chk(cudaMalloc(&d_off, atom_off.size() * sizeof(size_t)));
    if (e == cudaSuccess)
      chk(cudaMemcpy(d_off, atom_off.data(), atom_off.size() * sizeof(size_t), cudaMemcpyHostToDevice));
    if (e == cudaSuccess && max_n > 0) {
      if (pair_threads == 64) {
        if (smem_pairs64 > 48 * 1024)
          chk(cudaFuncSetAttribute(k_becke_pairs<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_pairs64)));
        const dim3 grid_dim(static_cast<unsigned>((max_n + 63) / 64), natoms);
        if (e == cudaSuccess) k_becke_pairs<64><<<grid_dim, 64, smem_pairs64>>>(natoms, d_pos, d_dist, d_aij, d_off, d_xyz, d_w);
      } else if (pair_threads == 32) {
        if (smem_pairs32 > 48 * 1024)
          chk(cudaFuncSetAttribute(k_becke_pairs<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_pairs32)));
        const dim3 grid_dim(static_cast<unsigned>((max_n + 31) / 32), natoms);
        if (e == cudaSuccess) k_becke_pairs<32><<<grid_dim, 32, smem_pairs32>>>(natoms, d_pos, d_dist, d_aij, d_off, d_xyz, d_w);
      } else {
        const dim3 grid_dim(static_cast<unsigned>((max_n + BECKE_THREADS - 1) / BECKE_THREADS), natoms);
        k_becke<<<grid_dim, BECKE_THREADS, smem>>>(natoms, d_pos, d_dist, d_aij, d_off, d_xyz, d_w);
      }
    }
    if (e == cudaSuccess) chk(cudaGetLastError());
    if (e == cudaSuccess) chk(cudaMemcpy(g->w.data(), d_w, total * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_pos); cudaFree(d_dist); cudaFree(d_aij); cudaFree(d_xyz); cudaFree(d_w); cudaFree(d_off);
    if (e != cudaSuccess) {
      delete g;
      return occ_xc_set_error((std::string("occ_xc_grid_create: CUDA error in the Becke partition: ") +
                               cudaGetErrorString(e)).c_str());
    }
  }
  build_hierarchy(g);
  *out = g;
  return 0;
}

size_t occ_xc_grid_npts(const occ_xc_grid *g) { return g ? g->w.size() : 0; }
size_t occ_xc_grid_nleaves(const occ_xc_grid *g) { return g ? g->leaves.size() : 0; }
int occ_xc_grid_atom_info(const occ_xc_grid *g, int atom, size_t *npts, int *nradial) {
  if (!g || atom < 0 || atom >= static_cast<int>(g->atom_npts.size()))
    return occ_xc_set_error("occ_xc_grid_atom_info: bad argument");
  if (npts) *npts = g->atom_npts[atom];
  if (nradial) *nradial = g->atom_nradial[atom];
  return 0;
}
int occ_xc_grid_copy(const occ_xc_grid *g, double *sorted_pts, double *sorted_w, occ_xc_leaf_t *leaves, size_t *perm,
                     double *pts, double *w) {
  if (!g) return occ_xc_set_error("occ_xc_grid_copy: null grid");
  const size_t n = g->w.size();
  if (sorted_pts) std::memcpy(sorted_pts, g->spts.data(), 3 * n * sizeof(double));
  if (sorted_w) std::memcpy(sorted_w, g->sw.data(), n * sizeof(double));
  if (leaves) std::memcpy(leaves, g->leaves.data(), g->leaves.size() * sizeof(occ_xc_leaf_t));
  if (perm) std::memcpy(perm, g->perm.data(), n * sizeof(size_t));
  if (pts) std::memcpy(pts, g->pts.data(), 3 * n * sizeof(double));
  if (w) std::memcpy(w, g->w.data(), n * sizeof(double));
  return 0;
}
void occ_xc_grid_destroy(occ_xc_grid *g) { delete g; }

} // extern "C"
