// oracle/src/grid.cpp — CPU ORACLE (test infrastructure): molecular quadrature grid + Morton leaves.
// Restates (paths under /root/reference):
//   src/numint/grid_utils.cpp:15-41,78-95      Lebedev levels, Bragg radii, Becke 3x polynomial
//   src/numint/grid_utils.cpp:125-200          Treutler-Ahlrichs adjustment, pairwise cell functions
//   src/numint/grid_utils.cpp:234-274          NWChem pruning
//   src/numint/grid_utils.cpp:398-514          LMG radial grid (inner/outer/h solves)
//   src/numint/grid_utils.cpp:576-758          generate_atom_grid
//   src/numint/molecular_grid.cpp:46-82,183-256  per-atom basis info, unique-element grids, partitioning
//   src/numint/molecular_grid_points.cpp:15-50   concatenation
//   src/numint/spatial_grid_hierarchy.cpp:11-147 Morton order, leaves, bounding spheres
#include "oracle.h"
#include "oracle_internal.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <utility>

namespace orc {

static const int kLebedevLevels[33] = {1,    6,    14,   26,   38,   50,   74,   86,   110,  146,  170,
                                       194,  230,  266,  302,  350,  434,  590,  770,  974,  1202, 1454,
                                       1730, 2030, 2354, 2702, 3074, 3470, 3890, 4334, 4802, 5294, 5810};

// Bragg-Slater radii in Angstrom, Z = 1..131 (grid_utils.cpp:22-41)
static const double kBragg[131] = {
    0.35, 1.40, 1.45, 1.05, 0.85, 0.70, 0.65, 0.60, 0.50, 1.50, 1.80, 1.50, 1.25, 1.10, 1.00, 1.00, 1.00, 1.80, 2.20,
    1.80, 1.60, 1.40, 1.35, 1.40, 1.40, 1.40, 1.35, 1.35, 1.35, 1.35, 1.30, 1.25, 1.15, 1.15, 1.15, 1.90, 2.35, 2.00,
    1.80, 1.55, 1.45, 1.45, 1.35, 1.30, 1.35, 1.40, 1.60, 1.55, 1.55, 1.45, 1.45, 1.40, 1.40, 2.10, 2.60, 2.15, 1.95,
    1.85, 1.85, 1.85, 1.85, 1.85, 1.85, 1.80, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.55, 1.45, 1.35, 1.35, 1.30,
    1.35, 1.35, 1.35, 1.50, 1.90, 1.80, 1.60, 1.90, 1.45, 2.10, 1.80, 2.15, 1.95, 1.80, 1.80, 1.75, 1.75, 1.75, 1.75,
    1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75,
    1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75, 1.75};

static const double kAngstromToBohr = 1.0 / 0.52917721067; // include/occ/core/units.h:7-8

static double atomic_radius(int Z) { // grid_utils.cpp:78-86
  if (Z < 1 || Z > 131) return 1.0 * kAngstromToBohr;
  return kBragg[Z - 1] * kAngstromToBohr;
}

static int level_at_or_above(int n) { // grid_utils.cpp:43-53
  for (int i = 0; i < 33; ++i)
    if (kLebedevLevels[i] >= n) return kLebedevLevels[i];
  throw std::runtime_error("Requested number of angular points exceeds 5810");
}
static int level_below(int n) { // grid_utils.cpp:55-69
  if (n <= kLebedevLevels[0]) return kLebedevLevels[0];
  for (int i = 32; i >= 0; --i)
    if (kLebedevLevels[i] < n) return kLebedevLevels[i];
  return kLebedevLevels[0];
}

// ---- LMG (Lindh, Malmqvist, Gagliardi, TCA 106, 178 (2001)) ----
static double lmg_inner(double max_error, double alpha_inner) { // grid_utils.cpp:398-409 (eq. 25, m = 0)
  const int m = 0;
  double r = 1.9 - std::log(1.0 / max_error);
  r = r * 2.0 / (m + 3.0);
  r = std::exp(r) / alpha_inner;
  return std::sqrt(r);
}

// grid_utils.cpp:412-441 (eq. 19): decimating step search; also exercised with alpha = DBL_MAX (a15 quirk).
static double lmg_outer(double max_error, double alpha_outer, int l, double guess) {
  const int m = 2 * l;
  double r = guess, r_old = 1.0e50, step = 0.5, f = 1.0e50;
  double sign = (f > max_error) ? 1.0 : -1.0, sign_old;
  while (std::fabs(r_old - r) > 1e-14) {
    const double c = std::tgamma((m + 3.0) / 2.0);
    const double a = std::pow(alpha_outer * r * r, (m + 1.0) / 2.0);
    const double e = std::exp(-alpha_outer * r * r);
    f = c * a * e;
    sign_old = sign;
    sign = (f > max_error) ? 1.0 : -1.0;
    if (r < 0.0) sign = 1.0;
    if (sign != sign_old) step *= 0.1;
    r_old = r;
    r += sign * step;
  }
  return r;
}

static double lmg_h(double max_error, int l, double guess) { // grid_utils.cpp:444-475 (eqs. 17, 18)
  const int m = 2 * l;
  double h = guess, h_old = 1.0e50, step = 0.1 * guess, f = 1.0e50;
  double sign = (f > max_error) ? -1.0 : 1.0, sign_old;
  while (std::fabs(h_old - h) > 1e-14) {
    const double c0 = 4.0 * std::sqrt(2.0) * M_PI;
    const double cm = std::tgamma(3.0 / 2.0) / std::tgamma((m + 3.0) / 2.0);
    const double p0 = 1.0 / h;
    const double e0 = std::exp(-M_PI * M_PI / (2.0 * h));
    const double pm = std::pow(M_PI / h, m / 2.0);
    const double rd0 = c0 * p0 * e0;
    f = cm * pm * rd0;
    sign_old = sign;
    sign = (f > max_error) ? -1.0 : 1.0;
    if (h < 0.0) sign = 1.0;
    if (sign != sign_old) step *= 0.1;
    h_old = h;
    h += sign * step;
  }
  return h;
}

struct Radial {
  std::vector<double> r, w;
};

static Radial lmg_radial(int Z, double precision, double alpha_max, int l_max, const double *alpha_min) {
  // grid_utils.cpp:478-514
  const double r_inner = lmg_inner(precision, 2 * alpha_max);
  double h = FLT_MAX;
  double r_outer = 0.0;
  const double br = atomic_radius(Z);
  for (int i = 0; i <= l_max; ++i) {
    if (alpha_min[i] > 0.0) {
      r_outer = std::max(r_outer, lmg_outer(precision, alpha_min[i], i, 4.0 * br));
      h = std::min(h, lmg_h(precision, i, 0.1 * (r_outer - r_inner)));
    }
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

// generate_treutler_alrichs_radial_grid, grid_utils.cpp:559-572 (points ascending)
static Radial treutler_ahlrichs_radial(size_t n) {
  Radial g;
  g.r.resize(n);
  g.w.resize(n);
  const double step = M_PI / (n + 1), ln2 = 1 / std::log(2.0);
  for (size_t i = 1; i <= n; ++i) {
    const double x = std::cos(i * step);
    const double tmp1 = ln2 * std::pow(1 + x, 0.6);
    const double tmp2 = std::log((1 - x) / 2);
    g.r[n - i] = -tmp1 * tmp2;
    g.w[n - i] = step * std::sin(i * step) * tmp1 * (-0.6 / (1 + x) * tmp2 + 1 / (1 - x));
  }
  return g;
}

// calculate_radial_points_orca, grid_settings.cpp:176-181 with get_periodic_row (:21-30)
static size_t radial_points_orca(double int_acc, int Z) {
  const int row = Z <= 2 ? 1 : (Z <= 10 ? 2 : (Z <= 18 ? 3 : (Z <= 36 ? 4 : (Z <= 54 ? 5 : (Z <= 86 ? 6 : 7)))));
  const double nr = (15.0 * int_acc - 40.0) + 5.0 * row;
  return static_cast<size_t>(std::max(1.0, std::round(nr)));
}

// prune_orca_scheme, grid_utils.cpp:296-335
static std::vector<int> prune_orca(int Z, const int regions[5], const std::vector<double> &radii) {
  const double R = atomic_radius(Z);
  std::vector<int> out(radii.size());
  for (size_t i = 0; i < radii.size(); ++i) {
    const double r = radii[i];
    const int region = r < 0.25 * R ? 0 : (r < 0.50 * R ? 1 : (r < 1.00 * R ? 2 : (r < 2.50 * R ? 3 : 4)));
    out[i] = level_at_or_above(regions[region]);
  }
  return out;
}

// grid_utils.cpp:234-274
static std::vector<int> prune_nwchem(int Z, int max_angular, const std::vector<double> &radii) {
  std::vector<int> out(radii.size());
  int lvl[5];
  if (max_angular < 50) {
    std::fill(out.begin(), out.end(), max_angular);
    return out;
  } else if (max_angular == 50) {
    lvl[0] = 5; lvl[1] = 6; lvl[2] = 6; lvl[3] = 6; lvl[4] = 5;
  } else {
    int i;
    for (i = 6; i < 33; ++i)
      if (kLebedevLevels[i] == max_angular) break;
    lvl[0] = 5; lvl[1] = 6; lvl[2] = i - 1; lvl[3] = i; lvl[4] = i - 1;
  }
  double alphas[4] = {0.1, 0.4, 0.8, 2.5};
  if (Z <= 2) {
    alphas[0] = 0.25; alphas[1] = 0.5; alphas[2] = 1.0; alphas[3] = 4.5;
  } else if (Z <= 10) {
    alphas[0] = 0.16666667; alphas[1] = 0.5; alphas[2] = 0.9; alphas[3] = 3.5;
  }
  const double radius = atomic_radius(Z);
  for (size_t i = 0; i < radii.size(); ++i) {
    const double scale = radii[i] / radius;
    int place = 0;
    while (place < 4 && !(alphas[place] > scale)) ++place;
    out[i] = kLebedevLevels[lvl[place]];
  }
  return out;
}

struct AtomGrid {
  int Z = 0;
  int nradial = 0;
  std::vector<double> pts; // 3 x n, column-major (xyz interleaved)
  std::vector<double> w;
};

static std::map<int, std::vector<double>> g_lebedev;

const std::vector<double> *lebedev_table(int n) {
  auto it = g_lebedev.find(n);
  return it == g_lebedev.end() ? nullptr : &it->second;
}

// grid_utils.cpp:576-758 (LMG branch, NWChem pruning)
static AtomGrid generate_atom_grid(int Z, const orc_grid_settings_t &s, double alpha_max, int l_max,
                                   const std::vector<double> &alpha_min) {
  Radial radial;
  const bool orca = s.angular_regions[0] > 0; // has_angular_regions(), :598
  if (orca) { // :600-611: Treutler-Ahlrichs radial grid scaled by the Bragg radius, r^2 folded into the weights
    const size_t n_radial = s.int_acc > 0 ? radial_points_orca(s.int_acc, Z) : 50; // :588-596
    radial = treutler_ahlrichs_radial(n_radial);
    const double rm = atomic_radius(Z);
    for (size_t i = 0; i < n_radial; ++i) {
      radial.r[i] *= rm;
      radial.w[i] *= rm;
      radial.w[i] *= radial.r[i] * radial.r[i];
    }
  } else if (alpha_max <= 0.0 || l_max <= 0 || alpha_min.empty()) { // grid_utils.cpp:617-626 defaults
    std::vector<double> def(4, 0.1);
    radial = lmg_radial(Z, s.radial_precision, 20.0, 3, def.data());
  } else {
    radial = lmg_radial(Z, s.radial_precision, alpha_max, l_max, alpha_min.data());
  }
  for (auto &w : radial.w) w *= 4 * M_PI; // :679
  int max_angular = s.max_angular_points;
  if (s.reduced_first_row_element_grid && Z < 3) max_angular = level_below(max_angular); // :682-687
  std::vector<int> n_angular;
  if (orca) { // :693-705
    int regions[5];
    for (int k = 0; k < 5; ++k)
      regions[k] = (s.reduced_first_row_element_grid && Z < 3) ? level_below(s.angular_regions[k]) : s.angular_regions[k];
    n_angular = prune_orca(Z, regions, radial.r);
  } else if (s.pruning_scheme == 1) { // PruningScheme::None, :719-722
    n_angular.assign(radial.r.size(), max_angular);
  } else {
    n_angular = prune_nwchem(Z, max_angular, radial.r);
  }
  size_t total = 0;
  for (int na : n_angular) total += level_at_or_above(na);
  AtomGrid g;
  g.Z = Z;
  g.nradial = static_cast<int>(radial.r.size());
  g.pts.resize(3 * total);
  g.w.resize(total);
  size_t off = 0;
  for (size_t i = 0; i < radial.r.size(); ++i) {
    const std::vector<double> *leb = lebedev_table(n_angular[i]);
    if (!leb) throw std::runtime_error("Lebedev table not registered for order " + std::to_string(n_angular[i]));
    const size_t na = leb->size() / 4;
    const double r = radial.r[i], w = radial.w[i];
    for (size_t k = 0; k < na; ++k) {
      g.pts[3 * (off + k) + 0] = (*leb)[4 * k + 0] * r;
      g.pts[3 * (off + k) + 1] = (*leb)[4 * k + 1] * r;
      g.pts[3 * (off + k) + 2] = (*leb)[4 * k + 2] * r;
      g.w[off + k] = (*leb)[4 * k + 3] * w;
    }
    off += na;
  }
  return g;
}

struct MolGrid {
  std::vector<double> pts, w;
  std::vector<size_t> atom_npts;
  std::vector<int> atom_nradial;
};

// molecular_grid.cpp:46-82 with grid_utils.cpp:149-200 (Becke cell functions, TA adjustment).
// Pair loops are kept in the reference's (i, j<i) order so each column's product is formed in the same order.
static void partition_atom_grid(AtomGrid &grid, int atom, int natoms, const double *pos, const std::vector<double> &radii,
                                const std::vector<double> &dists, bool ta_adjust) {
  const size_t n = grid.w.size();
  for (size_t p = 0; p < n; ++p) {
    grid.pts[3 * p + 0] += pos[3 * atom + 0];
    grid.pts[3 * p + 1] += pos[3 * atom + 1];
    grid.pts[3 * p + 2] += pos[3 * atom + 2];
  }
  std::vector<double> gd(natoms), cell(natoms);
  for (size_t p = 0; p < n; ++p) {
    const double x = grid.pts[3 * p], y = grid.pts[3 * p + 1], z = grid.pts[3 * p + 2];
    for (int a = 0; a < natoms; ++a) {
      const double dx = x - pos[3 * a], dy = y - pos[3 * a + 1], dz = z - pos[3 * a + 2];
      gd[a] = std::sqrt(dx * dx + dy * dy + dz * dz);
      cell[a] = 1.0;
    }
    for (int i = 0; i < natoms; ++i) {
      const double r_i = radii[i];
      for (int j = 0; j < i; ++j) {
        const double r_j = radii[j];
        double v = (gd[i] - gd[j]) / dists[static_cast<size_t>(i) * natoms + j];
        if (std::fabs(r_i - r_j) > 1e-14) {
          // TreutlerAlrichsAdjustment (grid_utils.cpp:125-135) or BeckeAdjustment (:137-147)
          const double xi = ta_adjust ? std::sqrt(r_i / r_j) : r_i / r_j;
          const double u = (xi - 1) / (xi + 1);
          const double a_ij = u / (u * u - 1.0);
          v += a_ij * (1 - v * v);
        }
        for (int k = 0; k < 3; ++k) v = (3 - v * v) * v * 0.5; // becke_partition, :88-95
        if (std::fabs(1.0 - v) < 1e-14) {
          cell[i] = 0.0;
        } else {
          cell[i] *= 0.5 * (1.0 - v);
          cell[j] *= 0.5 * (1.0 + v);
        }
      }
    }
    double sum = 0.0;
    for (int a = 0; a < natoms; ++a) sum += cell[a];
    grid.w[p] *= cell[atom] / sum; // molecular_grid.cpp:76-77
  }
}

} // namespace orc

extern "C" {

void orc_register_lebedev(int n, const double *xyzw) {
  orc::g_lebedev[n] = std::vector<double>(xyzw, xyzw + 4 * static_cast<size_t>(n));
}

int orc_lmg_radial(int Z, double precision, double alpha_max, int l_max, const double *alpha_min, double *r,
                   double *w, int cap) {
  orc::Radial g = orc::lmg_radial(Z, precision, alpha_max, l_max, alpha_min);
  const int n = static_cast<int>(g.r.size());
  for (int i = 0; i < n && i < cap; ++i) {
    if (r) r[i] = g.r[i];
    if (w) w[i] = g.w[i];
  }
  return n;
}

void *orc_molecular_grid_create(int natoms, const int *Z, const double *pos, const orc_basis_t *b,
                                const orc_grid_settings_t *settings, int nthreads) {
  using namespace orc;
  orc_grid_settings_t s = *settings;
  // molecular_grid.cpp:133-181 ensure_settings
  if (s.angular_regions[0] <= 0 && s.max_angular_points < s.min_angular_points)
    s.max_angular_points = s.min_angular_points + 1;
  s.max_angular_points = level_at_or_above(s.max_angular_points);
  s.min_angular_points = level_at_or_above(s.min_angular_points);

  // molecular_grid.cpp:183-256 initialize_from_basis
  int basis_lmax = 0;
  for (int sh = 0; sh < b->nsh; ++sh) basis_lmax = std::max(basis_lmax, b->l[sh]);
  std::vector<double> alpha_max(natoms, 0.0);
  std::vector<std::vector<double>> alpha_min(natoms);
  for (int a = 0; a < natoms; ++a) {
    bool any = false;
    std::vector<double> amin(basis_lmax + 1, DBL_MAX); // :214-216 (DBL_MAX for l the atom lacks)
    for (int sh = 0; sh < b->nsh; ++sh) {
      if (b->shell_atom[sh] != a) continue;
      any = true;
      const double *e = b->exps + b->prim_off[sh];
      for (int k = 0; k < b->nprim[sh]; ++k) {
        amin[b->l[sh]] = std::min(amin[b->l[sh]], e[k]);
        alpha_max[a] = std::max(alpha_max[a], e[k]);
      }
    }
    if (any) alpha_min[a] = amin;
  }
  std::vector<int> unique(Z, Z + natoms);
  std::sort(unique.begin(), unique.end());
  unique.erase(std::unique(unique.begin(), unique.end()), unique.end());
  std::map<int, AtomGrid> ugrids;
  for (int z : unique) {
    int first = 0;
    for (int a = 0; a < natoms; ++a)
      if (Z[a] == z) {
        first = a;
        break;
      }
    ugrids[z] = generate_atom_grid(z, s, alpha_max[first], basis_lmax, alpha_min[first]);
  }
  // grid_utils.cpp:108-123 interatomic distances; molecular_grid.cpp:297-302 default radii
  std::vector<double> dists(static_cast<size_t>(natoms) * natoms, 0.0), radii(natoms);
  for (int i = 0; i < natoms; ++i) {
    radii[i] = atomic_radius(Z[i]);
    for (int j = i + 1; j < natoms; ++j) {
      const double dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
      const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
      dists[static_cast<size_t>(i) * natoms + j] = d;
      dists[static_cast<size_t>(j) * natoms + i] = d;
    }
  }
  std::vector<AtomGrid> grids(natoms);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
  for (int a = 0; a < natoms; ++a) {
    grids[a] = ugrids.at(Z[a]);
    partition_atom_grid(grids[a], a, natoms, pos, radii, dists, s.treutler_alrichs_adjustment != 0);
  }
  auto *mg = new MolGrid();
  for (int a = 0; a < natoms; ++a) {
    mg->pts.insert(mg->pts.end(), grids[a].pts.begin(), grids[a].pts.end());
    mg->w.insert(mg->w.end(), grids[a].w.begin(), grids[a].w.end());
    mg->atom_npts.push_back(grids[a].w.size());
    mg->atom_nradial.push_back(grids[a].nradial);
  }
  return mg;
}

size_t orc_grid_npts(const void *g) { return static_cast<const orc::MolGrid *>(g)->w.size(); }
void orc_grid_copy(const void *g, double *pts, double *w) {
  auto *mg = static_cast<const orc::MolGrid *>(g);
  std::memcpy(pts, mg->pts.data(), sizeof(double) * mg->pts.size());
  std::memcpy(w, mg->w.data(), sizeof(double) * mg->w.size());
}
size_t orc_grid_atom_npts(const void *g, int atom) { return static_cast<const orc::MolGrid *>(g)->atom_npts[atom]; }
int orc_grid_atom_nradial(const void *g, int atom) { return static_cast<const orc::MolGrid *>(g)->atom_nradial[atom]; }
void orc_grid_destroy(void *g) { delete static_cast<orc::MolGrid *>(g); }

// ---- spatial_grid_hierarchy.cpp ----
static inline uint64_t dilate3(uint64_t x) { // :11-19 (standard 21-bit 3D bit spreading)
  x &= 0x1fffff;
  x = (x | (x << 32)) & 0x1f00000000ffffULL;
  x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
  x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
  x = (x | (x << 2)) & 0x1249249249249249ULL;
  return x;
}

size_t orc_hierarchy(size_t n, const double *pts, const double *w, double *spts, double *sw, size_t *perm,
                     orc_leaf_t *leaves, size_t leaves_cap) {
  if (n == 0) return 0;
  double lo[3], hi[3], ext[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = hi[d] = pts[d];
  }
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], pts[3 * i + d]);
      hi[d] = std::max(hi[d], pts[3 * i + d]);
    }
  for (int d = 0; d < 3; ++d) {
    ext[d] = hi[d] - lo[d];
    if (ext[d] < 1e-10) ext[d] = 1.0; // :60-62
  }
  std::vector<std::pair<uint64_t, size_t>> codes(n);
  const uint64_t scale = (1ULL << 21) - 1;
  for (size_t i = 0; i < n; ++i) {
    uint64_t q[3];
    for (int d = 0; d < 3; ++d) {
      double v = (pts[3 * i + d] - lo[d]) / ext[d];
      v = std::min(std::max(v, 0.0), 1.0);
      q[d] = static_cast<uint64_t>(v * scale);
    }
    codes[i] = {dilate3(q[0]) | (dilate3(q[1]) << 1) | (dilate3(q[2]) << 2), i};
  }
  std::sort(codes.begin(), codes.end()); // :79
  for (size_t i = 0; i < n; ++i) {
    const size_t o = codes[i].second;
    if (perm) perm[i] = o;
    spts[3 * i] = pts[3 * o];
    spts[3 * i + 1] = pts[3 * o + 1];
    spts[3 * i + 2] = pts[3 * o + 2];
    sw[i] = w[o];
  }
  // build_leaves :94-124 with SpatialHierarchySettings{128, 32, 256}
  const size_t target = 128, min_leaf = 32, max_leaf = 256;
  size_t nleaves = 0, offset = 0;
  while (offset < n) {
    const size_t remaining = n - offset;
    size_t count;
    if (remaining <= max_leaf)
      count = remaining;
    else if (remaining < 2 * min_leaf)
      count = remaining / 2;
    else
      count = std::min(target, remaining);
    if (nleaves < leaves_cap) {
      orc_leaf_t &lf = leaves[nleaves];
      lf.offset = offset;
      lf.count = count;
      // compute_bounding_sphere :126-147: centroid, max distance
      double c[3] = {0, 0, 0};
      for (size_t i = 0; i < count; ++i)
        for (int d = 0; d < 3; ++d) c[d] += spts[3 * (offset + i) + d];
      for (int d = 0; d < 3; ++d) lf.center[d] = c[d] / static_cast<double>(count);
      double rad = 0.0;
      for (size_t i = 0; i < count; ++i) {
        const double dx = spts[3 * (offset + i)] - lf.center[0], dy = spts[3 * (offset + i) + 1] - lf.center[1],
                     dz = spts[3 * (offset + i) + 2] - lf.center[2];
        rad = std::max(rad, std::sqrt(dx * dx + dy * dy + dz * dz));
      }
      lf.radius = rad;
    }
    ++nleaves;
    offset += count;
  }
  return nleaves;
}

} // extern "C"
