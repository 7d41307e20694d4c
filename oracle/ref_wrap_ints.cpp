// extern "C" entry points over the UNMODIFIED reference headers include/occ/ints/{boys,ecoeffs,rints}.h
// (header-only, standard library only) compiled by oracle/Makefile into oracle/_ref/libocc_ref.so.
// They pin the COSX oracle's Boys table, E coefficients and Hermite Coulomb integrals. Test infrastructure only.
#include <occ/ints/boys.h>
#include <occ/ints/ecoeffs.h>
#include <occ/ints/rints.h>
#include <cstring>
#include <memory>

using namespace occ::ints;

extern "C" {

// boys.h:332-375 — the table SemiNumericalExchange hands to its ESPEvaluator (occ::ints::boys().table())
int occ_ref_boys_table(double *out) {
  if (out) std::memcpy(out, boys().table(), BoysDefault::table_size_bytes());
  return BoysDefault::TableSize;
}
double occ_ref_boys(double x, int m) { return boys().compute(x, m); }
double occ_ref_boys_reference(double x, int m) { return BoysDefault::reference(x, m); }

// ecoeffs.h:150-198
int occ_ref_ecoeffs_1d(int la, int lb, double a, double b, double XAB, double *out) {
  auto E = std::make_unique<ECoeffs1DDynamic<double>>();
  compute_e_coeffs_1d_dynamic<double>(la, lb, a, b, XAB, *E);
  const int n = (la + 1) * (lb + 1) * (la + lb + 1);
  std::memcpy(out, E->data, sizeof(double) * n);
  return n;
}

// rints.h:158-227
int occ_ref_rints(int L, double p, double PCx, double PCy, double PCz, double *out) {
  auto R = std::make_unique<RIntsDynamic<double>>();
  compute_r_ints_dynamic<double, BoysParamsDefault>(boys().table(), L, p, PCx, PCy, PCz, *R);
  const int n = nhermsum(L);
  std::memcpy(out, R->data, sizeof(double) * n);
  return n;
}
}
