// extern "C" entry points over the UNMODIFIED reference sources compiled by oracle/Makefile into
// oracle/_ref/libocc_ref.so: gau2grid (src/3rdparty/gau2grid/*.c) exports its own C API; the
// Lebedev tables (src/numint/lebedev.cpp:5389-5503) need this one wrapper. Test infrastructure only.
#include <occ/numint/lebedev.h>
#include <cstring>
extern "C" int occ_ref_lebedev(int num_points, double *out_xyzw) {
  auto m = occ::numint::grid::lebedev(num_points);
  std::memcpy(out_xyzw, m.data(), sizeof(double) * 4 * static_cast<size_t>(num_points));
  return num_points;
}
