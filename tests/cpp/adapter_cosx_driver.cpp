// tests/cpp/adapter_cosx_driver.cpp — compile/link check of adapter/occ_cosx_b200.h; without a device the constructor
// must throw std::runtime_error (no CPU fallback), with one it builds K for two s shells and prints K(0,0).
#include "../../adapter/occ_cosx_b200.h"
#include <cstdio>
#include <cstring>

int main() {
  using namespace occ_b200;
  BasisArrays b;
  const double o1[3] = {0, 0, 0}, o2[3] = {0, 0, 1.4};
  const double e = 0.8, c = 1.0;
  b.add_shell(0, o1, 1, &e, &c, 0);
  b.add_shell(0, o2, 1, &e, &c, 1);
  std::vector<size_t> off = {0, 2, 4};
  const double pts[12] = {0.1, 0, 0, 0, 0.2, 0, 0, 0, 1.3, 0.2, 0, 1.5};
  const double w[4] = {0.5, 0.5, 0.5, 0.5};
  try {
    SemiNumericalExchangeB200 sgx(b, off, pts, w, nullptr);
    const double D[4] = {0.5, 0.1, 0.1, 0.5};
    double K[4];
    sgx.compute_K(SpinorbitalKind::Restricted, D, K);
    std::printf("ok K00=%.12f batches=%zu\n", K[0], sgx.num_batches());
  } catch (const std::runtime_error &err) {
    std::printf("runtime_error: %s\n", err.what());
    return std::strstr(err.what(), "CUDA") ? 0 : 2;
  }
  return 0;
}
