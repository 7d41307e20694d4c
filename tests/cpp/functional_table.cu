// tests/cpp/functional_table.cu — TEST HARNESS: evaluates the product's device functionals (occ_b200/csrc/xc_functionals.cuh,
// which are __host__ __device__) on the HOST for a table of points, so that the CPU test suite can compare them with the
// oracle and the reference's known answers without a GPU.  Not part of the product path.
//   usage: functional_table <func_id> <polarized 0|1> <family of the build 0|1|2> <n>   (stdin: rho, sigma, tau arrays; stdout: results)
#include "xc_functionals.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace xcb;

int main(int argc, char **argv) {
  if (argc != 5) return 2;
  int id = atoi(argv[1]);
  const int pol = atoi(argv[2]), fam = atoi(argv[3]);
  const size_t n = strtoull(argv[4], nullptr, 10);
  const int s = pol ? 2 : 1, g = pol ? 3 : 1;
  std::vector<double> rho(n * s), sigma(n * g), tau(n * s), out(n * (1 + s + g + s));
  if (fread(rho.data(), 8, rho.size(), stdin) != rho.size()) return 3;
  if (fread(sigma.data(), 8, sigma.size(), stdin) != sigma.size()) return 3;
  if (fread(tau.data(), 8, tau.size(), stdin) != tau.size()) return 3;
  if (xc_family_of(id) < 0) return 4;
  for (size_t i = 0; i < n; ++i) {
    XCPoint o;
    if (!pol) {
      if (fam == 0) xc_eval_unpolarized<0>(1, &id, rho[i], sigma[i], tau[i], o);
      else if (fam == 1) xc_eval_unpolarized<1>(1, &id, rho[i], sigma[i], tau[i], o);
      else xc_eval_unpolarized<2>(1, &id, rho[i], sigma[i], tau[i], o);
    } else {
      const double *r = &rho[2 * i], *sg = &sigma[3 * i], *t = &tau[2 * i];
      if (fam == 0) xc_eval_polarized<0>(1, &id, r[0], r[1], sg[0], sg[1], sg[2], t[0], t[1], o);
      else if (fam == 1) xc_eval_polarized<1>(1, &id, r[0], r[1], sg[0], sg[1], sg[2], t[0], t[1], o);
      else xc_eval_polarized<2>(1, &id, r[0], r[1], sg[0], sg[1], sg[2], t[0], t[1], o);
    }
    double *q = &out[i * (1 + s + g + s)];
    *q++ = o.exc;
    for (int k = 0; k < s; ++k) *q++ = o.vrho[k];
    for (int k = 0; k < g; ++k) *q++ = o.vsigma[k];
    for (int k = 0; k < s; ++k) *q++ = o.vtau[k];
  }
  fwrite(out.data(), 8, out.size(), stdout);
  return 0;
}
