// Test driver for adapter/occ_dft_b200.h: reads a case written by tests/test_adapter_cpp.py, runs
// XCBuild::compute_K_dft through the C ABI and writes K, Exc, N back.  With --expect-error it checks that a
// failure inside the library surfaces as std::runtime_error (the reference's error convention).
#include "../../adapter/occ_dft_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

template <class T> static std::vector<T> read_vec(std::ifstream &f) {
  long long n = 0;
  f.read(reinterpret_cast<char *>(&n), sizeof(n));
  std::vector<T> v(static_cast<size_t>(n));
  f.read(reinterpret_cast<char *>(v.data()), sizeof(T) * v.size());
  return v;
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const bool expect_error = argc > 3 && std::strcmp(argv[3], "--expect-error") == 0;
  // --devices a,b,c : one leaf shard per listed GPU, driven from this process (occ_xc_multi_*)
  std::vector<int> devices;
  for (int a = 3; a + 1 < argc; ++a)
    if (std::strcmp(argv[a], "--devices") == 0)
      for (const char *p = argv[a + 1]; *p;) {
        devices.push_back(std::atoi(p));
        while (*p && *p != ',') ++p;
        if (*p == ',') ++p;
      }
  bool want_gradient = false, want_nlc = false;
  for (int a = 3; a < argc; ++a) {
    if (std::strcmp(argv[a], "--gradient") == 0) want_gradient = true; // XCBuild::compute_xc_gradient appended to the output
    if (std::strcmp(argv[a], "--nlc") == 0) want_nlc = true;           // NonLocalCorrelation: energy, gradient, Vxc instead
  }
  std::ifstream f(argv[1], std::ios::binary);
  auto meta = read_vec<int>(f); // pure, nspin, nfunc ids...
  occ_b200::BasisArrays basis;
  basis.pure = meta[0] != 0;
  const int nspin = meta[1];
  std::vector<int> func_ids(meta.begin() + 2, meta.end());
  auto l = read_vec<int>(f), nprim = read_vec<int>(f), atom = read_vec<int>(f);
  auto origin = read_vec<double>(f), exps = read_vec<double>(f), coefs = read_vec<double>(f);
  size_t off = 0;
  for (size_t s = 0; s < l.size(); ++s) {
    basis.add_shell(l[s], &origin[3 * s], nprim[s], &exps[off], &coefs[off], atom[s]);
    off += nprim[s];
  }
  auto pts = read_vec<double>(f), w = read_vec<double>(f);
  auto leaf_raw = read_vec<double>(f); // offset, count, cx, cy, cz, radius per leaf
  std::vector<occ_xc_leaf_t> leaves(leaf_raw.size() / 6);
  for (size_t i = 0; i < leaves.size(); ++i) {
    leaves[i].offset = static_cast<size_t>(leaf_raw[6 * i]);
    leaves[i].count = static_cast<size_t>(leaf_raw[6 * i + 1]);
    for (int d = 0; d < 3; ++d) leaves[i].center[d] = leaf_raw[6 * i + 2 + d];
    leaves[i].radius = leaf_raw[6 * i + 5];
  }
  auto D = read_vec<double>(f);     // mo.D exactly as the reference stores it (NOT doubled)
  auto Cocc = read_vec<double>(f);  // mo.Cocc: nbf x nocc_alpha [, nbf x nocc_beta]; may be empty
  auto nocc = read_vec<int>(f);     // {nocc_alpha, nocc_beta} or empty
  if (want_nlc) { // the case's grid is the NLC grid: + sort permutation and atom-grid prefix sums
    auto perm64 = read_vec<long long>(f), off64 = read_vec<long long>(f);
    std::vector<size_t> perm(perm64.begin(), perm64.end()), offs(off64.begin(), off64.end());
    try {
      occ_b200::NonLocalCorrelation nlc(basis, w.size(), pts.data(), w.data(), leaves, perm, offs, true, -1, devices);
      const size_t nbf = static_cast<size_t>(basis.nbf()), natoms = offs.size() - 1;
      std::vector<double> Vxc(nbf * nbf), g(3 * natoms);
      const double e = nlc.vv10(D.data(), Vxc.data());
      const double eg = nlc.compute_gradient(D.data(), g.data());
      std::ofstream o(argv[2], std::ios::binary);
      const double sc[2] = {e, eg};
      o.write(reinterpret_cast<const char *>(sc), sizeof(sc));
      o.write(reinterpret_cast<const char *>(g.data()), sizeof(double) * g.size());
      o.write(reinterpret_cast<const char *>(Vxc.data()), sizeof(double) * Vxc.size());
      std::printf("ok nlc e=%.12f\n", e);
      return 0;
    } catch (const std::runtime_error &e) {
      std::printf("runtime_error: %s\n", e.what());
      return 1;
    }
  }
  try {
    occ_b200::XCBuild xc(basis, w.size(), pts.data(), w.data(), leaves, func_ids,
                         nspin == 2 ? occ_b200::SpinorbitalKind::Unrestricted : occ_b200::SpinorbitalKind::Restricted,
                         1e-10, 1e-10, -1, 0, 1, nocc.empty() ? -1 : nocc[0], nocc.empty() ? -1 : nocc[1], false, devices,
                         want_gradient);
    std::vector<double> K(D.size());
    xc.compute_K_dft(D.data(), K.data());
    if (!nocc.empty()) { // the factorised entry (with D and with the density formed on the device) must agree
      std::vector<double> K2(D.size()), K3(D.size());
      xc.compute_K_dft_mo(D.data(), Cocc.data(), K2.data());
      xc.compute_K_dft_mo(nullptr, Cocc.data(), K3.data());
      for (size_t i = 0; i < K.size(); ++i)
        if (std::fabs(K2[i] - K[i]) > 1e-10 || std::fabs(K3[i] - K[i]) > 1e-10) {
          std::printf("factorised K differs at %zu: %.15g %.15g %.15g\n", i, K[i], K2[i], K3[i]);
          return 1;
        }
    }
    const double *Kp = xc.compute_K_dft(D.data());   // page-locked variant must give the same matrix
    for (size_t i = 0; i < K.size(); ++i)
      if (std::fabs(Kp[i] - K[i]) > 1e-12) {
        std::printf("pinned K differs at %zu\n", i);
        return 1;
      }
    std::ofstream o(argv[2], std::ios::binary);
    const double sc[3] = {xc.exchange_correlation_energy(), xc.total_density_alpha(), xc.total_density_beta()};
    o.write(reinterpret_cast<const char *>(sc), sizeof(sc));
    o.write(reinterpret_cast<const char *>(K.data()), sizeof(double) * K.size());
    if (want_gradient) { // DFT::compute_xc_gradient(mo) without NLC, D = mo.D; with the orbitals when the case has them
      int natoms = 0;
      for (int a : atom) natoms = std::max(natoms, a + 1);
      std::vector<double> g(3 * static_cast<size_t>(natoms)), g2(g.size());
      xc.compute_xc_gradient(D.data(), g.data());
      if (!nocc.empty()) {
        xc.compute_xc_gradient_mo(D.data(), Cocc.data(), g2.data());
        for (size_t i = 0; i < g.size(); ++i)
          if (std::fabs(g2[i] - g[i]) > 1e-9) {
            std::printf("factorised gradient differs at %zu: %.15g %.15g\n", i, g[i], g2[i]);
            return 1;
          }
      }
      o.write(reinterpret_cast<const char *>(g.data()), sizeof(double) * g.size());
    }
    std::printf("ok deriv=%d exc=%.12f\n", xc.density_derivative(), sc[0]);
    return expect_error ? 1 : 0;
  } catch (const std::runtime_error &e) {
    std::printf("runtime_error: %s\n", e.what());
    return expect_error ? 0 : 1;
  }
}
