// oracle/src/oracle_internal.h — CPU ORACLE internals (test infrastructure, not product code).
#pragma once
#include "oracle.h"
#include <array>
#include <vector>

#define ORC_MAX_L 8

namespace orc {

int ncart(int l);
int nsph(int l);
void cart_powers(int l, std::vector<std::array<int, 3>> &out);
const std::vector<double> &c2s_matrix(int l);
void collocate_shell(int l, int pure, size_t npts, const double *xyz, int nprim, const double *coefs,
                     const double *exps, const double *center, int deriv, double *phi, double *phi_x,
                     double *phi_y, double *phi_z);

// value, gradient and xx, xy, xz, yy, yz, zz of one shell, component-major, no common_fac (oracle/src/gradient.cpp)
void collocate_shell_deriv2(int l, int pure, size_t npts, const double *xyz, int nprim, const double *coefs,
                            const double *exps, const double *center, double *out[10]);

struct Basis {
  explicit Basis(const orc_basis_t *b);
  orc_basis_t raw;
  std::vector<int> first_bf;
  int nbf;
  int max_shell_size;
};

const std::vector<double> *lebedev_table(int n); // rows of (x,y,z,w)

// C = alpha * op(A) * op(B) + beta * C, column-major; binds OpenBLAS when orc_set_blas succeeded.
void dgemm(char transa, char transb, int m, int n, int k, double alpha, const double *A, int lda,
           const double *B, int ldb, double beta, double *C, int ldc);

// pointwise functional evaluation on libxc-layout arrays, accumulating (+=) into the outputs
int xc_eval_add(int func_id, int polarized, size_t n, const double *rho, const double *sigma, const double *tau,
                double *exc, double *vrho, double *vsigma, double *vtau);

} // namespace orc
