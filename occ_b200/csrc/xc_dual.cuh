// xc_dual.cuh — forward-mode dual numbers for the device-side exchange-correlation functionals.
// Each functional is one templated energy density e(rho_a, rho_b, sigma_aa, sigma_ab, sigma_bb, tau_a, tau_b);
// vrho / vsigma / vtau (what libxc's xc_*_exc_vxc returns to src/dft/functional.cpp:38-57 in the reference)
// are the N directional derivatives carried alongside the value.  N is the number of active inputs of the
// family/spin case (1..7), so LDA kernels do not pay for sigma/tau slots.
#pragma once
#include <math.h>

#define XC_HD __host__ __device__ __forceinline__

namespace xcb {

template <int N> struct Dual {
  double v;
  double d[N];
  XC_HD Dual() {}
  XC_HD Dual(double x) : v(x) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = 0.0;
  }
  XC_HD static Dual seed(double x, int k) {
    Dual r(x);
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (i == k) ? 1.0 : 0.0;
    return r;
  }
};

// f(a) with f' known
template <int N> XC_HD Dual<N> lift(const Dual<N> &a, double f, double df) {
  Dual<N> r;
  r.v = f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = df * a.d[i];
  return r;
}

template <int N> XC_HD Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> XC_HD Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> XC_HD Dual<N> operator-(const Dual<N> &a) {
  Dual<N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> XC_HD Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> XC_HD Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> XC_HD Dual<N> operator+(const Dual<N> &a, double b) {
  Dual<N> r = a;
  r.v += b;
  return r;
}
template <int N> XC_HD Dual<N> operator+(double a, const Dual<N> &b) { return b + a; }
template <int N> XC_HD Dual<N> operator-(const Dual<N> &a, double b) {
  Dual<N> r = a;
  r.v -= b;
  return r;
}
template <int N> XC_HD Dual<N> operator-(double a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -b.d[i];
  return r;
}
template <int N> XC_HD Dual<N> operator*(const Dual<N> &a, double b) { return lift(a, a.v * b, b); }
template <int N> XC_HD Dual<N> operator*(double a, const Dual<N> &b) { return lift(b, a * b.v, a); }
template <int N> XC_HD Dual<N> operator/(const Dual<N> &a, double b) {
  const double inv = 1.0 / b;
  return lift(a, a.v * inv, inv);
}
template <int N> XC_HD Dual<N> operator/(double a, const Dual<N> &b) {
  const double inv = 1.0 / b.v;
  const double q = a * inv;
  return lift(b, q, -q * inv);
}

template <int N> XC_HD Dual<N> sqrt(const Dual<N> &a) {
  const double s = ::sqrt(a.v);
  return lift(a, s, 0.5 / s);
}
template <int N> XC_HD Dual<N> cbrt(const Dual<N> &a) {
  const double c = ::cbrt(a.v);
  return lift(a, c, c / (3.0 * a.v));
}
template <int N> XC_HD Dual<N> exp(const Dual<N> &a) {
  const double e = ::exp(a.v);
  return lift(a, e, e);
}
template <int N> XC_HD Dual<N> expm1(const Dual<N> &a) { return lift(a, ::expm1(a.v), ::exp(a.v)); }
template <int N> XC_HD Dual<N> log(const Dual<N> &a) { return lift(a, ::log(a.v), 1.0 / a.v); }
template <int N> XC_HD Dual<N> atan(const Dual<N> &a) { return lift(a, ::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int N> XC_HD Dual<N> asinh(const Dual<N> &a) {
  return lift(a, ::asinh(a.v), 1.0 / ::sqrt(1.0 + a.v * a.v));
}
template <int N> XC_HD Dual<N> erf(const Dual<N> &a) {
  return lift(a, ::erf(a.v), 1.1283791670955126 * ::exp(-a.v * a.v)); // 2/sqrt(pi)
}
template <int N> XC_HD Dual<N> pow(const Dual<N> &a, double p) {
  const double f = ::pow(a.v, p);
  return lift(a, f, p * f / a.v);
}
template <int N> XC_HD Dual<N> sqr(const Dual<N> &a) { return lift(a, a.v * a.v, 2.0 * a.v); }
template <int N> XC_HD double val(const Dual<N> &a) { return a.v; }

XC_HD double sqr(double a) { return a * a; }
XC_HD double val(double a) { return a; }

} // namespace xcb
