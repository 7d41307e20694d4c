// oracle/src/dual.h — CPU ORACLE (test infrastructure): forward-mode dual numbers with N directional
// derivatives, so each functional is written ONCE as an energy density and every libxc-style
// potential (vrho, vsigma, vtau) falls out by the chain rule instead of hand-derived formulas.
#pragma once
#include <cmath>

namespace orc {

template <int N> struct Dual {
  double v;
  double d[N];
  Dual() : v(0) {
    for (int i = 0; i < N; ++i) d[i] = 0;
  }
  Dual(double x) : v(x) { // NOLINT (implicit on purpose)
    for (int i = 0; i < N; ++i) d[i] = 0;
  }
  static Dual var(double x, int k) {
    Dual r(x);
    r.d[k] = 1.0;
    return r;
  }
};

template <int N> inline Dual<N> chain(const Dual<N> &a, double f, double df) {
  Dual<N> r;
  r.v = f;
  for (int i = 0; i < N; ++i) r.d[i] = df * a.d[i];
  return r;
}

template <int N> inline Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a.v + b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> inline Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a.v - b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> inline Dual<N> operator-(const Dual<N> &a) {
  Dual<N> r;
  r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> inline Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> inline Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) {
  Dual<N> r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> inline Dual<N> operator+(const Dual<N> &a, double b) { return a + Dual<N>(b); }
template <int N> inline Dual<N> operator+(double a, const Dual<N> &b) { return Dual<N>(a) + b; }
template <int N> inline Dual<N> operator-(const Dual<N> &a, double b) { return a - Dual<N>(b); }
template <int N> inline Dual<N> operator-(double a, const Dual<N> &b) { return Dual<N>(a) - b; }
template <int N> inline Dual<N> operator*(const Dual<N> &a, double b) { return chain(a, a.v * b, b); }
template <int N> inline Dual<N> operator*(double a, const Dual<N> &b) { return chain(b, a * b.v, a); }
template <int N> inline Dual<N> operator/(const Dual<N> &a, double b) { return chain(a, a.v / b, 1.0 / b); }
template <int N> inline Dual<N> operator/(double a, const Dual<N> &b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N> &operator+=(Dual<N> &a, const Dual<N> &b) { return a = a + b; }
template <int N> inline Dual<N> &operator-=(Dual<N> &a, const Dual<N> &b) { return a = a - b; }
template <int N> inline Dual<N> &operator*=(Dual<N> &a, const Dual<N> &b) { return a = a * b; }

template <int N> inline bool operator<(const Dual<N> &a, double b) { return a.v < b; }
template <int N> inline bool operator>(const Dual<N> &a, double b) { return a.v > b; }
template <int N> inline bool operator<=(const Dual<N> &a, double b) { return a.v <= b; }
template <int N> inline bool operator>=(const Dual<N> &a, double b) { return a.v >= b; }
template <int N> inline bool operator<(const Dual<N> &a, const Dual<N> &b) { return a.v < b.v; }
template <int N> inline bool operator>(const Dual<N> &a, const Dual<N> &b) { return a.v > b.v; }

template <int N> inline Dual<N> sqrt(const Dual<N> &a) {
  const double s = std::sqrt(a.v);
  return chain(a, s, 0.5 / s);
}
template <int N> inline Dual<N> cbrt(const Dual<N> &a) {
  const double c = std::cbrt(a.v);
  return chain(a, c, c / (3.0 * a.v));
}
template <int N> inline Dual<N> exp(const Dual<N> &a) {
  const double e = std::exp(a.v);
  return chain(a, e, e);
}
template <int N> inline Dual<N> log(const Dual<N> &a) { return chain(a, std::log(a.v), 1.0 / a.v); }
template <int N> inline Dual<N> log1p(const Dual<N> &a) { return chain(a, std::log1p(a.v), 1.0 / (1.0 + a.v)); }
template <int N> inline Dual<N> expm1(const Dual<N> &a) { return chain(a, std::expm1(a.v), std::exp(a.v)); }
template <int N> inline Dual<N> atan(const Dual<N> &a) { return chain(a, std::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int N> inline Dual<N> asinh(const Dual<N> &a) {
  return chain(a, std::asinh(a.v), 1.0 / std::sqrt(1.0 + a.v * a.v));
}
template <int N> inline Dual<N> erf(const Dual<N> &a) {
  return chain(a, std::erf(a.v), 1.1283791670955126 * std::exp(-a.v * a.v)); // 2/sqrt(pi)
}
template <int N> inline Dual<N> pow(const Dual<N> &a, double p) {
  const double f = std::pow(a.v, p);
  return chain(a, f, p * std::pow(a.v, p - 1.0));
}

inline double sqrt(double x) { return std::sqrt(x); }
inline double cbrt(double x) { return std::cbrt(x); }
inline double exp(double x) { return std::exp(x); }
inline double log(double x) { return std::log(x); }
inline double log1p(double x) { return std::log1p(x); }
inline double expm1(double x) { return std::expm1(x); }
inline double atan(double x) { return std::atan(x); }
inline double asinh(double x) { return std::asinh(x); }
inline double erf(double x) { return std::erf(x); }
inline double pow(double x, double p) { return std::pow(x, p); }
inline double value(double x) { return x; }
template <int N> inline double value(const Dual<N> &x) { return x.v; }

} // namespace orc
