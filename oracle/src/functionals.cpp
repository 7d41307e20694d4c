// oracle/src/functionals.cpp — CPU ORACLE (test infrastructure): the exchange-correlation functionals
// the reference obtains from libxc 6.2.2 (an un-vendored dependency, 3rdparty/CMakeLists.txt:215-234;
// call sites src/dft/functional.cpp:29-63).  libxc's source is not available offline, so each
// functional is restated from its publication in libxc's conventions:
//   eps = energy per particle;  vrho = d(rho*eps)/d rho_s;  vsigma = d(rho*eps)/d sigma (sigma = |grad rho|^2,
//   polarized order aa, ab, bb);  vtau = d(rho*eps)/d tau_s with tau = 1/2 sum |grad psi|^2.
// Every functional is ONE templated energy density F(ra, rb, saa, sab, sbb, ta, tb); potentials come
// from forward-mode AD (dual.h).  The unpolarized entry evaluates F(rho/2, rho/2, sigma/4, ...).
//
// Pinned by the reference's own tests: gga_x_pbe eps (tests/dft_tests.cpp:71-111); Slater+VWN3(=VWN5-P
// at zeta=0), PBE x+c, B88+LYP, r2SCAN x+c energies through tests/isolated_xc_tests.cpp:71-294.
// NOT pinned by any reference test ("parity unpinned"): VWN-RPA, the zeta != 0 interpolations of the
// three VWN variants, every polarized branch, and all potentials (those are checked by finite
// differences of F and by the polarized(zeta=0) == unpolarized identity in tests/test_oracle_functionals.py).
#include "dual.h"
#include "oracle.h"
#include "oracle_internal.h"
#include <cmath>

namespace orc {
namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kDensThreshold = 1e-15; // below this total density every output is zero (libxc dens_threshold)
constexpr double kSpinFloor = 1e-15;     // per-spin density floor in polarized mode (libxc floors at dens_threshold)
constexpr double kSigmaFloor = 1e-40;    // keeps sqrt(sigma) differentiable at sigma = 0
constexpr double kTauFloor = 1e-20;      // libxc tau_threshold

template <class T> inline T sq(const T &x) { return x * x; }

template <class T> struct Vars {
  T ra, rb, saa, sab, sbb, ta, tb;
};

// ------------------------------------------------------------------------------------------------
// LDA exchange (Dirac/Slater): e = -(3/4)(6/pi)^(1/3) sum_s rho_s^(4/3)
template <class T> T F_lda_x(const Vars<T> &v) {
  const double c = -0.75 * std::cbrt(6.0 / kPi);
  return c * (v.ra * cbrt(v.ra) + v.rb * cbrt(v.rb));
}

// ------------------------------------------------------------------------------------------------
// Vosko, Wilk, Nusair, Can. J. Phys. 58, 1200 (1980). Index 0: paramagnetic, 1: ferromagnetic, 2: spin stiffness
const double vwn_A[3] = {0.0310907, 0.01554535, -1.0 / (6.0 * kPi * kPi)};
const double vwn5_b[3] = {3.72744, 7.06042, 1.13107};
const double vwn5_c[3] = {12.9352, 18.0578, 13.0045};
const double vwn5_x0[3] = {-0.10498, -0.32500, -0.0047584};
const double rpa_b[3] = {13.0720, 20.1231, 1.06835};
const double rpa_c[3] = {42.7198, 101.578, 11.4813};
const double rpa_x0[3] = {-0.409286, -0.743294, -0.228344};

template <class T> T vwn_aux(double A, double b, double c, double x0, const T &rs) {
  const T x = sqrt(rs);
  const double Q = std::sqrt(4.0 * c - b * b);
  const T X = rs + b * x + c;
  const double X0 = x0 * x0 + b * x0 + c;
  const T at = atan(Q / (2.0 * x + b));
  return A * (log(rs / X) + (2.0 * b / Q) * at -
              (b * x0 / X0) * (log(sq(x - x0) / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
}

template <class T> T f_zeta(const T &z) { // spin-polarisation function f(zeta)
  const T p = 1.0 + z, m = 1.0 - z;
  return (p * cbrt(p) + m * cbrt(m) - 2.0) / (std::pow(2.0, 4.0 / 3.0) - 2.0);
}
const double kFpp0 = 4.0 / (9.0 * (std::cbrt(2.0) - 1.0)); // f''(0) = 1.7099209341613657

template <class T> void rs_zeta(const Vars<T> &v, T &rho, T &rs, T &zeta) {
  rho = v.ra + v.rb;
  rs = cbrt(3.0 / (4.0 * kPi * rho));
  zeta = (v.ra - v.rb) / rho;
}

// variant: 0 = VWN5 (lda_c_vwn), 1 = VWN-RPA (lda_c_vwn_rpa), 2 = VWN3 (lda_c_vwn_3)
template <class T> T eps_vwn(int variant, const T &rs, const T &z) {
  const T fz = f_zeta(z);
  const T z4 = sq(sq(z));
  if (variant == 1) {
    const T eP = vwn_aux(vwn_A[0], rpa_b[0], rpa_c[0], rpa_x0[0], rs);
    const T eF = vwn_aux(vwn_A[1], rpa_b[1], rpa_c[1], rpa_x0[1], rs);
    return eP * (1.0 - fz) + eF * fz;
  }
  const T eP = vwn_aux(vwn_A[0], vwn5_b[0], vwn5_c[0], vwn5_x0[0], rs);
  const T eF = vwn_aux(vwn_A[1], vwn5_b[1], vwn5_c[1], vwn5_x0[1], rs);
  const T dmc = eF - eP;
  if (variant == 0) {
    const T ac = vwn_aux(vwn_A[2], vwn5_b[2], vwn5_c[2], vwn5_x0[2], rs);
    return eP + ac * fz * (1.0 - z4) / kFpp0 + dmc * fz * z4;
  }
  // VWN3: RPA spin stiffness scaled by the Monte-Carlo/RPA ratio of the para->ferro difference
  const T ePr = vwn_aux(vwn_A[0], rpa_b[0], rpa_c[0], rpa_x0[0], rs);
  const T eFr = vwn_aux(vwn_A[1], rpa_b[1], rpa_c[1], rpa_x0[1], rs);
  const T acr = vwn_aux(vwn_A[2], rpa_b[2], rpa_c[2], rpa_x0[2], rs);
  const T drpa = eFr - ePr;
  return eP + (dmc / drpa) * acr * fz * (1.0 - z4) / kFpp0 + dmc * fz * z4;
}
template <class T, int VARIANT> T F_vwn(const Vars<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  return rho * eps_vwn(VARIANT, rs, z);
}

// ------------------------------------------------------------------------------------------------
// Becke, Phys. Rev. A 38, 3098 (1988): beta = 0.0042, gamma = 6
template <class T> T b88_spin(const T &r, const T &s) {
  const double Cx = 0.375 * std::cbrt(3.0 / kPi) * std::pow(4.0, 2.0 / 3.0); // 0.9305257363491
  const double beta = 0.0042, gamma = 6.0;
  const T r43 = r * cbrt(r);
  const T x = sqrt(s) / r43;
  return -r43 * (Cx + beta * sq(x) / (1.0 + gamma * beta * x * asinh(x)));
}
template <class T> T F_b88(const Vars<T> &v) { return b88_spin(v.ra, v.saa) + b88_spin(v.rb, v.sbb); }
// Becke, J. Chem. Phys. 85, 7184 (1986) ("B86b", libxc gga_x_b86_mgc, id 105; reference methods b86bpbe / b86bpbeh):
// E_x = E_x^LSDA - beta sum_s rho_s^(4/3) x_s^2 / (1 + gamma x_s^2)^(4/5), beta = 0.00375, gamma = 0.007.
// "parity unpinned": the reference holds no number for it; the two constants are the published ones.
template <class T> T b86b_spin(const T &r, const T &s) {
  const double Cx = 0.375 * std::cbrt(3.0 / kPi) * std::pow(4.0, 2.0 / 3.0);
  const T r43 = r * cbrt(r);
  const T x2 = s / sq(r43);
  return -r43 * (Cx + 0.00375 * x2 / pow(1.0 + 0.007 * x2, 0.8));
}
template <class T> T F_b86b(const Vars<T> &v) { return b86b_spin(v.ra, v.saa) + b86b_spin(v.rb, v.sbb); }

// ------------------------------------------------------------------------------------------------
// Lee, Yang, Parr, Phys. Rev. B 37, 785 (1988) in the gradient-only form of Miehlich, Savin, Stoll,
// Preuss, Chem. Phys. Lett. 157, 200 (1989).
template <class T> T F_lyp(const Vars<T> &v) {
  const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
  const double CF = 0.3 * std::pow(3.0 * kPi * kPi, 2.0 / 3.0);
  const T rho = v.ra + v.rb;
  const T rm13 = 1.0 / cbrt(rho);
  const T den = 1.0 + d * rm13;
  const T omega = exp(-c * rm13) / den * pow(rho, -11.0 / 3.0);
  const T delta = c * rm13 + d * rm13 / den;
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T ra83 = sq(v.ra) * cbrt(sq(v.ra)), rb83 = sq(v.rb) * cbrt(sq(v.rb));
  const T t1 = -4.0 * a / den * v.ra * v.rb / rho;
  const T brack = std::pow(2.0, 11.0 / 3.0) * CF * (ra83 + rb83) + (47.0 / 18.0 - 7.0 * delta / 18.0) * stot -
                  (2.5 - delta / 18.0) * (v.saa + v.sbb) -
                  (delta - 11.0) / 9.0 * (v.ra / rho * v.saa + v.rb / rho * v.sbb);
  const T t2 = -a * b * omega *
               (v.ra * v.rb * brack - (2.0 / 3.0) * sq(rho) * stot + ((2.0 / 3.0) * sq(rho) - sq(v.ra)) * v.sbb +
                ((2.0 / 3.0) * sq(rho) - sq(v.rb)) * v.saa);
  return t1 + t2;
}

// ------------------------------------------------------------------------------------------------
// Perdew, Burke, Ernzerhof, Phys. Rev. Lett. 77, 3865 (1996)
template <class T> T ex_unif(const T &n) { return -0.75 * std::cbrt(3.0 / kPi) * n * cbrt(n); }
template <class T> T s2_of(const T &n, const T &sigma) { // reduced gradient squared
  const double c = 4.0 * std::pow(3.0 * kPi * kPi, 2.0 / 3.0);
  return sigma / (c * sq(n) * cbrt(sq(n)));
}
template <class T> T pbe_x_n(const T &n, const T &sigma) {
  const double kappa = 0.804, mu = 0.2195149727645171;
  const T s2 = s2_of(n, sigma);
  return ex_unif(n) * (1.0 + kappa - kappa / (1.0 + mu * s2 / kappa));
}
template <class T> T F_pbe_x(const Vars<T> &v) { // spin scaling E[ra,rb] = (E[2ra] + E[2rb])/2
  return 0.5 * (pbe_x_n(2.0 * v.ra, 4.0 * v.saa) + pbe_x_n(2.0 * v.rb, 4.0 * v.sbb));
}

// Perdew & Wang, Phys. Rev. B 45, 13244 (1992), "modified" constants (more digits) as used by PBE/SCAN codes
struct PwSet {
  double A, a1, b1, b2, b3, b4;
};
const PwSet pw_mod[3] = {{0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294},
                         {0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517},
                         {0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671}};
template <class T> T pw_G(const PwSet &p, const T &rs) {
  const T srs = sqrt(rs);
  const T q1 = 2.0 * p.A * (p.b1 * srs + p.b2 * rs + p.b3 * rs * srs + p.b4 * sq(rs));
  return -2.0 * p.A * (1.0 + p.a1 * rs) * log(1.0 + 1.0 / q1);
}
template <class T> T pw_dG(const PwSet &p, const T &rs) { // d pw_G / d rs
  const T srs = sqrt(rs);
  const T q1 = 2.0 * p.A * (p.b1 * srs + p.b2 * rs + p.b3 * rs * srs + p.b4 * sq(rs));
  const T dq1 = p.A * (p.b1 / srs + 2.0 * p.b2 + 3.0 * p.b3 * srs + 4.0 * p.b4 * rs);
  return -2.0 * p.A * p.a1 * log(1.0 + 1.0 / q1) + 2.0 * p.A * (1.0 + p.a1 * rs) * dq1 / (sq(q1) + q1);
}
template <class T> T eps_pw(const T &rs, const T &z) {
  const T e0 = pw_G(pw_mod[0], rs), e1 = pw_G(pw_mod[1], rs), mac = pw_G(pw_mod[2], rs); // mac = -alpha_c
  const T fz = f_zeta(z), z4 = sq(sq(z));
  return e0 - mac * fz * (1.0 - z4) / kFpp0 + (e1 - e0) * fz * z4;
}
template <class T> T deps_pw_drs(const T &rs, const T &z) {
  const T e0 = pw_dG(pw_mod[0], rs), e1 = pw_dG(pw_mod[1], rs), mac = pw_dG(pw_mod[2], rs);
  const T fz = f_zeta(z), z4 = sq(sq(z));
  return e0 - mac * fz * (1.0 - z4) / kFpp0 + (e1 - e0) * fz * z4;
}
// 1 +- zeta as 2 rho_s / rho (not 1 +- (ra - rb) / rho): keeps the relative accuracy of a spin density at its floor,
// which matters where derivatives behave like (1 - zeta)^(-1/3)
template <class T> void zeta_plus_minus(const Vars<T> &v, T &zp, T &zm) {
  const T rho = v.ra + v.rb;
  zp = 2.0 * v.ra / rho;
  zm = 2.0 * v.rb / rho;
}
template <class T> T phi_zeta(const Vars<T> &v) {
  T zp, zm;
  zeta_plus_minus(v, zp, zm);
  return 0.5 * (cbrt(sq(zp)) + cbrt(sq(zm)));
}

template <class T> T F_pbe_c(const Vars<T> &v) {
  const double beta = 0.06672455060314922, gamma = (1.0 - std::log(2.0)) / (kPi * kPi);
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T ec = eps_pw(rs, z);
  const T phi = phi_zeta(v);
  const T phi3 = phi * sq(phi);
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  // t^2 = sigma / (4 phi^2 ks^2 rho^2), ks^2 = 4 kF / pi, kF = (3 pi^2 rho)^(1/3)
  const T kF = cbrt(3.0 * kPi * kPi * rho);
  const T t2 = stot * kPi / (16.0 * sq(phi) * kF * sq(rho));
  const T A = (beta / gamma) / expm1(-ec / (gamma * phi3));
  const T At2 = A * t2;
  const T H = gamma * phi3 * log(1.0 + (beta / gamma) * t2 * (1.0 + At2) / (1.0 + At2 + sq(At2)));
  return rho * (ec + H);
}

// ------------------------------------------------------------------------------------------------
// r2SCAN: Furness, Kaplan, Ning, Perdew, Sun, J. Phys. Chem. Lett. 11, 8208 (2020) (+ erratum)
const double r2_eta = 0.001, r2_dp2 = 0.361;
const double r2_cx[8] = {1.0,           -0.667,         -0.4445555,     -0.663086601049,
                         1.451297044490, -0.887998041597, 0.234528941479, -0.023185843322};
const double r2_cc[8] = {1.0,           -0.64,          -0.4352,        -1.535685604549,
                         3.061560252175, -1.915710236206, 0.516884468372, -0.051848879792};

template <class T> T r2_switch(const T &alpha, const double *poly, double c1, double c2, double dd) {
  if (alpha < 0.0) return exp(-c1 * alpha / (1.0 - alpha));
  if (alpha <= 2.5) {
    T r = poly[7];
    for (int i = 6; i >= 0; --i) r = r * alpha + poly[i];
    return r;
  }
  return -dd * exp(c2 / (1.0 - alpha));
}

template <class T> T r2scan_x_n(const T &n, const T &sigma_in, const T &tau) {
  const double k0 = 0.174, k1 = 0.065, mu = 10.0 / 81.0, a1 = 4.9479;
  const double c1x = 0.667, c2x = 0.8, dx = 1.24;
  double sum_icx = 0.0;
  for (int i = 1; i < 8; ++i) sum_icx += i * r2_cx[i];
  const double C2 = sum_icx * k0; // = -sum i c_xi (1 - h0x) = -0.162742
  const double Ceta = 20.0 / 27.0 + r2_eta * 5.0 / 3.0;
  const T &sigma = sigma_in; // already bounded by 8 n tau at the input stage (xc_eval_add)
  const T p = s2_of(n, sigma);
  const T tauW = sigma / (8.0 * n);
  const T tauU = 0.3 * std::pow(3.0 * kPi * kPi, 2.0 / 3.0) * n * cbrt(sq(n));
  const T alpha = (tau - tauW) / (tauU + r2_eta * tauW);
  const T x = (Ceta * C2 * exp(-sq(p) / std::pow(r2_dp2, 4)) + mu) * p;
  const T h1 = 1.0 + k1 - k1 / (1.0 + x / k1);
  const double h0 = 1.0 + k0;
  const T fx = r2_switch(alpha, r2_cx, c1x, c2x, dx);
  const T gx = 1.0 - exp(-a1 / sqrt(sqrt(p)));
  return ex_unif(n) * (h1 + fx * (h0 - h1)) * gx;
}
template <class T> T F_r2scan_x(const Vars<T> &v) {
  return 0.5 * (r2scan_x_n(2.0 * v.ra, 4.0 * v.saa, 2.0 * v.ta) + r2scan_x_n(2.0 * v.rb, 4.0 * v.sbb, 2.0 * v.tb));
}

template <class T> T F_r2scan_c(const Vars<T> &v) {
  const double b1c = 0.0285764, b2c = 0.0889, b3c = 0.125541, chi_inf = 0.128026;
  const double gamma = (1.0 - std::log(2.0)) / (kPi * kPi);
  const double c1c = 0.64, c2c = 1.5, dc = 0.7;
  double dfc2 = 0.0;
  for (int i = 1; i < 8; ++i) dfc2 += i * r2_cc[i];
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T tau = v.ta + v.tb;
  const T p = s2_of(rho, stot);
  T zp, zm;
  zeta_plus_minus(v, zp, zm);
  const T ds = 0.5 * (zp * cbrt(sq(zp)) + zm * cbrt(sq(zm)));   // ((1+z)^(5/3) + (1-z)^(5/3))/2
  const T dxz = 0.5 * (zp * cbrt(zp) + zm * cbrt(zm));           // ((1+z)^(4/3) + (1-z)^(4/3))/2
  const T phi = phi_zeta(v);
  const T phi3 = phi * sq(phi);
  // iso-orbital indicator of the total density
  const T tauW = stot / (8.0 * rho);
  const T tauU = 0.3 * std::pow(3.0 * kPi * kPi, 2.0 / 3.0) * rho * cbrt(sq(rho)) * ds;
  const T alpha = (tau - tauW) / (tauU + r2_eta * tauW);
  const T fc = r2_switch(alpha, r2_cc, c1c, c2c, dc);

  // alpha = 0 limit
  const T srs = sqrt(rs);
  const T den0 = 1.0 + b2c * srs + b3c * rs;
  const T e_lda0 = -b1c / den0;
  const T de_lda0 = b1c * (0.5 * b2c / srs + b3c) / sq(den0);
  const T z2 = sq(z);
  const T z12 = sq(sq(z2)) * sq(z2);
  const T gc = (1.0 - 2.3631 * (dxz - 1.0)) * (1.0 - z12);
  const T w0 = expm1(-e_lda0 / b1c);
  const T ginf = 1.0 / sqrt(sqrt(1.0 + 4.0 * chi_inf * p));
  const T H0 = b1c * log(1.0 + w0 * (1.0 - ginf));
  const T eps0 = (e_lda0 + H0) * gc;

  // alpha = 1 limit
  const T e_lsda1 = eps_pw(rs, z);
  const T de_lsda1 = deps_pw_drs(rs, z);
  const T beta = 0.066725 * (1.0 + 0.1 * rs) / (1.0 + 0.1778 * rs);
  const T t2 = std::pow(3.0 * kPi * kPi / 16.0, 2.0 / 3.0) * p / (sq(phi) * rs);
  const T w1 = expm1(-e_lsda1 / (gamma * phi3));
  const T y = beta * t2 / (gamma * w1);
  const T e_lsda0 = e_lda0 * gc;
  const T de_lsda0 = de_lda0 * gc;
  const T dy = dfc2 / (27.0 * gamma * ds * phi3 * w1) *
               (20.0 * rs * (de_lsda0 - de_lsda1) - 45.0 * r2_eta * (e_lsda0 - e_lsda1)) * p *
               exp(-sq(p) / std::pow(r2_dp2, 4));
  const T g1 = 1.0 / sqrt(sqrt(1.0 + 4.0 * (y - dy)));
  const T H1 = gamma * phi3 * log(1.0 + w1 * (1.0 - g1));
  const T eps1 = e_lsda1 + H1;

  return rho * (eps1 + fc * (eps0 - eps1));
}

// ------------------------------------------------------------------------------------------------
// Perdew-Wang 1992 LSDA correlation as libxc's lda_c_pw (original 1992 constants) / lda_c_pw_mod
const PwSet pw_orig[3] = {{0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294},
                          {0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517},
                          {0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671}};
template <class T, int MOD> T F_pw(const Vars<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  if (MOD) return rho * eps_pw(rs, z);
  const T e0 = pw_G(pw_orig[0], rs), e1 = pw_G(pw_orig[1], rs), mac = pw_G(pw_orig[2], rs);
  const T fz = f_zeta(z), z4 = sq(sq(z));
  return rho * (e0 - mac * fz * (1.0 - z4) / 1.709921 + (e1 - e0) * fz * z4);
}

// ------------------------------------------------------------------------------------------------
// TPSS: Tao, Perdew, Staroverov, Scuseria, Phys. Rev. Lett. 91, 146401 (2003).
// Pinned by the reference's own known-answer test (tests/dft_tests.cpp:138-191: eps of x and c, and
// vrho / vsigma / vtau of c at four points, 1e-6).
template <class T> T tpss_x_n(const T &n, const T &sigma_in, const T &tau) {
  const double kappa = 0.804, b = 0.40, c = 1.59096, e = 1.537, mu = 0.21951;
  const double se = std::sqrt(e);
  const T &sigma = sigma_in; // z = tau_W / tau <= 1 through the input-stage bound sigma <= 8 n tau
  const T p = s2_of(n, sigma);
  const T z = sigma / (8.0 * n * tau);
  const T tauU = 0.3 * std::pow(3.0 * kPi * kPi, 2.0 / 3.0) * n * cbrt(sq(n));
  const T alpha = (tau - sigma / (8.0 * n)) / tauU;
  const T qb = 0.45 * (alpha - 1.0) / sqrt(1.0 + b * alpha * (alpha - 1.0)) + (2.0 / 3.0) * p;
  const T z2 = sq(z);
  const T num = (10.0 / 81.0 + c * z2 / sq(1.0 + z2)) * p + (146.0 / 2025.0) * sq(qb) -
                (73.0 / 405.0) * qb * sqrt(0.5 * sq(0.6 * z) + 0.5 * sq(p)) + sq(10.0 / 81.0) / kappa * sq(p) +
                2.0 * se * (10.0 / 81.0) * sq(0.6 * z) + e * mu * p * sq(p);
  const T x = num / sq(1.0 + se * p);
  return ex_unif(n) * (1.0 + kappa - kappa / (1.0 + x / kappa));
}
template <class T> T F_tpss_x(const Vars<T> &v) {
  return 0.5 * (tpss_x_n(2.0 * v.ra, 4.0 * v.saa, 2.0 * v.ta) + tpss_x_n(2.0 * v.rb, 4.0 * v.sbb, 2.0 * v.tb));
}
// PBE correlation energy per particle of a fully spin-polarised density n with gradient square sigma
template <class T> T pbe_c_eps_ferro(const T &n, const T &sigma) {
  const double beta = 0.06672455060314922, gamma = (1.0 - std::log(2.0)) / (kPi * kPi);
  const double phi = 0.5 * std::cbrt(4.0), phi3 = phi * phi * phi; // phi(zeta = 1) = 2^(-1/3)
  const T rs = cbrt(3.0 / (4.0 * kPi * n));
  const T ec = pw_G(pw_mod[1], rs);
  const T kF = cbrt(3.0 * kPi * kPi * n);
  const T t2 = sigma * kPi / (16.0 * phi * phi * kF * sq(n));
  const T A = (beta / gamma) / expm1(-ec / (gamma * phi3));
  const T At2 = A * t2;
  return ec + gamma * phi3 * log(1.0 + (beta / gamma) * t2 * (1.0 + At2) / (1.0 + At2 + sq(At2)));
}
template <class T> T F_tpss_c(const Vars<T> &v) {
  const double d = 2.8;
  const T rho = v.ra + v.rb;
  const T zeta = (v.ra - v.rb) / rho;
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T tau = v.ta + v.tb;
  const T z = stot / (8.0 * rho * tau);
  // C(zeta, xi): xi = |grad zeta| / (2 (3 pi^2 n)^(1/3)), |grad zeta|^2 = 4 (nb^2 saa - 2 na nb sab + na^2 sbb) / n^4
  const T zeta2 = sq(zeta);
  const T C0 = 0.53 + 0.87 * zeta2 + 0.50 * sq(zeta2) + 2.26 * zeta2 * sq(zeta2);
  const T gz2 = 4.0 * (sq(v.rb) * v.saa - 2.0 * v.ra * v.rb * v.sab + sq(v.ra) * v.sbb) / sq(sq(rho));
  const T xi2 = gz2 / (4.0 * cbrt(sq(3.0 * kPi * kPi * rho)));
  T zp, zm;
  zeta_plus_minus(v, zp, zm);
  const T inv43 = 0.5 * (1.0 / (zp * cbrt(zp)) + 1.0 / (zm * cbrt(zm)));
  const T C = C0 / sq(sq(1.0 + xi2 * inv43));
  const T e_pbe = F_pbe_c(v) / rho;
  T ea = pbe_c_eps_ferro(v.ra, v.saa), eb = pbe_c_eps_ferro(v.rb, v.sbb);
  if (value(ea) < value(e_pbe)) ea = e_pbe; // eps~_sigma = max(eps_PBE(n_sigma, 0), eps_PBE(na, nb))
  if (value(eb) < value(e_pbe)) eb = e_pbe;
  const T z2 = sq(z);
  const T pkzb = e_pbe * (1.0 + C * z2) - (1.0 + C) * z2 * (v.ra * ea + v.rb * eb) / rho;
  return rho * pkzb * (1.0 + d * pkzb * z2 * z);
}

// ------------------------------------------------------------------------------------------------
template <class T> T F_b3lyp(const Vars<T> &v) { // libxc hyb_gga_xc_b3lyp: a0 = 0.20, ax = 0.72, ac = 0.81
  return 0.08 * F_lda_x(v) + 0.72 * F_b88(v) + 0.19 * F_vwn<T, 1>(v) + 0.81 * F_lyp(v);
}

// libxc hyb_gga_xc_b3lyp5 (VWN5 in place of VWN-RPA), hyb_gga_xc_pbeh (PBE0: 0.75 PBE-x + PBE-c, exx 0.25),
// hyb_gga_xc_bhandh (0.5 LDA-x + LYP, exx 0.5), hyb_gga_xc_bhandhlyp (0.5 B88 + LYP, exx 0.5)
template <class T> T F_b3lyp5(const Vars<T> &v) {
  return 0.08 * F_lda_x(v) + 0.72 * F_b88(v) + 0.19 * F_vwn<T, 0>(v) + 0.81 * F_lyp(v);
}
template <class T> T F_pbeh(const Vars<T> &v) { return 0.75 * F_pbe_x(v) + F_pbe_c(v); }
template <class T> T F_bhandh(const Vars<T> &v) { return 0.5 * F_lda_x(v) + F_lyp(v); }
template <class T> T F_bhandhlyp(const Vars<T> &v) { return 0.5 * F_b88(v) + F_lyp(v); }

// ------------------------------------------------------------------------------------------------
// B97-type power-series GGAs with a short-range (erf-attenuated) LSDA exchange: wB97X of Chai & Head-Gordon,
// J. Chem. Phys. 128, 084106 (2008) [libxc hyb_gga_xc_wb97x, id 464; reference table src/dft/dft_method.cpp:35].
//   E_xc = sum_s e_x^LSDA(rho_s) F(a_s) g_x(x_s^2) + sum_s e_c,ss g_ss(x_s^2) + e_c,ab g_ab((x_a^2 + x_b^2)/2)
//   x_s = |grad rho_s| / rho_s^(4/3);  g(x^2) = sum_i c_i u^i,  u = gamma x^2 / (1 + gamma x^2),
//   gamma_x = 0.004, gamma_ss = 0.2, gamma_ab = 0.006 (Becke, J. Chem. Phys. 107, 8554 (1997));
//   e_c,ss / e_c,ab: Stoll partition of the PW92 correlation energy with the constants AS PRINTED in 1992 (libxc's
//   lda_c_pw set; the more-digit "modified" set moves the water/3-21G golden by 4.5e-6 Eh, see tests/test_isolated_xc.py);
//   F(a): attenuation of the LSDA exchange hole by erfc(omega r)/r, a_s = omega / (2 k_Fs), k_Fs = (6 pi^2 rho_s)^(1/3)
//   (Gill, Adamson, Pople, Mol. Phys. 88, 1005 (1996); eq. 7-8 of the wB97X paper), omega = 0.3.
// The exact-exchange part (100 % long range + 15.7706 % short range) is the Fock build's business
// (DFT::compute_JK -> range-separated K); xc_hyb_cam_coef: omega 0.3, alpha 1.0, beta -0.842294.
// Pinned by the reference's own isolated-XC golden -6.577307030585 (tests/isolated_xc_tests.cpp:296-319).
template <class T> T sr_attenuation(const T &a) {
  if (value(a) >= 1.0) { // asymptotic series (the closed form cancels to 1/(36 a^2) from O(a^4) terms)
    const T i2 = 1.0 / sq(a);
    return i2 * (1.0 / 36.0 + i2 * (-1.0 / 960.0 + i2 * (1.0 / 26880.0 + i2 * (-1.0 / 829440.0 + i2 * (1.0 / 28385280.0 +
           i2 * (-1.0 / 1073479680.0 + i2 * (1.0 / 44590694400.0 + i2 * (-1.0 / 2021444812800.0 +
           i2 * (1.0 / 99407521382400.0)))))))));
  }
  const T b = expm1(-0.25 / sq(a));
  const T c = 2.0 * sq(a) * b + 0.5;
  return 1.0 - (8.0 / 3.0) * a * (std::sqrt(kPi) * erf(0.5 / a) + 2.0 * a * (b - c));
}
template <class T> T b97_series(double gamma, const double *c, const T &x2) {
  const T u = gamma * x2 / (1.0 + gamma * x2);
  return c[0] + u * (c[1] + u * (c[2] + u * (c[3] + u * c[4])));
}
struct B97Set {
  double cx[5], css[5], cab[5], omega;
};
const B97Set kWB97X = {{0.842294, 0.726479, 1.04760, -5.70635, 13.2794},
                       {1.00000, -4.33879, 18.2308, -31.7430, 17.2901},
                       {1.00000, 2.37031, -11.3995, 6.58405, -3.78132},
                       0.3};
template <class T> T F_b97_family(const B97Set &P, const Vars<T> &v) {
  const double cx_lsda = -0.75 * std::cbrt(6.0 / kPi);
  const T r43a = v.ra * cbrt(v.ra), r43b = v.rb * cbrt(v.rb);
  const T x2a = v.saa / sq(r43a), x2b = v.sbb / sq(r43b);
  // exchange
  T ex = T(0.0);
  const T *r43[2] = {&r43a, &r43b};
  const T *x2[2] = {&x2a, &x2b};
  const T *rs_[2] = {&v.ra, &v.rb};
  for (int s = 0; s < 2; ++s) {
    T att = T(1.0);
    if (P.omega > 0.0) att = sr_attenuation(P.omega / (2.0 * cbrt(6.0 * kPi * kPi * *rs_[s])));
    ex = ex + cx_lsda * *r43[s] * att * b97_series(0.004, P.cx, *x2[s]);
  }
  // correlation, Stoll partition of PW92
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T ec_tot = F_pw<T, 0>(v);
  const T rsa = cbrt(3.0 / (4.0 * kPi * v.ra)), rsb = cbrt(3.0 / (4.0 * kPi * v.rb));
  const T ec_aa = v.ra * pw_G(pw_orig[1], rsa), ec_bb = v.rb * pw_G(pw_orig[1], rsb);
  const T ec_ab = ec_tot - ec_aa - ec_bb;
  return ex + ec_aa * b97_series(0.2, P.css, x2a) + ec_bb * b97_series(0.2, P.css, x2b) +
         ec_ab * b97_series(0.006, P.cab, 0.5 * (x2a + x2b));
}
template <class T> T F_wb97x(const Vars<T> &v) { return F_b97_family(kWB97X, v); }
// wB97X-V (Mardirossian & Head-Gordon, Phys. Chem. Chem. Phys. 16, 9904 (2014), Table 3; libxc hyb_gga_xc_wb97x_v, id 466;
// reference table share/methods/dft_methods.json "wb97x-v", dispersion "vv10" with b = 6.0, C = 0.01): seven linear
// coefficients on the same three series, omega = 0.3, 16.7 % short-range + 100 % long-range exact exchange
// (xc_hyb_cam_coef: alpha 1.0, beta -0.833).  Pinned by the reference's golden -6.539400300836
// (tests/isolated_xc_tests.cpp:340-364: the local part only, the VV10 energy is post-SCF).
const B97Set kWB97XV = {{0.833, 0.603, 1.194, 0.0, 0.0}, {0.556, -0.257, 0.0, 0.0, 0.0}, {1.219, -1.850, 0.0, 0.0, 0.0}, 0.3};
template <class T> T F_wb97x_v(const Vars<T> &v) { return F_b97_family(kWB97XV, v); }
// B97 itself (Becke, J. Chem. Phys. 107, 8554 (1997), Table: c_x, c_ss, c_ab; 19.43 % exact exchange; libxc
// hyb_gga_xc_b97, id 407; reference table src/dft/dft_method.cpp:18): full-range exchange (omega = 0), three terms.
// "parity unpinned" beyond the machinery it shares with wB97X (pinned): the nine coefficients are the published ones.
const B97Set kB97 = {{0.8094, 0.5073, 0.7481, 0.0, 0.0}, {0.1737, 2.3487, -2.4868, 0.0, 0.0},
                     {0.9454, 0.7471, -4.5961, 0.0, 0.0}, 0.0};
template <class T> T F_b97(const Vars<T> &v) { return F_b97_family(kB97, v); }

// ------------------------------------------------------------------------------------------------
// Perdew-Zunger 1981 (Phys. Rev. B 23, 5048, Appendix C): eps_c of the unpolarized / polarized gas, f(zeta) interpolation
// [libxc lda_c_pz, id 9]; local part of P86.
struct PzSet {
  double gamma, beta1, beta2, A, B, C, D;
};
const PzSet pz_u = {-0.1423, 1.0529, 0.3334, 0.0311, -0.048, 0.0020, -0.0116};
const PzSet pz_p = {-0.0843, 1.3981, 0.2611, 0.01555, -0.0269, 0.0007, -0.0048};
template <class T> T pz_eps(const PzSet &p, const T &rs) {
  if (value(rs) >= 1.0) return p.gamma / (1.0 + p.beta1 * sqrt(rs) + p.beta2 * rs);
  return p.A * log(rs) + p.B + p.C * rs * log(rs) + p.D * rs;
}
template <class T> T F_pz81(const Vars<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T eu = pz_eps(pz_u, rs), ep = pz_eps(pz_p, rs);
  return rho * (eu + f_zeta(z) * (ep - eu));
}
// Rasolt & Geldart, Phys. Rev. B 34, 1325 (1986): gradient coefficient C_c(rs) = C_xc(rs) - C_x, C_x = -0.001667
template <class T> T rg_Cc(const T &rs) {
  return 0.001667 + (0.002568 + 0.023266 * rs + 7.389e-6 * sq(rs)) /
                        (1.0 + 8.723 * rs + 0.472 * sq(rs) + 1.0e4 * 7.389e-6 * rs * sq(rs));
}
// Perdew, Phys. Rev. B 33, 8822 (1986), eqs. 4-9 [libxc gga_c_p86, id 132; reference method bp86,
// src/dft/dft_method.cpp:30]: E_c = E_c^PZ81 + int exp(-Phi) C(n) |grad n|^2 / (d n^(4/3)),
// Phi = 1.745 f~ (C(inf) / C(n)) |grad n| / n^(7/6), f~ = 0.11, d = 2^(1/3) sqrt(((1+z)/2)^(5/3) + ((1-z)/2)^(5/3)).
// "parity unpinned": no reference-held value; published formula.
template <class T> T F_p86(const Vars<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T sigma = v.saa + 2.0 * v.sab + v.sbb;
  const T C = rg_Cc(rs);
  const double Cinf = 0.001667 + 0.002568;
  const T n16 = sqrt(cbrt(rho));
  const T Phi = 1.745 * 0.11 * (Cinf / C) * sqrt(sigma) / (rho * n16);
  T zp, zm;
  zeta_plus_minus(v, zp, zm);
  const T a = 0.5 * zp, b = 0.5 * zm;
  const T d = std::cbrt(2.0) * sqrt(a * cbrt(sq(a)) + b * cbrt(sq(b)));
  return F_pz81(v) + exp(-Phi) * C * sigma / (d * rho * cbrt(rho));
}
// Perdew & Wang 1991 correlation (Perdew et al., Phys. Rev. B 46, 6671 (1992), eqs. 10-15; the form quoted in the PBE paper,
// Phys. Rev. Lett. 77, 3865, ref. 3) [libxc gga_c_pw91, id 134; component of hyb_gga_xc_b3pw91]:
//   eps_c = eps^PW92 + H0 + H1
//   H0 = (beta^2 / 2 alpha) phi^3 ln[1 + (2 alpha / beta) (t^2 + A t^4) / (1 + A t^2 + A^2 t^4)]
//   A  = (2 alpha / beta) / (exp(-2 alpha eps^PW92 / (phi^3 beta^2)) - 1)
//   H1 = nu [C_c(rs) - C_c(0) - 3 C_x / 7] phi^3 t^2 exp(-100 phi^4 (k_s^2 / k_F^2) t^2)
//   alpha = 0.09, beta = nu C_c(0), nu = (16 / pi)(3 pi^2)^(1/3), C_c(0) = 0.004235, C_x = -0.001667
// Pinned by E_c^PW91 of the hydrogen and helium atoms in Table I of the PBE paper (-0.0065, -0.0450 Eh), tests/test_oracle_pins.py.
template <class T> T F_pw91_c(const Vars<T> &v) {
  const double alpha = 0.09, Cc0 = 0.004235, Cx = -0.001667;
  const double nu = 16.0 / kPi * std::cbrt(3.0 * kPi * kPi), beta = nu * Cc0;
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T ec = F_pw<T, 0>(v) / rho;
  const T phi = phi_zeta(v);
  const T phi3 = phi * sq(phi);
  const T sigma = v.saa + 2.0 * v.sab + v.sbb;
  const T kF = cbrt(3.0 * kPi * kPi * rho);
  const T t2 = sigma * kPi / (16.0 * sq(phi) * kF * sq(rho));
  const T A = (2.0 * alpha / beta) / expm1(-2.0 * alpha * ec / (phi3 * beta * beta));
  const T H0 = (beta * beta / (2.0 * alpha)) * phi3 *
               log(1.0 + (2.0 * alpha / beta) * (t2 + A * sq(t2)) / (1.0 + A * t2 + sq(A) * sq(t2)));
  const T ks2_kF2 = 4.0 / (kPi * kF);
  const T H1 = nu * (rg_Cc(rs) - Cc0 - 3.0 * Cx / 7.0) * phi3 * t2 * exp(-100.0 * sq(sq(phi)) * ks2_kF2 * t2);
  return rho * (ec + H0 + H1);
}
// libxc hyb_gga_xc_b3pw91 (id 401): a0 = 0.20, ax = 0.72, ac = 0.81 over {lda_x, gga_x_b88, lda_c_pw, gga_c_pw91}
// (Becke, J. Chem. Phys. 98, 5648 (1993))
template <class T> T F_b3pw91(const Vars<T> &v) {
  return 0.08 * F_lda_x(v) + 0.72 * F_b88(v) + 0.19 * F_pw<T, 0>(v) + 0.81 * F_pw91_c(v);
}

// libxc hyb_gga_xc_b3p86 (id 403): a0 = 0.20, ax = 0.72, ac = 0.81 over {lda_x, gga_x_b88, lda_c_vwn_rpa, gga_c_p86}
template <class T> T F_b3p86(const Vars<T> &v) {
  return 0.08 * F_lda_x(v) + 0.72 * F_b88(v) + 0.19 * F_vwn<T, 1>(v) + 0.81 * F_p86(v);
}

template <class T> bool F_dispatch(int id, const Vars<T> &v, T &out) {
  switch (id) {
  case ORC_HYB_GGA_XC_WB97X_V: out = F_wb97x_v(v); return true;
  case ORC_HYB_GGA_XC_B3P86: out = F_b3p86(v); return true;
  case ORC_GGA_X_B86_MGC: out = F_b86b(v); return true;
  case ORC_HYB_GGA_XC_WB97X: out = F_wb97x(v); return true;
  case ORC_HYB_GGA_XC_B97: out = F_b97(v); return true;
  case ORC_HYB_GGA_XC_B3PW91: out = F_b3pw91(v); return true;
  case ORC_GGA_C_PW91: out = F_pw91_c(v); return true;
  case ORC_GGA_C_P86: out = F_p86(v); return true;
  case ORC_LDA_C_PZ: out = F_pz81(v); return true;
  case ORC_LDA_C_PW: out = F_pw<T, 0>(v); return true;
  case ORC_LDA_C_PW_MOD: out = F_pw<T, 1>(v); return true;
  case ORC_MGGA_X_TPSS: out = F_tpss_x(v); return true;
  case ORC_MGGA_C_TPSS: out = F_tpss_c(v); return true;
  case ORC_HYB_GGA_XC_B3LYP5: out = F_b3lyp5(v); return true;
  case ORC_HYB_GGA_XC_PBEH: out = F_pbeh(v); return true;
  case ORC_HYB_GGA_XC_BHANDH: out = F_bhandh(v); return true;
  case ORC_HYB_GGA_XC_BHANDHLYP: out = F_bhandhlyp(v); return true;
  case ORC_LDA_X: out = F_lda_x(v); return true;
  case ORC_LDA_C_VWN: out = F_vwn<T, 0>(v); return true;
  case ORC_LDA_C_VWN_RPA: out = F_vwn<T, 1>(v); return true;
  case ORC_LDA_C_VWN_3: out = F_vwn<T, 2>(v); return true;
  case ORC_GGA_X_B88: out = F_b88(v); return true;
  case ORC_GGA_C_LYP: out = F_lyp(v); return true;
  case ORC_GGA_X_PBE: out = F_pbe_x(v); return true;
  case ORC_GGA_C_PBE: out = F_pbe_c(v); return true;
  case ORC_HYB_GGA_XC_B3LYP: out = F_b3lyp(v); return true;
  case ORC_MGGA_X_R2SCAN: out = F_r2scan_x(v); return true;
  case ORC_MGGA_C_R2SCAN: out = F_r2scan_c(v); return true;
  default: return false;
  }
}

} // namespace

int xc_eval_add(int id, int polarized, size_t n, const double *rho, const double *sigma, const double *tau,
                double *exc, double *vrho, double *vsigma, double *vtau) {
  const int fam = orc_xc_family(id);
  if (fam < 0) return 1;
  for (size_t i = 0; i < n; ++i) {
    if (!polarized) {
      const double r = rho[i];
      if (!(r > kDensThreshold)) continue;
      using D = Dual<3>;
      const D R = D::var(r, 0);
      const double t0 = fam >= 2 ? std::max(tau[i], kTauFloor) : kTauFloor;
      double s0 = fam >= 1 ? std::max(sigma[i], kSigmaFloor) : kSigmaFloor;
      // libxc meta-GGAs (XC_FLAGS_ENFORCE_FHC): the Fermi-hole-curvature bound sigma <= 8 rho tau is applied to the
      // INPUT VALUE; the potentials stay the plain partial derivatives at the bounded point (no chain rule through
      // the bound) — this is what reproduces vrho/vsigma/vtau of tests/dft_tests.cpp:138-191, where z = 1 + 1 ulp
      if (fam >= 2) s0 = std::min(s0, 8.0 * r * t0);
      const D S = D::var(s0, 1);
      const D Tt = D::var(t0, 2);
      Vars<D> v{0.5 * R, 0.5 * R, 0.25 * S, 0.25 * S, 0.25 * S, 0.5 * Tt, 0.5 * Tt};
      D F;
      F_dispatch(id, v, F);
      exc[i] += F.v / r;
      vrho[i] += F.d[0];
      if (fam >= 1) vsigma[i] += F.d[1];
      if (fam >= 2) vtau[i] += F.d[2];
    } else {
      const double ra0 = rho[2 * i], rb0 = rho[2 * i + 1];
      if (!(ra0 + rb0 > kDensThreshold)) continue;
      using D = Dual<7>;
      const double ra = std::max(ra0, kSpinFloor), rb = std::max(rb0, kSpinFloor);
      double saa = kSigmaFloor, sab = 0.0, sbb = kSigmaFloor, ta = kTauFloor, tb = kTauFloor;
      if (fam >= 1) {
        saa = std::max(sigma[3 * i], kSigmaFloor);
        sbb = std::max(sigma[3 * i + 2], kSigmaFloor);
        sab = sigma[3 * i + 1];
        const double lim = 0.5 * (saa + sbb); // libxc: |sigma_ab| <= (sigma_aa + sigma_bb)/2
        sab = std::min(std::max(sab, -lim), lim);
      }
      if (fam >= 2) {
        ta = std::max(tau[2 * i], kTauFloor);
        tb = std::max(tau[2 * i + 1], kTauFloor);
        // per-spin-channel Fermi-hole-curvature bound, value only (see the unpolarized branch)
        saa = std::min(saa, 8.0 * ra * ta);
        sbb = std::min(sbb, 8.0 * rb * tb);
        const double lim2 = 0.5 * (saa + sbb);
        sab = std::min(std::max(sab, -lim2), lim2);
      }
      Vars<D> v{D::var(ra, 0), D::var(rb, 1), D::var(saa, 2), D::var(sab, 3), D::var(sbb, 4), D::var(ta, 5), D::var(tb, 6)};
      D F;
      F_dispatch(id, v, F);
      exc[i] += F.v / (ra + rb);
      vrho[2 * i] += F.d[0];
      vrho[2 * i + 1] += F.d[1];
      if (fam >= 1) {
        vsigma[3 * i] += F.d[2];
        vsigma[3 * i + 1] += F.d[3];
        vsigma[3 * i + 2] += F.d[4];
      }
      if (fam >= 2) {
        vtau[2 * i] += F.d[5];
        vtau[2 * i + 1] += F.d[6];
      }
    }
  }
  return 0;
}

} // namespace orc

extern "C" {

int orc_xc_family(int id) {
  switch (id) {
  case ORC_LDA_X:
  case ORC_LDA_C_VWN:
  case ORC_LDA_C_VWN_RPA:
  case ORC_LDA_C_VWN_3:
  case ORC_LDA_C_PW:
  case ORC_LDA_C_PW_MOD:
  case ORC_LDA_C_PZ:
    return 0;
  case ORC_HYB_GGA_XC_B97:
  case ORC_HYB_GGA_XC_B3PW91:
  case ORC_GGA_C_PW91:
  case ORC_GGA_C_P86:
  case ORC_GGA_X_B88:
  case ORC_GGA_C_LYP:
  case ORC_GGA_X_PBE:
  case ORC_GGA_C_PBE:
  case ORC_HYB_GGA_XC_B3LYP:
  case ORC_HYB_GGA_XC_B3LYP5:
  case ORC_HYB_GGA_XC_PBEH:
  case ORC_HYB_GGA_XC_BHANDH:
  case ORC_HYB_GGA_XC_BHANDHLYP:
  case ORC_HYB_GGA_XC_WB97X:
  case ORC_HYB_GGA_XC_WB97X_V:
  case ORC_HYB_GGA_XC_B3P86:
  case ORC_GGA_X_B86_MGC:
    return 1;
  case ORC_MGGA_X_R2SCAN:
  case ORC_MGGA_C_R2SCAN:
  case ORC_MGGA_X_TPSS:
  case ORC_MGGA_C_TPSS:
    return 2;
  default:
    return -1;
  }
}

double orc_xc_exx(int id) { // xc_hyb_exx_coef (include/occ/dft/functional.h:557-579)
  switch (id) {
  case ORC_HYB_GGA_XC_B3LYP:
  case ORC_HYB_GGA_XC_B3LYP5:
  case ORC_HYB_GGA_XC_B3PW91:
  case ORC_HYB_GGA_XC_B3P86:
    return 0.20;
  case ORC_HYB_GGA_XC_B97:
    return 0.1943;
  case ORC_HYB_GGA_XC_PBEH:
    return 0.25;
  case ORC_HYB_GGA_XC_BHANDH:
  case ORC_HYB_GGA_XC_BHANDHLYP:
    return 0.50;
  case ORC_HYB_GGA_XC_WB97X_V: // beta = -0.833 at omega = 0.3
  case ORC_HYB_GGA_XC_WB97X:
    return 1.0; // libxc: xc_hyb_exx_coef returns cam_alpha (the full-range fraction); beta = -0.842294 at omega = 0.3
  default:
    return 0.0;
  }
}

int orc_xc_eval(int id, int polarized, size_t n, const double *rho, const double *sigma, const double *tau,
                double *exc, double *vrho, double *vsigma, double *vtau) {
  const int fam = orc_xc_family(id);
  if (fam < 0) return 1;
  const int s = polarized ? 2 : 1;
  for (size_t i = 0; i < n; ++i) exc[i] = 0.0;
  for (size_t i = 0; i < n * s; ++i) vrho[i] = 0.0;
  if (fam >= 1 && vsigma)
    for (size_t i = 0; i < n * (polarized ? 3 : 1); ++i) vsigma[i] = 0.0;
  if (fam >= 2 && vtau)
    for (size_t i = 0; i < n * s; ++i) vtau[i] = 0.0;
  return orc::xc_eval_add(id, polarized, n, rho, sigma, tau, exc, vrho, vsigma, vtau);
}

} // extern "C"
