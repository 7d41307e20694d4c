// xc_functionals.cuh — the LDA / GGA / meta-GGA exchange-correlation functionals of the XC build, written
// for the device (and callable on the host for table checks).  They replace the libxc 6.2.2 calls the
// reference makes per grid block (src/dft/functional.cpp:18-65: xc_lda_exc_vxc / xc_gga_exc_vxc /
// xc_mgga_exc_vxc) for the functional ids reachable from src/dft/dft_method.cpp:14-37:
//   lda_x(1) lda_c_vwn(7) lda_c_vwn_rpa(8) lda_c_vwn_3(30) gga_x_pbe(101) gga_x_b88(106) gga_c_pbe(130)
//   gga_c_lyp(131) hyb_gga_xc_b3lyp(402) mgga_x_r2scan(497) mgga_c_r2scan(498) mgga_x_tpss(202) mgga_c_tpss(231)
//   hyb_gga_xc_b3lyp5(475) hyb_gga_xc_pbeh(406) hyb_gga_xc_bhandh(435) hyb_gga_xc_bhandhlyp(436) lda_c_pw(12)
//   hyb_gga_xc_wb97x_v(466) hyb_gga_xc_b3p86(403) gga_x_b86_mgc(105)
//   lda_c_pw_mod(13) hyb_gga_xc_wb97x(464) hyb_gga_xc_b97(407) hyb_gga_xc_b3pw91(401) gga_c_pw91(134) gga_c_p86(132)
//   lda_c_pz(9)
// libxc conventions: exc = energy per particle; vrho = d(rho*exc)/d rho_s; vsigma = d(rho*exc)/d sigma with
// sigma = grad rho . grad rho (polarized order aa, ab, bb); vtau = d(rho*exc)/d tau_s, tau = 1/2 sum |grad psi|^2.
// Each functional is a single energy density in (rho_a, rho_b, sigma_aa, sigma_ab, sigma_bb, tau_a, tau_b);
// the potentials come from the dual-number derivatives (xc_dual.cuh).  No libxc, no CPU fallback.
#pragma once
#include "xc_dual.cuh"

namespace xcb {

enum : int {
  ID_LDA_X = 1, ID_LDA_C_VWN = 7, ID_LDA_C_VWN_RPA = 8, ID_LDA_C_VWN_3 = 30, ID_GGA_X_PBE = 101,
  ID_GGA_X_B88 = 106, ID_GGA_C_PBE = 130, ID_GGA_C_LYP = 131, ID_HYB_GGA_XC_B3LYP = 402,
  ID_MGGA_X_R2SCAN = 497, ID_MGGA_C_R2SCAN = 498, ID_LDA_C_PW = 12, ID_LDA_C_PW_MOD = 13, ID_MGGA_X_TPSS = 202,
  ID_MGGA_C_TPSS = 231, ID_HYB_GGA_XC_PBEH = 406, ID_HYB_GGA_XC_BHANDH = 435, ID_HYB_GGA_XC_BHANDHLYP = 436,
  ID_HYB_GGA_XC_B3LYP5 = 475, ID_HYB_GGA_XC_WB97X = 464, ID_HYB_GGA_XC_B97 = 407, ID_HYB_GGA_XC_B3PW91 = 401,
  ID_GGA_C_PW91 = 134, ID_GGA_C_P86 = 132, ID_LDA_C_PZ = 9, ID_HYB_GGA_XC_WB97X_V = 466, ID_HYB_GGA_XC_B3P86 = 403,
  ID_GGA_X_B86_MGC = 105
};

XC_HD double xc_exx_of(int id) { // xc_hyb_exx_coef (include/occ/dft/functional.h:557-579)
  switch (id) {
  case ID_HYB_GGA_XC_B3LYP: case ID_HYB_GGA_XC_B3LYP5: case ID_HYB_GGA_XC_B3PW91: case ID_HYB_GGA_XC_B3P86:
    return 0.20;
  case ID_HYB_GGA_XC_B97: return 0.1943; // Becke, J. Chem. Phys. 107, 8554 (1997)
  case ID_HYB_GGA_XC_PBEH: return 0.25;
  case ID_HYB_GGA_XC_BHANDH: case ID_HYB_GGA_XC_BHANDHLYP: return 0.50;
  case ID_HYB_GGA_XC_WB97X_V:
  case ID_HYB_GGA_XC_WB97X: return 1.0; // cam_alpha: libxc's xc_hyb_exx_coef returns the full-range fraction
  default: return 0.0;
  }
}

// xc_hyb_cam_coef (include/occ/dft/functional.h:581-603): E_x^HF = alpha K + beta K^SR(omega); omega = 0: not range separated
XC_HD void xc_cam_of(int id, double *omega, double *alpha, double *beta) {
  *omega = 0.0;
  *alpha = xc_exx_of(id);
  *beta = 0.0;
  if (id == ID_HYB_GGA_XC_WB97X) {
    *omega = 0.3;
    *alpha = 1.0;
    *beta = -(1.0 - 0.157706);
  }
  if (id == ID_HYB_GGA_XC_WB97X_V) {
    *omega = 0.3;
    *alpha = 1.0;
    *beta = -(1.0 - 0.167);
  }
}

XC_HD int xc_family_of(int id) {
  switch (id) {
  case ID_LDA_X: case ID_LDA_C_VWN: case ID_LDA_C_VWN_RPA: case ID_LDA_C_VWN_3: case ID_LDA_C_PW:
  case ID_LDA_C_PW_MOD: case ID_LDA_C_PZ: return 0;
  case ID_HYB_GGA_XC_B97: case ID_HYB_GGA_XC_B3PW91: case ID_GGA_C_PW91: case ID_GGA_C_P86:
  case ID_GGA_X_PBE: case ID_GGA_X_B88: case ID_GGA_C_PBE: case ID_GGA_C_LYP: case ID_HYB_GGA_XC_B3LYP:
  case ID_HYB_GGA_XC_B3LYP5: case ID_HYB_GGA_XC_PBEH: case ID_HYB_GGA_XC_BHANDH: case ID_HYB_GGA_XC_BHANDHLYP:
  case ID_HYB_GGA_XC_WB97X: case ID_HYB_GGA_XC_WB97X_V: case ID_HYB_GGA_XC_B3P86: case ID_GGA_X_B86_MGC: return 1;
  case ID_MGGA_X_R2SCAN: case ID_MGGA_C_R2SCAN: case ID_MGGA_X_TPSS: case ID_MGGA_C_TPSS: return 2;
  default: return -1;
  }
}

// thresholds (libxc-style): nothing is evaluated below DENS_THR total density; spin densities, sigma and
// tau are floored so that every expression stays differentiable
#define XC_DENS_THR 1e-15
#define XC_SPIN_FLOOR 1e-15
#define XC_SIGMA_FLOOR 1e-40
#define XC_TAU_FLOOR 1e-20

#define XC_PI 3.14159265358979323846
#define XC_LDAX_C (-0.9305257363491)          /* -(3/4)(6/pi)^(1/3) */
#define XC_TWO43M2 0.5198420997897464         /* 2^(4/3) - 2 */
#define XC_FPP0 1.7099209341613653            /* f''(0) = 4/(9(2^(1/3)-1)) */
#define XC_B88_CX 0.9305257363491             /* (3/8)(3/pi)^(1/3) 4^(2/3) */
#define XC_CF 2.871234000188191               /* (3/10)(3 pi^2)^(2/3) */
#define XC_TWO113 12.699208415745595          /* 2^(11/3) */
#define XC_EXU_C (-0.7385587663820223)        /* -(3/4)(3/pi)^(1/3) */
#define XC_S2_C 38.283120002509214            /* 4 (3 pi^2)^(2/3) */
#define XC_GAMMA 0.0310906908696549           /* (1 - ln 2)/pi^2 */
#define XC_3PI2 29.608813203268074
#define XC_RS_C 0.238732414637843             /* 3/(4 pi) */

template <class T> struct XV {
  T ra, rb, saa, sab, sbb, ta, tb;
};

// ---------------------------------------------------------------------------------------------
template <class T> XC_HD T e_lda_x(const XV<T> &v) { // Dirac exchange, spin-resolved
  return XC_LDAX_C * (v.ra * cbrt(v.ra) + v.rb * cbrt(v.rb));
}

// ---------------------------------------------------------------------------------------------
// Vosko-Wilk-Nusair 1980. Parameter sets: 0 paramagnetic, 1 ferromagnetic, 2 spin stiffness;
// MC = fit to Ceperley-Alder (VWN5), RPA = fit to RPA data.
struct VwnP { double A, b, c, x0; };
XC_HD VwnP vwn_mc(int k) {
  switch (k) {
  case 0: return {0.0310907, 3.72744, 12.9352, -0.10498};
  case 1: return {0.01554535, 7.06042, 18.0578, -0.32500};
  default: return {-1.0 / (6.0 * XC_PI * XC_PI), 1.13107, 13.0045, -0.0047584};
  }
}
XC_HD VwnP vwn_rpa(int k) {
  switch (k) {
  case 0: return {0.0310907, 13.0720, 42.7198, -0.409286};
  case 1: return {0.01554535, 20.1231, 101.578, -0.743294};
  default: return {-1.0 / (6.0 * XC_PI * XC_PI), 1.06835, 11.4813, -0.228344};
  }
}
template <class T> XC_HD T vwn_q(const VwnP p, const T &rs, const T &x) {
  const double Q = ::sqrt(4.0 * p.c - p.b * p.b);
  const double X0 = p.x0 * p.x0 + p.b * p.x0 + p.c;
  const T X = rs + p.b * x + p.c;
  const T at = atan(Q / (2.0 * x + p.b));
  const T xm = x - p.x0;
  return p.A * (log(rs / X) + (2.0 * p.b / Q) * at -
                (p.b * p.x0 / X0) * (log(xm * xm / X) + (2.0 * (p.b + 2.0 * p.x0) / Q) * at));
}
template <class T> XC_HD T f_zeta(const T &z) {
  const T p = 1.0 + z, m = 1.0 - z;
  return (p * cbrt(p) + m * cbrt(m) - 2.0) / XC_TWO43M2;
}
template <class T> XC_HD void rs_zeta(const XV<T> &v, T &rho, T &rs, T &z) {
  rho = v.ra + v.rb;
  rs = cbrt(XC_RS_C / rho);
  z = (v.ra - v.rb) / rho;
}
// VARIANT 0: lda_c_vwn (VWN5), 1: lda_c_vwn_rpa, 2: lda_c_vwn_3
template <int VARIANT, class T> XC_HD T e_vwn(const XV<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T x = sqrt(rs);
  const T fz = f_zeta(z);
  if (VARIANT == 1) {
    const T eP = vwn_q(vwn_rpa(0), rs, x), eF = vwn_q(vwn_rpa(1), rs, x);
    return rho * (eP + (eF - eP) * fz);
  }
  const T z2 = z * z;
  const T z4 = z2 * z2;
  const T eP = vwn_q(vwn_mc(0), rs, x), eF = vwn_q(vwn_mc(1), rs, x);
  const T dmc = eF - eP;
  if (VARIANT == 0) {
    const T ac = vwn_q(vwn_mc(2), rs, x);
    return rho * (eP + ac * fz * (1.0 - z4) / XC_FPP0 + dmc * fz * z4);
  }
  const T ePr = vwn_q(vwn_rpa(0), rs, x), eFr = vwn_q(vwn_rpa(1), rs, x), acr = vwn_q(vwn_rpa(2), rs, x);
  return rho * (eP + (dmc / (eFr - ePr)) * acr * fz * (1.0 - z4) / XC_FPP0 + dmc * fz * z4);
}

// ---------------------------------------------------------------------------------------------
// Becke 1988 exchange, beta = 0.0042, gamma = 6
template <class T> XC_HD T b88_spin(const T &r, const T &s) {
  const double beta = 0.0042, gam = 6.0;
  const T r43 = r * cbrt(r);
  const T x = sqrt(s) / r43;
  return -(r43 * (XC_B88_CX + beta * (x * x) / (1.0 + gam * beta * x * asinh(x))));
}
template <class T> XC_HD T e_b88(const XV<T> &v) { return b88_spin(v.ra, v.saa) + b88_spin(v.rb, v.sbb); }
// Becke 1986 "B86b" exchange (J. Chem. Phys. 85, 7184; libxc gga_x_b86_mgc = 105): beta = 0.00375, gamma = 0.007, power 4/5
template <class T> XC_HD T b86b_spin(const T &r, const T &s) {
  const T r43 = r * cbrt(r);
  const T x2 = s / (r43 * r43);
  return -(r43 * (XC_B88_CX + 0.00375 * x2 / pow(1.0 + 0.007 * x2, 0.8)));
}
template <class T> XC_HD T e_b86b(const XV<T> &v) { return b86b_spin(v.ra, v.saa) + b86b_spin(v.rb, v.sbb); }

// ---------------------------------------------------------------------------------------------
// Lee-Yang-Parr correlation in the Miehlich-Savin-Stoll-Preuss gradient form
template <class T> XC_HD T e_lyp(const XV<T> &v) {
  const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
  const T rho = v.ra + v.rb;
  const T rm13 = 1.0 / cbrt(rho);
  const T den = 1.0 + d * rm13;
  const T omega = exp(-c * rm13) / den * pow(rho, -11.0 / 3.0);
  const T delta = c * rm13 + d * rm13 / den;
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T ra2 = v.ra * v.ra, rb2 = v.rb * v.rb;
  const T ra83 = ra2 * cbrt(ra2), rb83 = rb2 * cbrt(rb2);
  const T rho2 = rho * rho;
  const T t1 = -4.0 * a / den * v.ra * v.rb / rho;
  const T brack = (XC_TWO113 * XC_CF) * (ra83 + rb83) + (47.0 / 18.0 - 7.0 * delta / 18.0) * stot -
                  (2.5 - delta / 18.0) * (v.saa + v.sbb) -
                  (delta - 11.0) / 9.0 * (v.ra / rho * v.saa + v.rb / rho * v.sbb);
  const T t2 = -(a * b) * omega *
               (v.ra * v.rb * brack - (2.0 / 3.0) * rho2 * stot + ((2.0 / 3.0) * rho2 - ra2) * v.sbb +
                ((2.0 / 3.0) * rho2 - rb2) * v.saa);
  return t1 + t2;
}

// ---------------------------------------------------------------------------------------------
// PBE 1996
template <class T> XC_HD T ex_unif(const T &n) { return XC_EXU_C * (n * cbrt(n)); }
template <class T> XC_HD T red_grad2(const T &n, const T &sigma) { // s^2 = sigma / (4 (3 pi^2)^(2/3) n^(8/3))
  const T n2 = n * n;
  return sigma / (XC_S2_C * (n2 * cbrt(n2)));
}
template <class T> XC_HD T pbe_x_total(const T &n, const T &sigma) {
  const double kappa = 0.804, mu = 0.2195149727645171;
  const T s2 = red_grad2(n, sigma);
  return ex_unif(n) * (1.0 + kappa - kappa / (1.0 + mu * s2 / kappa));
}
template <class T> XC_HD T e_pbe_x(const XV<T> &v) { // E[ra, rb] = (E[2 ra] + E[2 rb]) / 2
  return 0.5 * (pbe_x_total(2.0 * v.ra, 4.0 * v.saa) + pbe_x_total(2.0 * v.rb, 4.0 * v.sbb));
}

// Perdew-Wang 1992 LSDA correlation, "modified" (more-digit) constants
struct PwP { double A, a1, b1, b2, b3, b4; };
XC_HD PwP pw_set(int k) {
  switch (k) {
  case 0: return {0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294};
  case 1: return {0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517};
  default: return {0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671};
  }
}
template <class T> XC_HD T pw_g(const PwP p, const T &rs, const T &srs) {
  const T q1 = 2.0 * p.A * (p.b1 * srs + p.b2 * rs + p.b3 * rs * srs + p.b4 * (rs * rs));
  return -2.0 * p.A * (1.0 + p.a1 * rs) * log(1.0 + 1.0 / q1);
}
template <class T> XC_HD T pw_dg(const PwP p, const T &rs, const T &srs) { // d pw_g / d rs
  const T q1 = 2.0 * p.A * (p.b1 * srs + p.b2 * rs + p.b3 * rs * srs + p.b4 * (rs * rs));
  const T dq1 = p.A * (p.b1 / srs + 2.0 * p.b2 + 3.0 * p.b3 * srs + 4.0 * p.b4 * rs);
  return -2.0 * p.A * p.a1 * log(1.0 + 1.0 / q1) + 2.0 * p.A * (1.0 + p.a1 * rs) * dq1 / (q1 * q1 + q1);
}
template <class T> XC_HD T eps_pw(const T &rs, const T &z) {
  const T srs = sqrt(rs);
  const T e0 = pw_g(pw_set(0), rs, srs), e1 = pw_g(pw_set(1), rs, srs), mac = pw_g(pw_set(2), rs, srs);
  const T fz = f_zeta(z);
  const T z2 = z * z;
  const T z4 = z2 * z2;
  return e0 - mac * fz * (1.0 - z4) / XC_FPP0 + (e1 - e0) * fz * z4;
}
template <class T> XC_HD T deps_pw_drs(const T &rs, const T &z) {
  const T srs = sqrt(rs);
  const T e0 = pw_dg(pw_set(0), rs, srs), e1 = pw_dg(pw_set(1), rs, srs), mac = pw_dg(pw_set(2), rs, srs);
  const T fz = f_zeta(z);
  const T z2 = z * z;
  const T z4 = z2 * z2;
  return e0 - mac * fz * (1.0 - z4) / XC_FPP0 + (e1 - e0) * fz * z4;
}
// 1 + zeta and 1 - zeta formed as 2 rho_s / rho rather than 1 +- (ra - rb) / rho: near full polarisation (a spin
// density at its 1e-15 floor) the latter carries an O(1e-16) absolute rounding error in a quantity of size 1e-14, and
// the functions below have derivatives like (1 - zeta)^(-1/3) there — two libm's then differ in the 3rd digit.
template <class T> XC_HD void zeta_pm(const XV<T> &v, T &zp, T &zm) {
  const T inv = 2.0 / (v.ra + v.rb);
  zp = v.ra * inv;
  zm = v.rb * inv;
}
template <class T> XC_HD T phi_zeta(const XV<T> &v) {
  T p, m;
  zeta_pm(v, p, m);
  return 0.5 * (cbrt(p * p) + cbrt(m * m));
}
template <class T> XC_HD T e_pbe_c(const XV<T> &v) {
  const double beta = 0.06672455060314922;
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T ec = eps_pw(rs, z);
  const T phi = phi_zeta(v);
  const T phi2 = phi * phi;
  const T phi3 = phi * phi2;
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T kF = cbrt(XC_3PI2 * rho);
  const T t2 = stot * XC_PI / (16.0 * phi2 * kF * (rho * rho));
  const T A = (beta / XC_GAMMA) / expm1(-ec / (XC_GAMMA * phi3));
  const T At2 = A * t2;
  const T H = XC_GAMMA * phi3 * log(1.0 + (beta / XC_GAMMA) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
  return rho * (ec + H);
}

// ---------------------------------------------------------------------------------------------
// r2SCAN (Furness, Kaplan, Ning, Perdew, Sun 2020)
#define XC_R2_ETA 0.001
#define XC_R2_DP24 0.016983563041 /* dp2^4, dp2 = 0.361 */
template <class T> XC_HD T r2_poly(const T &a, const double c0, const double c1, const double c2, const double c3,
                                   const double c4, const double c5, const double c6, const double c7) {
  T r = c7 * a + c6;
  r = r * a + c5;
  r = r * a + c4;
  r = r * a + c3;
  r = r * a + c2;
  r = r * a + c1;
  r = r * a + c0;
  return r;
}
template <class T> XC_HD T r2_fx(const T &alpha) {
  if (val(alpha) < 0.0) return exp(-0.667 * alpha / (1.0 - alpha));
  if (val(alpha) <= 2.5)
    return r2_poly(alpha, 1.0, -0.667, -0.4445555, -0.663086601049, 1.451297044490, -0.887998041597,
                   0.234528941479, -0.023185843322);
  return -1.24 * exp(0.8 / (1.0 - alpha));
}
template <class T> XC_HD T r2_fc(const T &alpha) {
  if (val(alpha) < 0.0) return exp(-0.64 * alpha / (1.0 - alpha));
  if (val(alpha) <= 2.5)
    return r2_poly(alpha, 1.0, -0.64, -0.4352, -1.535685604549, 3.061560252175, -1.915710236206, 0.516884468372,
                   -0.051848879792);
  return -0.7 * exp(1.5 / (1.0 - alpha));
}
template <class T> XC_HD T r2scan_x_total(const T &n, const T &sigma_in, const T &tau) {
  const double k0 = 0.174, k1 = 0.065, mu = 10.0 / 81.0, a1 = 4.9479;
  const double C2 = -0.9353000875519996 * k0; // sum_i i c_xi times k0
  const double Ceta = 20.0 / 27.0 + XC_R2_ETA * 5.0 / 3.0;
  const T &sigma = sigma_in; // sigma <= 8 n tau already holds (input-stage bound in xc_eval_*)
  const T p = red_grad2(n, sigma);
  const T tauW = sigma / (8.0 * n);
  const T n2 = n * n;
  const T tauU = XC_CF * (n * cbrt(n2));
  const T alpha = (tau - tauW) / (tauU + XC_R2_ETA * tauW);
  const T x = (Ceta * C2 * exp(-(p * p) / XC_R2_DP24) + mu) * p;
  const T h1 = 1.0 + k1 - k1 / (1.0 + x / k1);
  const double h0 = 1.0 + k0;
  const T fx = r2_fx(alpha);
  const T gx = 1.0 - exp(-a1 / sqrt(sqrt(p)));
  return ex_unif(n) * (h1 + fx * (h0 - h1)) * gx;
}
template <class T> XC_HD T e_r2scan_x(const XV<T> &v) {
  return 0.5 * (r2scan_x_total(2.0 * v.ra, 4.0 * v.saa, 2.0 * v.ta) +
                r2scan_x_total(2.0 * v.rb, 4.0 * v.sbb, 2.0 * v.tb));
}
template <class T> XC_HD T e_r2scan_c(const XV<T> &v) {
  const double b1c = 0.0285764, b2c = 0.0889, b3c = 0.125541, chi_inf = 0.128026;
  const double dfc2 = -0.7114023342889984; // sum_i i c_ci
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T tau = v.ta + v.tb;
  const T p = red_grad2(rho, stot);
  T zp, zm;
  zeta_pm(v, zp, zm);
  const T ds = 0.5 * (zp * cbrt(zp * zp) + zm * cbrt(zm * zm));
  const T dxz = 0.5 * (zp * cbrt(zp) + zm * cbrt(zm));
  const T phi = phi_zeta(v);
  const T phi2 = phi * phi;
  const T phi3 = phi * phi2;
  const T tauW = stot / (8.0 * rho);
  const T rho2 = rho * rho;
  const T tauU = XC_CF * (rho * cbrt(rho2)) * ds;
  const T alpha = (tau - tauW) / (tauU + XC_R2_ETA * tauW);
  const T fc = r2_fc(alpha);

  // alpha -> 0 limit
  const T srs = sqrt(rs);
  const T den0 = 1.0 + b2c * srs + b3c * rs;
  const T e_lda0 = -b1c / den0;
  const T de_lda0 = b1c * (0.5 * b2c / srs + b3c) / (den0 * den0);
  const T z2 = z * z;
  const T z4 = z2 * z2;
  const T z12 = z4 * z4 * z4;
  const T gc = (1.0 - 2.3631 * (dxz - 1.0)) * (1.0 - z12);
  const T w0 = expm1(-e_lda0 / b1c);
  const T ginf = 1.0 / sqrt(sqrt(1.0 + 4.0 * chi_inf * p));
  const T H0 = b1c * log(1.0 + w0 * (1.0 - ginf));
  const T eps0 = (e_lda0 + H0) * gc;

  // alpha = 1 limit
  const T e_lsda1 = eps_pw(rs, z);
  const T de_lsda1 = deps_pw_drs(rs, z);
  const T beta = 0.066725 * (1.0 + 0.1 * rs) / (1.0 + 0.1778 * rs);
  const T t2 = 1.5073033983379012 * p / (phi2 * rs); // (3 pi^2/16)^(2/3)
  const T w1 = expm1(-e_lsda1 / (XC_GAMMA * phi3));
  const T y = beta * t2 / (XC_GAMMA * w1);
  const T e_lsda0 = e_lda0 * gc;
  const T de_lsda0 = de_lda0 * gc;
  const T dy = dfc2 / (27.0 * XC_GAMMA * ds * phi3 * w1) *
               (20.0 * rs * (de_lsda0 - de_lsda1) - 45.0 * XC_R2_ETA * (e_lsda0 - e_lsda1)) * p *
               exp(-(p * p) / XC_R2_DP24);
  const T g1 = 1.0 / sqrt(sqrt(1.0 + 4.0 * (y - dy)));
  const T H1 = XC_GAMMA * phi3 * log(1.0 + w1 * (1.0 - g1));
  const T eps1 = e_lsda1 + H1;
  return rho * (eps1 + fc * (eps0 - eps1));
}

// ---------------------------------------------------------------------------------------------
// Perdew-Wang 1992 with the constants as printed in 1992 (libxc lda_c_pw); eps_pw above is lda_c_pw_mod
template <class T> XC_HD T e_pw92(const XV<T> &v, bool mod) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  if (mod) return rho * eps_pw(rs, z);
  const T srs = sqrt(rs);
  const PwP p0 = {0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294};
  const PwP p1 = {0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517};
  const PwP p2 = {0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671};
  const T e0 = pw_g(p0, rs, srs), e1 = pw_g(p1, rs, srs), mac = pw_g(p2, rs, srs);
  const T fz = f_zeta(z);
  const T z2 = z * z;
  const T z4 = z2 * z2;
  return rho * (e0 - mac * fz * (1.0 - z4) / 1.709921 + (e1 - e0) * fz * z4);
}

// ---------------------------------------------------------------------------------------------
// TPSS (Tao, Perdew, Staroverov, Scuseria 2003); known answers: tests/dft_tests.cpp:138-191 of the reference
template <class T> XC_HD T tpss_x_total(const T &n, const T &sigma, const T &tau) {
  const double kappa = 0.804, b = 0.40, c = 1.59096, e = 1.537, mu = 0.21951;
  const double sqe = 1.2397580409095959; // sqrt(e)
  const double m = 10.0 / 81.0;
  const T p = red_grad2(n, sigma);
  const T tauW = sigma / (8.0 * n);
  const T z = tauW / tau;
  const T n2 = n * n;
  const T alpha = (tau - tauW) / (XC_CF * (n * cbrt(n2)));
  const T am1 = alpha - 1.0;
  const T qb = 0.45 * am1 / sqrt(1.0 + b * alpha * am1) + (2.0 / 3.0) * p;
  const T z2 = z * z, p2 = p * p;
  const T opz2 = 1.0 + z2;
  const T num = (m + c * z2 / (opz2 * opz2)) * p + (146.0 / 2025.0) * (qb * qb) -
                (73.0 / 405.0) * qb * sqrt(0.18 * z2 + 0.5 * p2) + (m * m / kappa) * p2 + (2.0 * sqe * m * 0.36) * z2 +
                (e * mu) * (p * p2);
  const T den = 1.0 + sqe * p;
  const T x = num / (den * den);
  return ex_unif(n) * (1.0 + kappa - kappa / (1.0 + x / kappa));
}
template <class T> XC_HD T e_tpss_x(const XV<T> &v) {
  return 0.5 * (tpss_x_total(2.0 * v.ra, 4.0 * v.saa, 2.0 * v.ta) + tpss_x_total(2.0 * v.rb, 4.0 * v.sbb, 2.0 * v.tb));
}
template <class T> XC_HD T pbe_c_eps_one_spin(const T &n, const T &sigma) { // eps_c^PBE(n, 0, grad n, 0): zeta = 1
  const double beta = 0.06672455060314922;
  const double phi2 = 0.6299605249474366, phi3 = 0.5; // phi(1) = 2^(-1/3)
  const T rs = cbrt(XC_RS_C / n);
  const T ec = pw_g(pw_set(1), rs, sqrt(rs));
  const T kF = cbrt(XC_3PI2 * n);
  const T t2 = sigma * XC_PI / (16.0 * phi2 * kF * (n * n));
  const T A = (beta / XC_GAMMA) / expm1(-ec / (XC_GAMMA * phi3));
  const T At2 = A * t2;
  return ec + XC_GAMMA * phi3 * log(1.0 + (beta / XC_GAMMA) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
}
template <class T> XC_HD T e_tpss_c(const XV<T> &v) {
  const T rho = v.ra + v.rb;
  const T zeta = (v.ra - v.rb) / rho;
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T z = stot / (8.0 * rho * (v.ta + v.tb));
  const T zt2 = zeta * zeta;
  const T zt4 = zt2 * zt2;
  const T C0 = 0.53 + 0.87 * zt2 + 0.50 * zt4 + 2.26 * (zt2 * zt4);
  const T rho2 = rho * rho;
  // xi^2 = |grad zeta|^2 / (4 (3 pi^2 n)^(2/3)), |grad zeta|^2 = 4 |nb grad na - na grad nb|^2 / n^4
  const T kF = cbrt(XC_3PI2 * rho);
  const T xi2 = (v.rb * v.rb * v.saa - 2.0 * (v.ra * v.rb) * v.sab + v.ra * v.ra * v.sbb) / (rho2 * rho2 * (kF * kF));
  T zp, zm;
  zeta_pm(v, zp, zm);
  const T w = 1.0 + 0.5 * xi2 * (1.0 / (zp * cbrt(zp)) + 1.0 / (zm * cbrt(zm)));
  const T w2 = w * w;
  const T C = C0 / (w2 * w2);
  const T e_pbe = e_pbe_c(v) / rho;
  T ea = pbe_c_eps_one_spin(v.ra, v.saa), eb = pbe_c_eps_one_spin(v.rb, v.sbb);
  if (val(ea) < val(e_pbe)) ea = e_pbe;
  if (val(eb) < val(e_pbe)) eb = e_pbe;
  const T z2 = z * z;
  const T pkzb = e_pbe * (1.0 + C * z2) - (1.0 + C) * z2 * ((v.ra * ea + v.rb * eb) / rho);
  return rho * pkzb * (1.0 + 2.8 * pkzb * (z2 * z));
}

// ---------------------------------------------------------------------------------------------
// wB97X (Chai & Head-Gordon, J. Chem. Phys. 128, 084106 (2008); libxc hyb_gga_xc_wb97x = 464): B97 power series in
// u = gamma x^2 / (1 + gamma x^2), x_s = |grad rho_s| / rho_s^(4/3), on (i) the LSDA exchange attenuated by the
// short-range operator erfc(omega r)/r, omega = 0.3, (ii) the same-spin and (iii) the opposite-spin parts of the
// Stoll-partitioned PW92 correlation (1992 constants).  Golden: tests/isolated_xc_tests.cpp:296-319 of the reference.
template <class T> XC_HD T wb97_attenuation(const T &a) { // eq. (8) of the paper; its 1/a^2 series for a >= 1, where
  const T a2 = a * a;                                      // the closed form cancels from O(a^4) terms to 1/(36 a^2)
  if (val(a) >= 1.0) {
    const T y = 1.0 / a2;
    T r = y * (1.0 / 99407521382400.0) - 1.0 / 2021444812800.0;
    r = r * y + 1.0 / 44590694400.0;
    r = r * y - 1.0 / 1073479680.0;
    r = r * y + 1.0 / 28385280.0;
    r = r * y - 1.0 / 829440.0;
    r = r * y + 1.0 / 26880.0;
    r = r * y - 1.0 / 960.0;
    r = r * y + 1.0 / 36.0;
    return r * y;
  }
  const T e = exp(-0.25 / a2);
  return 1.0 - (8.0 / 3.0) * a *
                   (1.7724538509055159 * erf(0.5 / a) - 3.0 * a + 4.0 * a * a2 + (2.0 * a - 4.0 * a * a2) * e);
}
template <class T> XC_HD T b97_poly(const T &x2, const double gam, const double c0, const double c1, const double c2,
                                    const double c3, const double c4) {
  const T gx = gam * x2;
  const T u = gx / (1.0 + gx);
  T r = c4 * u + c3;
  r = r * u + c2;
  r = r * u + c1;
  return r * u + c0;
}
template <class T> XC_HD T pw92_ferro_1992(const T &n) { // eps_c of the fully polarized gas of density n
  const PwP p1 = {0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517};
  const T rs = cbrt(XC_RS_C / n);
  return pw_g(p1, rs, sqrt(rs));
}
template <class T> XC_HD T e_wb97x(const XV<T> &v) {
  const double omega = 0.3;
  const T ra43 = v.ra * cbrt(v.ra), rb43 = v.rb * cbrt(v.rb);
  const T xa2 = v.saa / (ra43 * ra43), xb2 = v.sbb / (rb43 * rb43);
  const T kfa = cbrt(2.0 * XC_3PI2 * v.ra), kfb = cbrt(2.0 * XC_3PI2 * v.rb);
  const T ex = XC_LDAX_C * (ra43 * wb97_attenuation(0.5 * omega / kfa) *
                                b97_poly(xa2, 0.004, 0.842294, 0.726479, 1.04760, -5.70635, 13.2794) +
                            rb43 * wb97_attenuation(0.5 * omega / kfb) *
                                b97_poly(xb2, 0.004, 0.842294, 0.726479, 1.04760, -5.70635, 13.2794));
  const T eaa = v.ra * pw92_ferro_1992(v.ra), ebb = v.rb * pw92_ferro_1992(v.rb);
  const T eab = e_pw92(v, false) - eaa - ebb;
  return ex + eaa * b97_poly(xa2, 0.2, 1.0, -4.33879, 18.2308, -31.7430, 17.2901) +
         ebb * b97_poly(xb2, 0.2, 1.0, -4.33879, 18.2308, -31.7430, 17.2901) +
         eab * b97_poly(0.5 * (xa2 + xb2), 0.006, 1.0, 2.37031, -11.3995, 6.58405, -3.78132);
}

// wB97X-V (Mardirossian & Head-Gordon, PCCP 16, 9904 (2014); libxc hyb_gga_xc_wb97x_v = 466): the wB97X form with seven
// linear coefficients, omega = 0.3; the VV10 part (b = 6.0, C = 0.01) is occ_xc_vv10.  Golden: isolated_xc_tests.cpp:340-364.
template <class T> XC_HD T e_wb97x_v(const XV<T> &v) {
  const double omega = 0.3;
  const T ra43 = v.ra * cbrt(v.ra), rb43 = v.rb * cbrt(v.rb);
  const T xa2 = v.saa / (ra43 * ra43), xb2 = v.sbb / (rb43 * rb43);
  const T kfa = cbrt(2.0 * XC_3PI2 * v.ra), kfb = cbrt(2.0 * XC_3PI2 * v.rb);
  const T ex = XC_LDAX_C * (ra43 * wb97_attenuation(0.5 * omega / kfa) * b97_poly(xa2, 0.004, 0.833, 0.603, 1.194, 0.0, 0.0) +
                            rb43 * wb97_attenuation(0.5 * omega / kfb) * b97_poly(xb2, 0.004, 0.833, 0.603, 1.194, 0.0, 0.0));
  const T eaa = v.ra * pw92_ferro_1992(v.ra), ebb = v.rb * pw92_ferro_1992(v.rb);
  const T eab = e_pw92(v, false) - eaa - ebb;
  return ex + eaa * b97_poly(xa2, 0.2, 0.556, -0.257, 0.0, 0.0, 0.0) + ebb * b97_poly(xb2, 0.2, 0.556, -0.257, 0.0, 0.0, 0.0) +
         eab * b97_poly(0.5 * (xa2 + xb2), 0.006, 1.219, -1.850, 0.0, 0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// B97 (Becke, J. Chem. Phys. 107, 8554 (1997); libxc hyb_gga_xc_b97 = 407, 19.43 % exact exchange): the same power series
// as wB97X on the full-range LSDA exchange and the Stoll-partitioned PW92 correlation, three terms each.
template <class T> XC_HD T e_b97(const XV<T> &v) {
  const T ra43 = v.ra * cbrt(v.ra), rb43 = v.rb * cbrt(v.rb);
  const T xa2 = v.saa / (ra43 * ra43), xb2 = v.sbb / (rb43 * rb43);
  const T ex = XC_LDAX_C * (ra43 * b97_poly(xa2, 0.004, 0.8094, 0.5073, 0.7481, 0.0, 0.0) +
                            rb43 * b97_poly(xb2, 0.004, 0.8094, 0.5073, 0.7481, 0.0, 0.0));
  const T eaa = v.ra * pw92_ferro_1992(v.ra), ebb = v.rb * pw92_ferro_1992(v.rb);
  const T eab = e_pw92(v, false) - eaa - ebb;
  return ex + eaa * b97_poly(xa2, 0.2, 0.1737, 2.3487, -2.4868, 0.0, 0.0) +
         ebb * b97_poly(xb2, 0.2, 0.1737, 2.3487, -2.4868, 0.0, 0.0) +
         eab * b97_poly(0.5 * (xa2 + xb2), 0.006, 0.9454, 0.7471, -4.5961, 0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// Perdew-Zunger 1981 parametrisation of the Ceperley-Alder correlation energy (libxc lda_c_pz = 9), the local part of P86
template <class T> XC_HD T pz81_eps(const T &rs, const double g, const double b1, const double b2, const double A,
                                    const double B, const double C, const double D) {
  if (val(rs) >= 1.0) return g / (1.0 + b1 * sqrt(rs) + b2 * rs);
  const T l = log(rs);
  return A * l + B + C * rs * l + D * rs;
}
template <class T> XC_HD T e_pz81(const XV<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T eu = pz81_eps(rs, -0.1423, 1.0529, 0.3334, 0.0311, -0.048, 0.0020, -0.0116);
  const T ep = pz81_eps(rs, -0.0843, 1.3981, 0.2611, 0.01555, -0.0269, 0.0007, -0.0048);
  return rho * (eu + f_zeta(z) * (ep - eu));
}
// Rasolt-Geldart gradient coefficient shared by P86 and PW91: C(rs) = 0.001667 + (0.002568 + a rs + b rs^2) /
// (1 + c rs + d rs^2 + 1e4 b rs^3)
template <class T> XC_HD T rasolt_geldart_C(const T &rs) {
  const T rs2 = rs * rs;
  return 0.001667 + (0.002568 + 0.023266 * rs + 7.389e-6 * rs2) / (1.0 + 8.723 * rs + 0.472 * rs2 + 0.07389 * (rs2 * rs));
}
// Perdew 1986 gradient correction (Phys. Rev. B 33, 8822; libxc gga_c_p86 = 132):
//   E = E^PZ81 + int exp(-Phi) C(n) |grad n|^2 / (d n^(4/3)),  Phi = 1.745 f~ [C(inf) / C(n)] |grad n| / n^(7/6), f~ = 0.11,
//   d = 2^(1/3) sqrt(((1 + zeta)/2)^(5/3) + ((1 - zeta)/2)^(5/3))
template <class T> XC_HD T e_p86(const XV<T> &v) {
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T Cn = rasolt_geldart_C(rs);
  const T r16 = sqrt(cbrt(rho)); // n^(1/6)
  const T n76 = rho * r16;
  const T Phi = (1.745 * 0.11 * 0.004235) / Cn * sqrt(stot) / n76;
  const T zp = 2.0 * v.ra / rho, zm = 2.0 * v.rb / rho; // 1 +- zeta
  const T hp = 0.5 * zp, hm = 0.5 * zm;
  const T d = 1.2599210498948732 * sqrt(hp * cbrt(hp * hp) + hm * cbrt(hm * hm));
  return e_pz81(v) + exp(-Phi) * Cn * stot / (d * rho * cbrt(rho));
}
// Perdew-Wang 1991 correlation (libxc gga_c_pw91 = 134): eps = eps^PW92 + H0 + H1,
//   H0 = (beta^2 / 2 alpha) phi^3 ln[1 + (2 alpha / beta) (t^2 + A t^4) / (1 + A t^2 + A^2 t^4)],
//   A = (2 alpha / beta) / (exp(-2 alpha eps / (phi^3 beta^2)) - 1),  alpha = 0.09, beta = nu Cc(0), nu = (16/pi)(3 pi^2)^(1/3),
//   H1 = nu [Cc(rs) - Cc(0) - 3 Cx / 7] phi^3 t^2 exp(-100 phi^4 (ks^2 / kF^2) t^2),  Cc(0) = 0.004235, Cx = -0.001667
template <class T> XC_HD T e_pw91_c(const XV<T> &v) {
  const double alpha = 0.09, Cc0 = 0.004235, Cx = -0.001667;
  const double nu = 16.0 / XC_PI * 3.0936677262801355; // (3 pi^2)^(1/3)
  const double beta = nu * Cc0;
  T rho, rs, z;
  rs_zeta(v, rho, rs, z);
  const T ec = e_pw92(v, false) / rho;
  const T phi = phi_zeta(v);
  const T phi2 = phi * phi;
  const T phi3 = phi * phi2;
  const T stot = v.saa + 2.0 * v.sab + v.sbb;
  const T kF = cbrt(XC_3PI2 * rho);
  const T t2 = stot * XC_PI / (16.0 * phi2 * kF * (rho * rho)); // t = |grad n| / (2 phi ks n), ks^2 = 4 kF / pi
  const T A = (2.0 * alpha / beta) / expm1(-2.0 * alpha * ec / (phi3 * beta * beta));
  const T At2 = A * t2;
  const T H0 = (beta * beta / (2.0 * alpha)) * phi3 * log(1.0 + (2.0 * alpha / beta) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
  const T H1 = nu * (rasolt_geldart_C(rs) - Cc0 - 3.0 * Cx / 7.0) * phi3 * t2 *
               exp(-100.0 * (phi2 * phi2) * (4.0 / (XC_PI * kF)) * t2);
  return rho * (ec + H0 + H1);
}

// ---------------------------------------------------------------------------------------------
template <class T> XC_HD T e_b3lyp(const XV<T> &v) { // libxc hyb_gga_xc_b3lyp: a0 = 0.20, ax = 0.72, ac = 0.81
  return 0.08 * e_lda_x(v) + 0.72 * e_b88(v) + 0.19 * e_vwn<1>(v) + 0.81 * e_lyp(v);
}

// The functionals added in round 2 are reached through real calls: inlining them as well into every (family, spin) kernel
// took the build of this unit from under a minute to five (cicc + ptxas on one giant function per dual-number width).
#define XC_HD_CALL __host__ __device__ __noinline__
template <class T> XC_HD_CALL T call_b97(const XV<T> &v) { return e_b97(v); }
template <class T> XC_HD_CALL T call_pz81(const XV<T> &v) { return e_pz81(v); }
template <class T> XC_HD_CALL T call_p86(const XV<T> &v) { return e_p86(v); }
template <class T> XC_HD_CALL T call_pw91_c(const XV<T> &v) { return e_pw91_c(v); }
// libxc hyb_gga_xc_b3pw91: a0 = 0.20, ax = 0.72, ac = 0.81 over {lda_x, gga_x_b88, lda_c_pw, gga_c_pw91}
template <class T> XC_HD_CALL T call_b3pw91(const XV<T> &v) {
  return 0.08 * e_lda_x(v) + 0.72 * e_b88(v) + 0.19 * e_pw92(v, false) + 0.81 * e_pw91_c(v);
}

template <class T> XC_HD_CALL T call_wb97x_v(const XV<T> &v) { return e_wb97x_v(v); }
template <class T> XC_HD_CALL T call_b86b(const XV<T> &v) { return e_b86b(v); }
// libxc hyb_gga_xc_b3p86: a0 = 0.20, ax = 0.72, ac = 0.81 over {lda_x, gga_x_b88, lda_c_vwn_rpa, gga_c_p86}
template <class T> XC_HD_CALL T call_b3p86(const XV<T> &v) {
  return 0.08 * e_lda_x(v) + 0.72 * e_b88(v) + 0.19 * e_vwn<1>(v) + 0.81 * e_p86(v);
}

template <class T> XC_HD T xc_energy_density(int id, const XV<T> &v) {
  switch (id) {
  case ID_HYB_GGA_XC_WB97X_V: return call_wb97x_v(v);
  case ID_HYB_GGA_XC_B3P86: return call_b3p86(v);
  case ID_GGA_X_B86_MGC: return call_b86b(v);
  case ID_LDA_X: return e_lda_x(v);
  case ID_LDA_C_VWN: return e_vwn<0>(v);
  case ID_LDA_C_VWN_RPA: return e_vwn<1>(v);
  case ID_LDA_C_VWN_3: return e_vwn<2>(v);
  case ID_GGA_X_B88: return e_b88(v);
  case ID_GGA_C_LYP: return e_lyp(v);
  case ID_GGA_X_PBE: return e_pbe_x(v);
  case ID_GGA_C_PBE: return e_pbe_c(v);
  case ID_HYB_GGA_XC_B3LYP: return e_b3lyp(v);
  case ID_MGGA_X_R2SCAN: return e_r2scan_x(v);
  case ID_MGGA_C_R2SCAN: return e_r2scan_c(v);
  case ID_MGGA_X_TPSS: return e_tpss_x(v);
  case ID_MGGA_C_TPSS: return e_tpss_c(v);
  case ID_LDA_C_PW: return e_pw92(v, false);
  case ID_LDA_C_PW_MOD: return e_pw92(v, true);
  case ID_HYB_GGA_XC_B3LYP5: return 0.08 * e_lda_x(v) + 0.72 * e_b88(v) + 0.19 * e_vwn<0>(v) + 0.81 * e_lyp(v);
  case ID_HYB_GGA_XC_PBEH: return 0.75 * e_pbe_x(v) + e_pbe_c(v);
  case ID_HYB_GGA_XC_BHANDH: return 0.5 * e_lda_x(v) + e_lyp(v);
  case ID_HYB_GGA_XC_BHANDHLYP: return 0.5 * e_b88(v) + e_lyp(v);
  case ID_HYB_GGA_XC_WB97X: return e_wb97x(v);
  case ID_HYB_GGA_XC_B97: return call_b97(v);
  case ID_LDA_C_PZ: return call_pz81(v);
  case ID_GGA_C_P86: return call_p86(v);
  case ID_GGA_C_PW91: return call_pw91_c(v);
  case ID_HYB_GGA_XC_B3PW91: return call_b3pw91(v);
  default: return T(0.0);
  }
}

struct XCPoint { // one grid point, summed over the functional components (functional.h:438-459 Result::operator+=)
  double exc;
  double vrho[2], vsigma[3], vtau[2];
};

// Unpolarized point. FAM: 0 LDA, 1 GGA, 2 meta-GGA = family of the BUILD (max over components); components of a
// lower family simply ignore the extra inputs, exactly like libxc called through Params of the build's family.
template <int FAM> XC_HD void xc_eval_unpolarized(int nfunc, const int *ids, double rho, double sigma, double tau,
                                                  XCPoint &o) {
  o.exc = 0.0;
  o.vrho[0] = o.vrho[1] = o.vsigma[0] = o.vsigma[1] = o.vsigma[2] = o.vtau[0] = o.vtau[1] = 0.0;
  if (!(rho > XC_DENS_THR)) return;
  constexpr int N = FAM + 1;
  typedef Dual<N> D;
  const D R = D::seed(rho, 0);
  const D S = FAM >= 1 ? D::seed(fmax(sigma, XC_SIGMA_FLOOR), FAM >= 1 ? 1 : 0) : D(XC_SIGMA_FLOOR);
  const D Tt = FAM >= 2 ? D::seed(fmax(tau, XC_TAU_FLOOR), FAM >= 2 ? 2 : 0) : D(XC_TAU_FLOOR);
  XV<D> v;
  v.ra = 0.5 * R;
  v.rb = v.ra;
  v.saa = 0.25 * S;
  v.sab = v.saa;
  v.sbb = v.saa;
  v.ta = 0.5 * Tt;
  v.tb = v.ta;
  // meta-GGA components see sigma bounded by 8 rho tau (libxc XC_FLAGS_ENFORCE_FHC): the bound changes the input
  // VALUE only, the potentials remain the plain partials there (pinned by tests/dft_tests.cpp:138-191, z = 1 + 1 ulp)
  XV<D> vm = v;
  if (FAM >= 2) {
    D Sm = S;
    Sm.v = fmin(S.v, 8.0 * rho * Tt.v);
    vm.saa = 0.25 * Sm;
    vm.sab = vm.saa;
    vm.sbb = vm.saa;
  }
  for (int f = 0; f < nfunc; ++f) {
    const D F = xc_energy_density(ids[f], (FAM >= 2 && xc_family_of(ids[f]) == 2) ? vm : v);
    o.exc += F.v / rho;
    o.vrho[0] += F.d[0];
    if (FAM >= 1) o.vsigma[0] += F.d[FAM >= 1 ? 1 : 0];
    if (FAM >= 2) o.vtau[0] += F.d[FAM >= 2 ? 2 : 0];
  }
}

template <int FAM> XC_HD void xc_eval_polarized(int nfunc, const int *ids, double ra0, double rb0, double saa0,
                                                double sab0, double sbb0, double ta0, double tb0, XCPoint &o) {
  o.exc = 0.0;
  o.vrho[0] = o.vrho[1] = o.vsigma[0] = o.vsigma[1] = o.vsigma[2] = o.vtau[0] = o.vtau[1] = 0.0;
  if (!(ra0 + rb0 > XC_DENS_THR)) return;
  constexpr int N = FAM == 0 ? 2 : (FAM == 1 ? 5 : 7);
  typedef Dual<N> D;
  const double ra = fmax(ra0, XC_SPIN_FLOOR), rb = fmax(rb0, XC_SPIN_FLOOR);
  XV<D> v;
  v.ra = D::seed(ra, 0);
  v.rb = D::seed(rb, 1);
  if (FAM >= 1) {
    const double saa = fmax(saa0, XC_SIGMA_FLOOR), sbb = fmax(sbb0, XC_SIGMA_FLOOR);
    const double lim = 0.5 * (saa + sbb); // |sigma_ab| <= (sigma_aa + sigma_bb)/2
    const double sab = fmin(fmax(sab0, -lim), lim);
    v.saa = D::seed(saa, N > 2 ? 2 : 0);
    v.sab = D::seed(sab, N > 3 ? 3 : 0);
    v.sbb = D::seed(sbb, N > 4 ? 4 : 0);
  } else {
    v.saa = D(XC_SIGMA_FLOOR);
    v.sab = D(0.0);
    v.sbb = D(XC_SIGMA_FLOOR);
  }
  if (FAM >= 2) {
    v.ta = D::seed(fmax(ta0, XC_TAU_FLOOR), N > 5 ? 5 : 0);
    v.tb = D::seed(fmax(tb0, XC_TAU_FLOOR), N > 6 ? 6 : 0);
  } else {
    v.ta = D(XC_TAU_FLOOR);
    v.tb = D(XC_TAU_FLOOR);
  }
  XV<D> vm = v;
  if (FAM >= 2) { // per-spin-channel Fermi-hole-curvature bound, value only (see xc_eval_unpolarized)
    vm.saa.v = fmin(v.saa.v, 8.0 * ra * v.ta.v);
    vm.sbb.v = fmin(v.sbb.v, 8.0 * rb * v.tb.v);
    const double lim2 = 0.5 * (vm.saa.v + vm.sbb.v);
    vm.sab.v = fmin(fmax(v.sab.v, -lim2), lim2);
  }
  for (int f = 0; f < nfunc; ++f) {
    const D F = xc_energy_density(ids[f], (FAM >= 2 && xc_family_of(ids[f]) == 2) ? vm : v);
    o.exc += F.v / (ra + rb);
    o.vrho[0] += F.d[0];
    o.vrho[1] += F.d[1];
    if (FAM >= 1) {
      o.vsigma[0] += F.d[N > 2 ? 2 : 0];
      o.vsigma[1] += F.d[N > 3 ? 3 : 0];
      o.vsigma[2] += F.d[N > 4 ? 4 : 0];
    }
    if (FAM >= 2) {
      o.vtau[0] += F.d[N > 5 ? 5 : 0];
      o.vtau[1] += F.d[N > 6 ? 6 : 0];
    }
  }
}

} // namespace xcb
