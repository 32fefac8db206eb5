// bgc_co2.cuh — carbonate chemistry of the reference's module co2calc as
// register-resident device functions for sm_100a.
//
//   co3_coeffs        <- comp_co3_coeffs     co2calc.F90:320-777
//   talk_residual     <- talk_row            co2calc.F90:1001-1092
//   solve_htotal      <- comp_htotal + drtsafe_row   co2calc.F90:781-997
//   co3_sat_vals      <- comp_co3_sat_vals   co2calc.F90:1096-1238
//
// The reference passes the equilibrium constants between its routines through
// module-level SAVE scalars (co2calc.F90:65-67), which makes it non-reentrant;
// here they are a per-thread struct that lives in registers.
//
// The root finder runs warp-synchronously: every lane of a warp stays in the
// iteration until __all_sync says the whole warp has converged; converged lanes
// are frozen by predication.  Control flow is therefore warp-uniform and a
// slow lane costs its own warp a few extra passes, never the whole block.
#pragma once
#include <cuda_runtime.h>
#include "bgc_math.cuh"

namespace bgc {

constexpr unsigned FULL_MASK = 0xffffffffu;

// co2calc.F90:30-59
constexpr double kRhoSw = 1.026;                 // g/cm^3
constexpr double kT0Kelvin = 273.15;             // co2calc's own constant, not T0_Kelvin_BGC
constexpr double kXacc = 1e-10;
constexpr int kMaxIt = 100;
constexpr double kSaltMin = 0.1;
constexpr double kDicMin = kSaltMin / 35.0 * 1944.0;
constexpr double kAlkMin = kSaltMin / 35.0 * 2225.0;
constexpr double kMassToVol = 1e6 * kRhoSw;
constexpr double kVolToMass = 1.0 / kMassToVol;
// LOG(c10) and LOG(1e-2) are folded at compile time by the reference compiler;
// pin the correctly rounded doubles instead of calling log() on the device.
constexpr double kLn10 = 2.302585092994046;
constexpr double kLn1em2 = -4.605170185988091;

// -log10(h): pH from the hydrogen-ion concentration (h is a normal, positive double)
__device__ __forceinline__ double ph_of_h(double h) {
#ifdef BGC_STRICT
  return -log10(h);
#else
  return -blog(h) * (1.0 / kLn10);
#endif
}
// The reference's bracket-growth loop has no exit (co2calc.F90:920-938; the
// abort at :931-933 is commented out).  The ratio x2/x1 squares on every pass,
// so 64 passes overflow any finite bracket: cap there and raise a status flag.
constexpr int kBracketGrowCap = 64;

struct Co3Consts {
  double k1, k2;   // NOT pressure corrected (captured before the correction, co2calc.F90:478,507)
  double ff;       // only meaningful when WANT_FF
  double kw, kb, ks, kf, k1p, k2p, k3p, ksi;
  double bt, st, ft;
};

struct Co3Totals { double dic, ta, pt, sit; };   // mol/kg, after the floors of co2calc.F90:843-846

// POP ref_pressure fit, co2calc.F90:371-372 (depth in m -> bar)
template <class EXP>
__device__ __forceinline__ double press_bar_of_depth(double depth, const EXP &ex) {
  return 0.059808 * (ex(-0.025 * depth) - 1.0) + 0.100766 * depth + 2.28405e-7 * (depth * depth);
}

// Pressure factor bexp((-deltaV + 0.5*Kappa*P)*P/(R*T)), Millero 1995.
template <class EXP>
__device__ __forceinline__ double kfac(double deltaV, double Kappa, double press_bar, double invRtk, const EXP &ex) {
  return ex((-deltaV + 0.5 * Kappa * press_bar) * press_bar * invRtk);
}

// `deep` is the reference's (k > 1): the pressure correction is keyed on the
// LEVEL INDEX, not on depth (co2calc.F90:480 ...), so level 1 is never corrected.
// K0 is never consumed by any caller and the Kfac factors of k1/k2 never leave
// comp_co3_coeffs (sk1/sk2 are captured first) -> neither is computed.
template <bool WANT_FF, class EXP>
__device__ __forceinline__ void co3_coeffs(bool deep, double depth, double temp, double salt,
                                           Co3Consts &c, const EXP &ex) {
  const double press_bar = press_bar_of_depth(depth, ex);

  const double salt_lim = gmax(salt, kSaltMin);
  const double tk = kT0Kelvin + temp;
  const double tk100 = tk * 1e-2;
  const double tk1002 = tk100 * tk100;
  const double invtk = frcp(tk);
  const double dlogtk = blog(tk);
  const double invRtk = (1.0 / 83.1451) * invtk;

  const double is = fdiv(19.924 * salt_lim, (1000.0 - 1.005 * salt_lim));
  const double is2 = is * is;
  const double sqrtis = sqrt(is);
  const double sqrts = sqrt(salt_lim);
  const double s2 = salt_lim * salt_lim;
  const double scl = cdiv(salt_lim, 1.80655, 1.0 / 1.80655);

  const double log_1_m_1p005em3_s = blog(1.0 - 0.001005 * salt_lim);
  double arg;

  if (WANT_FF) {   // Weiss & Price 1980, co2calc.F90:423-431
    arg = -162.8301 + fdiv(218.2968, tk100) + 90.9241 * (dlogtk + kLn1em2) - 1.47696 * tk1002 +
          salt_lim * (.025695 - .025225 * tk100 + 0.0049867 * tk1002);
    c.ff = ex(arg);
  } else {
    c.ff = 0.0;
  }

  // k1, k2: Lueker et al. 2000, total pH scale (k1_k2_pH_tot = .true. at every
  // call site on this path: co2calc.F90:285, BGC_mod.F90:2764)
  arg = 3633.86 * invtk - 61.2172 + 9.67770 * dlogtk - 0.011555 * salt_lim + 0.0001152 * s2;
  c.k1 = ex(-kLn10 * arg);
  arg = 471.78 * invtk + 25.9290 - 3.16967 * dlogtk - 0.01781 * salt_lim + 0.0001122 * s2;
  c.k2 = ex(-kLn10 * arg);

  // kb, Dickson 1990 (co2calc.F90:529-551)
  arg = (-8966.90 - 2890.53 * sqrts - 77.942 * salt_lim + 1.728 * salt_lim * sqrts - 0.0996 * s2) * invtk +
        (148.0248 + 137.1942 * sqrts + 1.62142 * salt_lim) +
        (-24.4344 - 25.085 * sqrts - 0.2474 * salt_lim) * dlogtk +
        0.053105 * sqrts * tk;
  // Pressure correction (Millero 1995).  Strict build: K = bexp(arg) * bexp(pressure term), the
  // reference's two factors.  Production build: one exponential of the summed argument
  // (identical up to the rounding of one addition, ~1e-16 relative * |arg|).
#ifdef BGC_STRICT
#define K_OF(arg_, dV_, Kap_) (ex(arg_) * (deep ? kfac((dV_), (Kap_), press_bar, invRtk, ex) : 1.0))
#else
#define K_OF(arg_, dV_, Kap_) ex((arg_) + (deep ? (-(dV_) + 0.5 * (Kap_) * press_bar) * press_bar * invRtk : 0.0))
#endif
  c.kb = K_OF(arg, -29.48 + (0.1622 - 0.002608 * temp) * temp, -2.84 * 0.001);
  // k1p, k2p, k3p: DOE 1994 (co2calc.F90:560-637)
  arg = -4576.752 * invtk + 115.525 - 18.453 * dlogtk +
        (-106.736 * invtk + 0.69171) * sqrts +
        (-0.65643 * invtk - 0.01844) * salt_lim;
  c.k1p = K_OF(arg, -14.51 + (0.1211 - 0.000321 * temp) * temp, (-2.67 + 0.0427 * temp) * 0.001);
  arg = -8814.715 * invtk + 172.0883 - 27.927 * dlogtk +
        (-160.340 * invtk + 1.3566) * sqrts +
        (0.37335 * invtk - 0.05778) * salt_lim;
  c.k2p = K_OF(arg, -23.12 + (0.1758 - 0.002647 * temp) * temp, (-5.15 + 0.09 * temp) * 0.001);
  arg = -3070.75 * invtk - 18.141 +
        (17.27039 * invtk + 2.81197) * sqrts +
        (-44.99486 * invtk - 0.09984) * salt_lim;
  c.k3p = K_OF(arg, -26.57 + (0.202 - 0.003042 * temp) * temp, (-4.08 + 0.0714 * temp) * 0.001);
  // ksi, Yao & Millero 1995 (co2calc.F90:647-669)
  arg = -8904.2 * invtk + 117.385 - 19.334 * dlogtk +
        (-458.79 * invtk + 3.5913) * sqrtis +
        (188.74 * invtk - 1.5998) * is +
        (-12.1652 * invtk + 0.07871) * is2 +
        log_1_m_1p005em3_s;
  c.ksi = K_OF(arg, -29.48 + (0.1622 - 0.002608 * temp) * temp, -2.84 * 0.001);
  // kw, Millero 1995 (co2calc.F90:681-700)
  arg = -13847.26 * invtk + 148.9652 - 23.6521 * dlogtk +
        (118.67 * invtk - 5.977 + 1.0495 * dlogtk) * sqrts -
        0.01615 * salt_lim;
  c.kw = K_OF(arg, -20.02 + (0.1119 - 0.001409 * temp) * temp, (-5.13 + 0.0794 * temp) * 0.001);
  // ks, Dickson 1990 (co2calc.F90:709-731)
  arg = -4276.1 * invtk + 141.328 - 23.093 * dlogtk +
        (-13856.0 * invtk + 324.57 - 47.986 * dlogtk) * sqrtis +
        (35474.0 * invtk - 771.54 + 114.723 * dlogtk) * is -
        2698.0 * invtk * is * sqrtis +
        1776.0 * invtk * is2 +
        log_1_m_1p005em3_s;
  c.ks = K_OF(arg, -18.03 + (0.0466 + 0.000316 * temp) * temp, (-4.53 + 0.09 * temp) * 0.001);

  // kf, Dickson & Riley 1979, uses the (corrected) ks (co2calc.F90:740-764)
  arg = 1.0 + fdiv((0.1400 / 96.062) * (scl), c.ks);
  const double log_1_p_tot_sulfate_div_ks = blog(arg);
  arg = 1590.2 * invtk - 12.641 + 1.525 * sqrtis + log_1_m_1p005em3_s + log_1_p_tot_sulfate_div_ks;
  c.kf = K_OF(arg, -9.78 - (0.009 + 0.000942 * temp) * temp, (-3.91 + 0.054 * temp) * 0.001);
#undef K_OF

  c.bt = 0.000232 / 10.811 * scl;   // co2calc.F90:773-775
  c.st = 0.14 / 96.062 * scl;
  c.ft = 0.000067 / 18.9984 * scl;
}

// Quantities of talk_row that do not depend on x: hoisted out of the iteration.
struct TalkInv {
  double k12, k12p, k123p, c, c_r, cks_;   // cks_ = c * ks
};

__device__ __forceinline__ TalkInv talk_invariants(const Co3Consts &k) {
  TalkInv t;
  t.k12 = k.k1 * k.k2;
  t.k12p = k.k1p * k.k2p;
  t.k123p = t.k12p * k.k3p;
  t.c = 1.0 + fdiv(k.st, k.ks);
  t.c_r = frcp(t.c);
  t.cks_ = t.c * k.ks;
  return t;
}

// Total-alkalinity residual fn(x) and d fn/dx, same term order as the reference.
__device__ __forceinline__ void talk_residual(const Co3Consts &k, const Co3Totals &t, const TalkInv &v,
                                              double x, double &fn, double &df) {
  const double x1 = x;
  const double x2 = x1 * x1;
  const double x3 = x2 * x1;
  const double a = x3 + k.k1p * x2 + v.k12p * x1 + v.k123p;
  const double da = 3.0 * x2 + 2.0 * k.k1p * x1 + v.k12p;
  const double b = x2 + k.k1 * x1 + v.k12;
  const double db = 2.0 * x1 + k.k1;
#ifdef BGC_STRICT
  const double x1_r = frcp(x1);
  const double a_r = frcp(a);
  const double b_r = frcp(b);
  const double kb_p_x1_r = frcp(k.kb + x1);
  const double ksi_p_x1_r = frcp(k.ksi + x1);
  const double c1_p_c_ks_x1_r_r = frcp(1.0 + v.cks_ * x1_r);
  const double c1_p_kf_x1_r_r = frcp(1.0 + k.kf * x1_r);
#else
  // The seven reciprocals of one evaluation from ONE division (batch inversion: prefix
  // products, one reciprocal, back-substitution).  1/(1 + c/x) is taken as x/(x + c).  The
  // product of the seven denominators stays within 1e-120..1e+120 for any bracket the
  // growth loop can produce, far from FP64 under/overflow.
  const double d3 = k.kb + x1, d4 = k.ksi + x1, d5 = x1 + v.cks_, d6 = x1 + k.kf;
  const double p1 = x1 * a, p2 = p1 * b, p3 = p2 * d3, p4 = p3 * d4, p5 = p4 * d5, p6 = p5 * d6;
  double r = frcp(p6);
  const double d6_r = r * p5; r = r * d6;
  const double d5_r = r * p4; r = r * d5;
  const double ksi_p_x1_r = r * p3; r = r * d4;
  const double kb_p_x1_r = r * p2; r = r * d3;
  const double b_r = r * p1; r = r * b;
  const double a_r = r * x1;
  const double x1_r = r * a;
  const double c1_p_c_ks_x1_r_r = x1 * d5_r;
  const double c1_p_kf_x1_r_r = x1 * d6_r;
#endif
  const double x2_r = x1_r * x1_r;
  const double a2_r = a_r * a_r;
  const double b2_r = b_r * b_r;

  fn = k.k1 * t.dic * x1 * b_r
     + 2.0 * t.dic * v.k12 * b_r
     + k.bt * k.kb * kb_p_x1_r
     + k.kw * x1_r
     + t.pt * v.k12p * x1 * a_r
     + 2.0 * t.pt * v.k123p * a_r
     + t.sit * k.ksi * ksi_p_x1_r
     - x1 * v.c_r
     - k.st * c1_p_c_ks_x1_r_r
     - k.ft * c1_p_kf_x1_r_r
     - t.pt * x3 * a_r
     - t.ta;

  df = k.k1 * t.dic * (b - x1 * db) * b2_r
     - 2.0 * t.dic * v.k12 * db * b2_r
     - k.bt * k.kb * kb_p_x1_r * kb_p_x1_r
     - k.kw * x2_r
     + (t.pt * v.k12p * (a - x1 * da)) * a2_r
     - 2.0 * t.pt * v.k123p * da * a2_r
     - t.sit * k.ksi * ksi_p_x1_r * ksi_p_x1_r
     - 1.0 * v.c_r
     - k.st * c1_p_c_ks_x1_r_r * c1_p_c_ks_x1_r_r * (v.cks_ * x2_r)
     - k.ft * c1_p_kf_x1_r_r * c1_p_kf_x1_r_r * k.kf * x2_r
     - t.pt * x2 * (3.0 * a - x1 * da) * a2_r;
}

__device__ __forceinline__ Co3Totals co3_totals(double dic_in, double ta_in, double pt_in, double sit_in) {
  Co3Totals t;   // co2calc.F90:843-846
  t.dic = gmax(dic_in, kDicMin) * kVolToMass;
  t.ta = gmax(ta_in, kAlkMin) * kVolToMass;
  t.pt = gmax(pt_in, 0.0) * kVolToMass;
  t.sit = gmax(sit_in, 0.0) * kVolToMass;
  return t;
}

// Status bits returned by the solver (accumulated into BgcStatus by the caller).
constexpr unsigned kSolveNoBracket = 1u;
constexpr unsigned kSolveNoConvergence = 2u;

// comp_htotal + drtsafe_row.  MUST be called by all 32 lanes of the warp
// (lanes without work pass benign inputs).  Returns htotal.
__device__ __forceinline__ double solve_htotal(const Co3Consts &k, const Co3Totals &t,
                                               double phlo, double phhi, unsigned &status) {
  const TalkInv v = talk_invariants(k);
#ifdef BGC_STRICT
  double x1 = exp10(-phhi);   // c10 ** (-phhi), co2calc.F90:848-849
  double x2 = exp10(-phlo);
#else
  double x1 = bexp(-phhi * kLn10);
  double x2 = bexp(-phlo * kLn10);
#endif

  double flo, fhi, f, df;
  talk_residual(k, t, v, x1, flo, df);
  talk_residual(k, t, v, x2, fhi, df);

  // bracket growth (co2calc.F90:920-938); not observed for oceanic inputs
  bool same_sign = (flo > 0.0 && fhi > 0.0) || (flo < 0.0 && fhi < 0.0);
  int grow = 0;
  while (__any_sync(FULL_MASK, same_sign)) {
    if (++grow > kBracketGrowCap) {
      if (same_sign) status |= kSolveNoBracket;
      break;
    }
    if (same_sign) {
      const double dxg = sqrt(fdiv(x2, x1));
      x2 = x2 * dxg;
      x1 = fdiv(x1, dxg);
      talk_residual(k, t, v, x1, flo, df);
      talk_residual(k, t, v, x2, fhi, df);
      same_sign = (flo > 0.0 && fhi > 0.0) || (flo < 0.0 && fhi < 0.0);
    }
  }

  double xlo, xhi;
  if (flo < 0.0) { xlo = x1; xhi = x2; } else { xlo = x2; xhi = x1; }
  double soln = 0.5 * (xlo + xhi);
  double dxold = fabs(xlo - xhi);
  double dx = dxold;

  talk_residual(k, t, v, soln, f, df);

  // co2calc.F90:960-991 with the reference's `mask` as the per-lane `live` flag
  bool live = true;
  for (int it = 1; it <= kMaxIt; ++it) {
    if (live) {
      const bool leave_bracket = ((soln - xhi) * df - f) * ((soln - xlo) * df - f) >= 0.0;
      const bool dx_decrease = fabs(2.0 * f) <= fabs(dxold * df);
      dxold = dx;
      if (leave_bracket || !dx_decrease) {
        dx = 0.5 * (xhi - xlo);
        soln = xlo + dx;
        if (xlo == soln) live = false;
      } else {
        dx = fdiv(-f, df);
        const double prev = soln;
        soln = soln + dx;
        if (prev == soln) live = false;
      }
      if (fabs(dx) < kXacc) live = false;
    }
    if (__all_sync(FULL_MASK, !live)) break;

    double fn, dfn;   // evaluated by every lane to keep the warp on one path
    talk_residual(k, t, v, soln, fn, dfn);
    if (live) {
      f = fn; df = dfn;
      if (f < 0.0) xlo = soln; else xhi = soln;
    }
  }
  if (live) status |= kSolveNoConvergence;   // reference: silent fall-through (co2calc.F90:993-995)
  return soln;
}

// comp_co3_sat_vals, Mucci 1983 + Millero 1979 (co2calc.F90:1096-1238)
template <class EXP>
__device__ __forceinline__ void co3_sat_vals(bool deep, double depth, double temp, double salt,
                                             double &co3_sat_calc, double &co3_sat_arag, const EXP &ex) {
  const double press_bar = press_bar_of_depth(depth, ex);
  const double salt_lim = gmax(salt, kSaltMin);
  const double tk = kT0Kelvin + temp;
  const double log10tk = cdiv(blog(tk), kLn10, 1.0 / kLn10);   // :1161-1164
  const double invtk = frcp(tk);
  const double invRtk = (1.0 / 83.1451) * invtk;
  const double sqrts = sqrt(salt_lim);
  const double s15 = sqrts * salt_lim;

  double arg = -171.9065 - 0.077993 * tk + 2839.319 * invtk + 71.595 * log10tk +
               (-0.77712 + 0.0028426 * tk + 178.34 * invtk) * sqrts -
               0.07711 * salt_lim + 0.0041249 * s15;
  const double arg_calc = kLn10 * arg;
  arg = -171.945 - 0.077993 * tk + 2903.293 * invtk + 71.595 * log10tk +
        (-0.068393 + 0.0017276 * tk + 88.135 * invtk) * sqrts -
        0.10018 * salt_lim + 0.0059415 * s15;
  const double arg_arag = kLn10 * arg;
  const double deltaV = -48.76 + 0.5304 * temp;
  const double Kappa = (-11.76 + 0.3692 * temp) * 0.001;
#ifdef BGC_STRICT
  double K_calc = ex(arg_calc), K_arag = ex(arg_arag);
  if (deep) {
    K_calc *= kfac(deltaV, Kappa, press_bar, invRtk, ex);
    K_arag *= kfac(deltaV + 2.8, Kappa, press_bar, invRtk, ex);
  }
#else   // one exponential of the summed argument, as in co3_coeffs
  const double K_calc = ex(arg_calc + (deep ? (-deltaV + 0.5 * Kappa * press_bar) * press_bar * invRtk : 0.0));
  const double K_arag = ex(arg_arag + (deep ? (-(deltaV + 2.8) + 0.5 * Kappa * press_bar) * press_bar * invRtk : 0.0));
#endif

  const double inv_Ca = fdiv((35.0 / 0.01028), salt_lim);
  co3_sat_calc = (K_calc * inv_Ca) * kMassToVol;
  co3_sat_arag = (K_arag * inv_Ca) * kMassToVol;
}

}  // namespace bgc
