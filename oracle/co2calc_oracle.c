/* co2calc_oracle.c — restatement of module co2calc (co2calc.F90), default
 * (non-CCSMCOUPLED) build: intrinsics EXP/LOG/SQRT -> libm.
 * TEST INFRASTRUCTURE ONLY (see bgc_oracle.h).  Pinned bit for bit against the machine-translated reference (bgc_oracle.h).
 *
 * Every expression keeps the reference's left-to-right evaluation order;
 * compile with -O2 -ffp-contract=off. */
#include "bgc_oracle.h"
#include <math.h>
#include <stddef.h>

/* co2calc.F90:30-59 */
static const double c0 = 0.0, c1 = 1.0, c2 = 2.0, c3 = 3.0, c10 = 10.0, c1000 = 1000.0,
                    p5 = 0.5, p001 = 0.001;
static const double rho_sw = 1.026;
static const double T0_Kelvin = 273.15;
static const double xacc_parm = 1e-10;
#define MAX_BRACKET_GROW_IT 3   /* :54 (documented intent; the abort is commented out) */
#define MAXIT 100               /* :55 */
/* the reference's bracket-growth DO loop has no exit (:920-938, abort commented
 * out).  For finite inputs a sign change is always reached because the bracket
 * ratio squares every pass; cap it so pathological inputs cannot hang. */
#define BRACKET_GROW_CAP 64
static const double salt_min = 0.1;
#define DIC_MIN (salt_min / 35.0 * 1944.0)
#define ALK_MIN (salt_min / 35.0 * 2225.0)

/* co2calc.F90:320-777 */
void oracle_comp_co3_coeffs(int k, double depth, double temp, double salt,
                            double *sk0, double *sk1, double *sk2, double *sff,
                            int k1_k2_pH_tot, OracleCo2Save *sv) {
  double k0, k1, k2, ff;
  double press_bar;
  double salt_lim, tk, is, scl, tk100, tk1002, invtk, dlogtk, is2, sqrtis, s2, sqrts,
         invRtk, arg, deltaV, Kappa, lnKfac, Kfac, log_1_m_1p005em3_s,
         log_1_p_tot_sulfate_div_ks;

  /* :371-372 */
  press_bar = 0.059808 * (exp(-0.025 * depth) - c1) + 0.100766 * depth +
              2.28405e-7 * (depth * depth);

  /* :386-415 */
  salt_lim = fmax(salt, salt_min);
  tk = T0_Kelvin + temp;
  tk100 = tk * 1e-2;
  tk1002 = tk100 * tk100;
  invtk = c1 / tk;
  dlogtk = log(tk);
  invRtk = (c1 / 83.1451) * invtk;

  is = 19.924 * salt_lim / (c1000 - 1.005 * salt_lim);
  is2 = is * is;
  sqrtis = sqrt(is);
  sqrts = sqrt(salt_lim);
  s2 = salt_lim * salt_lim;
  scl = salt_lim / 1.80655;

  arg = c1 - 0.001005 * salt_lim;
  log_1_m_1p005em3_s = log(arg);

  /* ff, Weiss & Price 1980 :423-431 */
  arg = -162.8301 + 218.2968 / tk100 + 90.9241 * (dlogtk + log(1e-2)) - 1.47696 * tk1002 +
        salt_lim * (.025695 - .025225 * tk100 + 0.0049867 * tk1002);
  ff = exp(arg);
  *sff = ff;

  /* K0, Weiss 1974 :437-444 */
  arg = 93.4517 / tk100 - 60.2409 + 23.3585 * (dlogtk + log(1e-2)) +
        salt_lim * (.023517 - 0.023656 * tk100 + 0.0047036 * tk1002);
  k0 = exp(arg);
  *sk0 = k0;

  /* k1 :461-490.  NB (Q2): sk1 is captured BEFORE the pressure correction. */
  if (k1_k2_pH_tot) {
    arg = 3633.86 * invtk - 61.2172 + 9.67770 * dlogtk - 0.011555 * salt_lim + 0.0001152 * s2;
  } else {
    arg = 3670.7 * invtk - 62.008 + 9.7944 * dlogtk - 0.0118 * salt_lim + 0.000116 * s2;
  }
  arg = -log(c10) * arg;
  k1 = exp(arg);
  *sk1 = k1;

  if (k > 1) {
    deltaV = -25.5 + 0.1271 * temp;
    Kappa = (-3.08 + 0.0877 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    k1 = k1 * Kfac;   /* dead: never leaves this routine */
  }

  /* k2 :492-519 */
  if (k1_k2_pH_tot) {
    arg = 471.78 * invtk + 25.9290 - 3.16967 * dlogtk - 0.01781 * salt_lim + 0.0001122 * s2;
  } else {
    arg = 1394.7 * invtk + 4.777 - 0.0184 * salt_lim + 0.000118 * s2;
  }
  arg = -log(c10) * arg;
  k2 = exp(arg);
  *sk2 = k2;

  if (k > 1) {
    deltaV = -15.82 - 0.0219 * temp;
    Kappa = (1.13 - 0.1475 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    k2 = k2 * Kfac;   /* dead */
  }
  (void)k1; (void)k2;

  /* kb :529-551 */
  arg = (-8966.90 - 2890.53 * sqrts - 77.942 * salt_lim + 1.728 * salt_lim * sqrts -
         0.0996 * s2) * invtk +
        (148.0248 + 137.1942 * sqrts + 1.62142 * salt_lim) +
        (-24.4344 - 25.085 * sqrts - 0.2474 * salt_lim) * dlogtk +
        0.053105 * sqrts * tk;
  sv->kb = exp(arg);
  if (k > 1) {
    deltaV = -29.48 + (0.1622 - 0.002608 * temp) * temp;
    Kappa = -2.84 * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->kb = sv->kb * Kfac;
  }

  /* k1p :560-580 */
  arg = -4576.752 * invtk + 115.525 - 18.453 * dlogtk +
        (-106.736 * invtk + 0.69171) * sqrts +
        (-0.65643 * invtk - 0.01844) * salt_lim;
  sv->k1p = exp(arg);
  if (k > 1) {
    deltaV = -14.51 + (0.1211 - 0.000321 * temp) * temp;
    Kappa = (-2.67 + 0.0427 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->k1p = sv->k1p * Kfac;
  }

  /* k2p :589-609 */
  arg = -8814.715 * invtk + 172.0883 - 27.927 * dlogtk +
        (-160.340 * invtk + 1.3566) * sqrts +
        (0.37335 * invtk - 0.05778) * salt_lim;
  sv->k2p = exp(arg);
  if (k > 1) {
    deltaV = -23.12 + (0.1758 - 0.002647 * temp) * temp;
    Kappa = (-5.15 + 0.09 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->k2p = sv->k2p * Kfac;
  }

  /* k3p :618-637 */
  arg = -3070.75 * invtk - 18.141 +
        (17.27039 * invtk + 2.81197) * sqrts +
        (-44.99486 * invtk - 0.09984) * salt_lim;
  sv->k3p = exp(arg);
  if (k > 1) {
    deltaV = -26.57 + (0.202 - 0.003042 * temp) * temp;
    Kappa = (-4.08 + 0.0714 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->k3p = sv->k3p * Kfac;
  }

  /* ksi :647-669 */
  arg = -8904.2 * invtk + 117.385 - 19.334 * dlogtk +
        (-458.79 * invtk + 3.5913) * sqrtis +
        (188.74 * invtk - 1.5998) * is +
        (-12.1652 * invtk + 0.07871) * is2 +
        log_1_m_1p005em3_s;
  sv->ksi = exp(arg);
  if (k > 1) {
    deltaV = -29.48 + (0.1622 - 0.002608 * temp) * temp;
    Kappa = -2.84 * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->ksi = sv->ksi * Kfac;
  }

  /* kw :681-700 */
  arg = -13847.26 * invtk + 148.9652 - 23.6521 * dlogtk +
        (118.67 * invtk - 5.977 + 1.0495 * dlogtk) * sqrts -
        0.01615 * salt_lim;
  sv->kw = exp(arg);
  if (k > 1) {
    deltaV = -20.02 + (0.1119 - 0.001409 * temp) * temp;
    Kappa = (-5.13 + 0.0794 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->kw = sv->kw * Kfac;
  }

  /* ks :709-731 */
  arg = -4276.1 * invtk + 141.328 - 23.093 * dlogtk +
        (-13856.0 * invtk + 324.57 - 47.986 * dlogtk) * sqrtis +
        (35474.0 * invtk - 771.54 + 114.723 * dlogtk) * is -
        2698.0 * invtk * is * sqrtis +
        1776.0 * invtk * is2 +
        log_1_m_1p005em3_s;
  sv->ks = exp(arg);
  if (k > 1) {
    deltaV = -18.03 + (0.0466 + 0.000316 * temp) * temp;
    Kappa = (-4.53 + 0.09 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->ks = sv->ks * Kfac;
  }

  /* kf :740-764 */
  arg = c1 + (0.1400 / 96.062) * (scl) / sv->ks;
  log_1_p_tot_sulfate_div_ks = log(arg);
  arg = 1590.2 * invtk - 12.641 + 1.525 * sqrtis +
        log_1_m_1p005em3_s + log_1_p_tot_sulfate_div_ks;
  sv->kf = exp(arg);
  if (k > 1) {
    deltaV = -9.78 - (0.009 + 0.000942 * temp) * temp;
    Kappa = (-3.91 + 0.054 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    sv->kf = sv->kf * Kfac;
  }

  /* :773-775 */
  sv->bt = 0.000232 / 10.811 * scl;
  sv->st = 0.14 / 96.062 * scl;
  sv->ft = 0.000067 / 18.9984 * scl;
}

/* co2calc.F90:1001-1092 */
void oracle_talk_row(double k1, double k2, double x, double *fn, double *df,
                     const OracleCo2Save *sv) {
  double x1, x1_r, x2, x2_r, x3, k12, k12p, k123p, a, a_r, a2_r, da, b, b_r, b2_r, db, c,
         c_r, kb_p_x1_r, ksi_p_x1_r, c1_p_c_ks_x1_r_r, c1_p_kf_x1_r_r;
  const double k1p = sv->k1p, k2p = sv->k2p, k3p = sv->k3p, kb = sv->kb, ksi = sv->ksi,
               kw = sv->kw, ks = sv->ks, kf = sv->kf, bt = sv->bt, st = sv->st, ft = sv->ft,
               dic = sv->dic, ta = sv->ta, pt = sv->pt, sit = sv->sit;

  x1 = x;
  x1_r = c1 / x1;
  x2 = x1 * x1;
  x2_r = x1_r * x1_r;
  x3 = x2 * x1;
  k12 = k1 * k2;
  k12p = k1p * k2p;
  k123p = k12p * k3p;
  a = x3 + k1p * x2 + k12p * x1 + k123p;
  a_r = c1 / a;
  a2_r = a_r * a_r;
  da = c3 * x2 + c2 * k1p * x1 + k12p;
  b = x2 + k1 * x1 + k12;
  b_r = c1 / b;
  b2_r = b_r * b_r;
  db = c2 * x1 + k1;
  c = c1 + st / ks;
  c_r = c1 / c;
  kb_p_x1_r = c1 / (kb + x1);
  ksi_p_x1_r = c1 / (ksi + x1);
  c1_p_c_ks_x1_r_r = c1 / (c1 + c * ks * x1_r);
  c1_p_kf_x1_r_r = c1 / (c1 + kf * x1_r);

  /* :1063-1074 */
  *fn = k1 * dic * x1 * b_r
        + c2 * dic * k12 * b_r
        + bt * kb * kb_p_x1_r
        + kw * x1_r
        + pt * k12p * x1 * a_r
        + c2 * pt * k123p * a_r
        + sit * ksi * ksi_p_x1_r
        - x1 * c_r
        - st * c1_p_c_ks_x1_r_r
        - ft * c1_p_kf_x1_r_r
        - pt * x3 * a_r
        - ta;

  /* :1080-1090 */
  *df = k1 * dic * (b - x1 * db) * b2_r
        - c2 * dic * k12 * db * b2_r
        - bt * kb * kb_p_x1_r * kb_p_x1_r
        - kw * x2_r
        + (pt * k12p * (a - x1 * da)) * a2_r
        - c2 * pt * k123p * da * a2_r
        - sit * ksi * ksi_p_x1_r * ksi_p_x1_r
        - c1 * c_r
        - st * c1_p_c_ks_x1_r_r * c1_p_c_ks_x1_r_r * (c * ks * x2_r)
        - ft * c1_p_kf_x1_r_r * c1_p_kf_x1_r_r * kf * x2_r
        - pt * x2 * (c3 * a - x1 * da) * a2_r;
}

/* co2calc.F90:872-997 */
void oracle_drtsafe_row(int k, double k1, double k2, double *x1, double *x2, double xacc,
                        double *soln, const OracleCo2Save *sv, OracleSolverStats *st) {
  int leave_bracket, dx_decrease, mask;
  int it;
  double temp;
  double xlo, xhi, flo, fhi, f, df, dxold, dx;
  (void)k;

  it = 0;
  for (;;) {   /* :920-938 */
    oracle_talk_row(k1, k2, *x1, &flo, &df, sv);
    oracle_talk_row(k1, k2, *x2, &fhi, &df, sv);
    if (st) st->talk_row_calls += 2;

    mask = (flo > c0 && fhi > c0) || (flo < c0 && fhi < c0);
    if (!mask) break;

    it = it + 1;
    if (st) st->bracket_grow += 1;
    if (it > BRACKET_GROW_CAP) break;   /* reference: abort commented out, loops on */

    dx = sqrt(*x2 / *x1);
    *x2 = *x2 * dx;
    *x1 = *x1 / dx;
  }

  if (flo < c0) {   /* :940-949 */
    xlo = *x1;
    xhi = *x2;
  } else {
    xlo = *x2;
    xhi = *x1;
    temp = flo;
    flo = fhi;
    fhi = temp;
  }
  *soln = p5 * (xlo + xhi);
  dxold = fabs(xlo - xhi);
  dx = dxold;

  oracle_talk_row(k1, k2, *soln, &f, &df, sv);
  if (st) st->talk_row_calls += 1;

  mask = 1;   /* :960-991 */
  for (it = 1; it <= MAXIT; ++it) {
    leave_bracket = ((*soln - xhi) * df - f) * ((*soln - xlo) * df - f) >= 0;
    dx_decrease = fabs(c2 * f) <= fabs(dxold * df);
    if (leave_bracket || !dx_decrease) {
      dxold = dx;
      dx = p5 * (xhi - xlo);
      *soln = xlo + dx;
      if (xlo == *soln) mask = 0;
    } else {
      dxold = dx;
      dx = -f / df;
      temp = *soln;
      *soln = *soln + dx;
      if (temp == *soln) mask = 0;
    }
    if (fabs(dx) < xacc) mask = 0;
    if (st) st->newton_iters += 1;

    if (!mask) return;

    oracle_talk_row(k1, k2, *soln, &f, &df, sv);
    if (st) st->talk_row_calls += 1;

    if (f < c0) {
      xlo = *soln;
      flo = f;
    } else {
      xhi = *soln;
      fhi = f;
    }
  }
  (void)flo; (void)fhi;
  if (st) st->no_convergence += 1;   /* reference: silent fall-through (:993-995) */
}

/* co2calc.F90:781-868 */
void oracle_comp_htotal(int k, double temp, double dic_in, double ta_in, double pt_in,
                        double sit_in, double k1, double k2, double *phlo, double *phhi,
                        double *htotal, OracleCo2Save *sv, OracleSolverStats *st) {
  double mass_to_vol, vol_to_mass, x1, x2;
  (void)temp;

  mass_to_vol = 1e6 * rho_sw;
  vol_to_mass = c1 / mass_to_vol;

  sv->dic = fmax(dic_in, DIC_MIN) * vol_to_mass;   /* :843-846 */
  sv->ta = fmax(ta_in, ALK_MIN) * vol_to_mass;
  sv->pt = fmax(pt_in, c0) * vol_to_mass;
  sv->sit = fmax(sit_in, c0) * vol_to_mass;

  x1 = pow(c10, -*phhi);   /* :848-849 */
  x2 = pow(c10, -*phlo);

  oracle_drtsafe_row(k, k1, k2, &x1, &x2, xacc_parm, htotal, sv, st);
}

/* co2calc.F90:75-210 */
void oracle_co2calc_1point(double depth, int locmip_k1_k2_bug_fix, int lcomp_co3_coeffs,
                           double temp, double salt, double dic_in, double ta_in,
                           double pt_in, double sit_in, double *phlo, double *phhi,
                           double *ph, double xco2_in, double atmpres, double *co2star,
                           double *dco2star, double *pCO2surf, double *dpco2,
                           OracleSolverStats *st) {
  OracleCo2Save sv;
  int k;
  double mass_to_vol, vol_to_mass, co2starair, htotal2;
  double press_bar, xco2, htotal, k0 = 0, k1 = 0, k2 = 0, ff = 0;

  mass_to_vol = 1e6 * rho_sw;
  vol_to_mass = c1 / mass_to_vol;
  (void)vol_to_mass;

  k = 1;   /* :149 */

  press_bar = 0.059808 * (exp(-0.025 * depth) - c1) + 0.100766 * depth +
              2.28405e-7 * (depth * depth);   /* :156-157 */

  if (lcomp_co3_coeffs) {   /* Q4: press_bar is passed as `depth` (:160) */
    oracle_comp_co3_coeffs(k, press_bar, temp, salt, &k0, &k1, &k2, &ff,
                           locmip_k1_k2_bug_fix, &sv);
  }

  oracle_comp_htotal(k, temp, dic_in, ta_in, pt_in, sit_in, k1, k2, phlo, phhi, &htotal,
                     &sv, st);

  xco2 = xco2_in * 1e-6;   /* :175 */

  htotal2 = htotal * htotal;   /* :184-189 */
  *co2star = sv.dic * htotal2 / (htotal2 + k1 * htotal + k1 * k2);
  co2starair = xco2 * ff * atmpres;
  *dco2star = co2starair - *co2star;
  *ph = -log10(htotal);

  *pCO2surf = *co2star / ff;   /* :196-197 */
  *dpco2 = *pCO2surf - xco2 * atmpres;

  *co2star = *co2star * mass_to_vol;   /* :204-208 */
  *dco2star = *dco2star * mass_to_vol;
  *pCO2surf = *pCO2surf * 1e6;
  *dpco2 = *dpco2 * 1e6;
}

/* co2calc.F90:214-316 */
void oracle_comp_CO3terms(int k, double depth, int lcomp_co3_coeffs, double temp,
                          double salt, double dic_in, double ta_in, double pt_in,
                          double sit_in, double *phlo, double *phhi, double *pH,
                          double *H2CO3, double *HCO3, double *CO3, OracleSolverStats *st) {
  OracleCo2Save sv;
  double mass_to_vol, htotal2, denom;
  double htotal, k0 = 0, k1 = 0, k2 = 0, ff = 0;

  mass_to_vol = 1e6 * rho_sw;

  if (lcomp_co3_coeffs) {   /* :284-286 */
    oracle_comp_co3_coeffs(k, depth, temp, salt, &k0, &k1, &k2, &ff, 1, &sv);
  }

  oracle_comp_htotal(k, temp, dic_in, ta_in, pt_in, sit_in, k1, k2, phlo, phhi, &htotal,
                     &sv, st);

  htotal2 = htotal * htotal;   /* :301-306 */
  denom = c1 / (htotal2 + k1 * htotal + k1 * k2);
  *H2CO3 = sv.dic * htotal2 * denom;
  *HCO3 = sv.dic * k1 * htotal * denom;
  *CO3 = sv.dic * k1 * k2 * denom;
  *pH = -log10(htotal);

  *H2CO3 = *H2CO3 * mass_to_vol;   /* :312-314 */
  *HCO3 = *HCO3 * mass_to_vol;
  *CO3 = *CO3 * mass_to_vol;
}

/* co2calc.F90:1096-1238 */
void oracle_comp_co3_sat_vals(int k, double depth, double temp, double salt,
                              double *co3_sat_calc, double *co3_sat_arag) {
  double mass_to_vol, press_bar;
  double salt_lim, tk, log10tk, invtk, sqrts, s15, invRtk, arg, K_calc, K_arag,
         deltaV = 0, Kappa = 0, lnKfac, Kfac, inv_Ca;

  mass_to_vol = 1e6 * rho_sw;

  press_bar = 0.059808 * (exp(-0.025 * depth) - c1) + 0.100766 * depth +
              2.28405e-7 * (depth * depth);   /* :1153-1154 */

  salt_lim = fmax(salt, salt_min);
  tk = T0_Kelvin + temp;
  log10tk = log(tk);
  log10tk = log10tk / log(c10);   /* :1164 */
  invtk = c1 / tk;
  invRtk = (c1 / 83.1451) * invtk;

  sqrts = sqrt(salt_lim);
  s15 = sqrts * salt_lim;

  /* :1180-1188 */
  arg = -171.9065 - 0.077993 * tk + 2839.319 * invtk + 71.595 * log10tk +
        (-0.77712 + 0.0028426 * tk + 178.34 * invtk) * sqrts -
        0.07711 * salt_lim + 0.0041249 * s15;
  arg = log(c10) * arg;
  K_calc = exp(arg);

  if (k > 1) {   /* :1190-1200 */
    deltaV = -48.76 + 0.5304 * temp;
    Kappa = (-11.76 + 0.3692 * temp) * p001;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    K_calc = K_calc * Kfac;
  }

  /* :1202-1210 */
  arg = -171.945 - 0.077993 * tk + 2903.293 * invtk + 71.595 * log10tk +
        (-0.068393 + 0.0017276 * tk + 88.135 * invtk) * sqrts -
        0.10018 * salt_lim + 0.0059415 * s15;
  arg = log(c10) * arg;
  K_arag = exp(arg);

  if (k > 1) {   /* :1212-1221 */
    deltaV = deltaV + 2.8;
    lnKfac = (-deltaV + p5 * Kappa * press_bar) * press_bar * invRtk;
    Kfac = exp(lnKfac);
    K_arag = K_arag * Kfac;
  }

  inv_Ca = (35.0 / 0.01028) / salt_lim;   /* :1227-1229 */
  *co3_sat_calc = K_calc * inv_Ca;
  *co3_sat_arag = K_arag * inv_Ca;

  *co3_sat_calc = *co3_sat_calc * mass_to_vol;   /* :1235-1236 */
  *co3_sat_arag = *co3_sat_arag * mass_to_vol;
}

/* ---------------------------------------------------------------- batched drivers */

void oracle_co2calc_points(int n, const double *depth, const double *temp,
                           const double *salt, const double *dic, const double *ta,
                           const double *pt, const double *sit, const double *phlo,
                           const double *phhi, const double *xco2, const double *atmpres,
                           double *ph, double *co2star, double *dco2star,
                           double *pco2surf, double *dpco2, OracleSolverStats *st,
                           int nthreads) {
  long tr = 0, bg = 0, ni = 0, nc = 0;
  int i;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) reduction(+ : tr, bg, ni, nc) schedule(static)
  for (i = 0; i < n; ++i) {
    OracleSolverStats s = {0, 0, 0, 0};
    double lo = phlo[i], hi = phhi[i];
    oracle_co2calc_1point(depth[i], 1, 1, temp[i], salt[i], dic[i], ta[i], pt[i], sit[i],
                          &lo, &hi, &ph[i], xco2[i], atmpres[i], &co2star[i], &dco2star[i],
                          &pco2surf[i], &dpco2[i], &s);
    tr += s.talk_row_calls; bg += s.bracket_grow; ni += s.newton_iters; nc += s.no_convergence;
  }
  if (st) { st->talk_row_calls += tr; st->bracket_grow += bg; st->newton_iters += ni;
            st->no_convergence += nc; }
}

void oracle_co3_coeffs_points(int n, const int *k, const double *depth, const double *temp,
                              const double *salt, double *out) {
  int i;
  for (i = 0; i < n; ++i) {
    OracleCo2Save sv;
    double k0, k1, k2, ff;
    double *o = out + (size_t)14 * i;
    oracle_comp_co3_coeffs(k[i], depth[i], temp[i], salt[i], &k0, &k1, &k2, &ff, 1, &sv);
    o[0] = k0; o[1] = k1; o[2] = k2; o[3] = ff; o[4] = sv.kw; o[5] = sv.kb; o[6] = sv.ks;
    o[7] = sv.kf; o[8] = sv.k1p; o[9] = sv.k2p; o[10] = sv.k3p; o[11] = sv.ksi;
    o[12] = sv.bt; o[13] = sv.st;
  }
}
