/* bgc_mod_oracle.c — restatement of module BGC_mod (BGC_mod.F90):
 * BGC_SourceSink, init_particulate_terms, compute_particulate_terms,
 * BGC_SurfaceFluxes, SCHMIDT_O2/CO2_singleValue, O2SAT_singleValue.
 * TEST INFRASTRUCTURE ONLY (see bgc_oracle.h).  Pinned bit for bit against the machine-translated reference (bgc_oracle.h).
 *
 * Arrays keep the reference's Fortran layout (level fastest).  Indices in the
 * macros below are 1-based like the Fortran they restate.  Expressions keep the
 * reference's evaluation order; compile with -O2 -ffp-contract=off. */
#include "bgc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ constants */
/* BGC_mod.F90:120-125 declares c0,c1,c2,c10,p5 as INTEGER(8) (Q5); every use is
 * value-equivalent to the doubles below (p5 -> 0 is never used). */
static const double c0 = 0.0, c1 = 1.0, c2 = 2.0, c10 = 10.0;

/* BGC_parms.F90:37-40 */
static const double spd = 86400.0;
#define dps (c1 / spd)
#define yps (c1 / (365.0 * spd))

/* BGC_parms.F90:327-339 */
#define parm_Red_D_C_P 117.0
#define parm_Red_D_N_P 16.0
#define parm_Red_D_O2_P 170.0
#define parm_Remin_D_O2_P 138.0
#define parm_Red_D_C_O2 (parm_Red_D_C_P / parm_Red_D_O2_P)
#define parm_Remin_D_C_O2 (parm_Red_D_C_P / parm_Remin_D_O2_P)
#define parm_Red_Fe_C 3.0e-6
#define parm_Red_D_C_O2_diaz (parm_Red_D_C_P / 150.0)

/* BGC_parms.F90:371-374 */
#define fe_scavenge_thres1 0.8e-3
#define fe_max_scale2 1200.0
/* :394-405 */
#define caco3_poc_min 0.4
#define spc_poc_fac 0.11
#define f_graze_sp_poc_lim 0.3
#define f_photosp_CaCO3 0.4
#define f_graze_CaCO3_remin 0.33
#define f_graze_si_remin 0.35
/* :411-412 */
#define r_Nfix_photo 1.25
/* :420-429 */
#define Q 0.137
#define Qp_zoo_pom 0.00855
#define Qfe_zoo 3.0e-6
#define gQsi_0 0.137
#define gQsi_max 0.685
#define gQsi_min 0.0457
#define QCaCO3_max 0.4
#define denitrif_C_N (parm_Red_D_C_P / 136.0)
/* :435-441 */
#define thres_z1 100.0e2
#define thres_z2 150.0e2
#define loss_thres_zoo 0.005
#define CaCO3_temp_thres1 6.0
#define CaCO3_temp_thres2 (-2.0)
#define CaCO3_sp_thres 4.0
/* :454-463 */
#define f_qsw_par 0.45
#define Tref 30.0
#define Q_10 1.5
/* :469-477 */
#define DOC_reminR ((c1 / 250.0) * dps)
#define DON_reminR ((c1 / 160.0) * dps)
#define DOFe_reminR ((c1 / 160.0) * dps)
#define DOP_reminR ((c1 / 160.0) * dps)
#define DONr_reminR ((c1 / (365.0 * 2.5)) * dps)
#define DOPr_reminR ((c1 / (365.0 * 2.5)) * dps)
#define DONrefract 0.08
#define DOPrefract 0.03
/* :488-489 */
#define xkw_coeff 8.6e-9

/* BGC_mod.F90:144-149 */
#define phlo_surf_init 7.0
#define phhi_surf_init 9.0
#define phlo_3d_init 6.0
#define phhi_3d_init 9.0
#define del_ph 0.20

#define mpercm 0.01

/* BGC_mod.F90:155-170 */
typedef struct sinking_particle {
  double diss, gamma, mass, rho;
  double sflux_in, hflux_in, prod, sflux_out, hflux_out, sed_loss, remin;
} sinking_particle;

/* Fortran-layout accessors, 1-based */
#define A2(p, k, c) ((p)[(size_t)((k)-1) + (size_t)nL * (size_t)((c)-1)])
#define A3(p, k, c, n) ((p)[(size_t)((k)-1) + (size_t)nL * ((size_t)((c)-1) + (size_t)nC * (size_t)((n)-1))])
#define F2(p, c, n) ((p)[(size_t)((c)-1) + (size_t)nC * (size_t)((n)-1)])
#define C1(p, c) ((p)[(size_t)((c)-1)])

static double sum4(const double *x) {   /* Fortran SUM over autotroph_cnt, in order */
  double s = 0.0;
  int i;
  for (i = 0; i < 4; ++i) s = s + x[i];
  return s;
}

/* BGC_mod.F90:2965-3005 */
double oracle_SCHMIDT_O2_singleValue(double SST) {
  const double a = 1638.0, b = 81.83, c = 1.483, d = 0.008004;
  return a + SST * (-b + SST * (c + SST * (-d)));
}

/* BGC_mod.F90:3091-3128 */
double oracle_SCHMIDT_CO2_singleValue(double SST) {
  const double a = 2073.1, b = 125.62, c = 3.6276, d = 0.043219;
  return a + SST * (-b + SST * (c + SST * (-d)));
}

/* BGC_mod.F90:3012-3083 */
double oracle_O2SAT_singleValue(double SST, double SSS, double T0_Kelvin_BGC) {
  const double a_0 = 2.00907, a_1 = 3.22014, a_2 = 4.05010, a_3 = 4.94457,
               a_4 = -2.56847E-1, a_5 = 3.88767, b_0 = -6.24523E-3, b_1 = -7.37614E-3,
               b_2 = -1.03410E-2, b_3 = -8.17083E-3, c_0 = -4.88682E-7;
  double TS, r;
  TS = log(((T0_Kelvin_BGC + 25.0) - SST) / (T0_Kelvin_BGC + SST));
  r = exp(a_0 + TS * (a_1 + TS * (a_2 + TS * (a_3 + TS * (a_4 + TS * a_5)))) +
          SSS * ((b_0 + TS * (b_1 + TS * (b_2 + TS * b_3))) + SSS * c_0));
  r = r / 0.0223916;
  return r;
}

/* BGC_mod.F90:2006-2109 */
static void init_particulate_terms(const BgcParams *p, sinking_particle *POC,
                                   sinking_particle *P_CaCO3, sinking_particle *P_SiO2,
                                   sinking_particle *dust, sinking_particle *P_iron,
                                   double *QA_dust_def, double NET_DUST_IN) {
  POC->diss = p->parm_POC_diss;
  POC->gamma = c0;
  POC->mass = 12.01;
  POC->rho = c0;

  P_CaCO3->diss = p->parm_CaCO3_diss;
  P_CaCO3->gamma = 0.30;
  P_CaCO3->mass = 100.09;
  P_CaCO3->rho = 0.05 * P_CaCO3->mass / POC->mass;

  P_SiO2->diss = p->parm_SiO2_diss;
  P_SiO2->gamma = 0.030;
  P_SiO2->mass = 60.08;
  P_SiO2->rho = 0.05 * P_SiO2->mass / POC->mass;

  dust->diss = 20000.0;
  dust->gamma = 0.97;
  dust->mass = 1.0e9;
  dust->rho = 0.05 * dust->mass / POC->mass;

  P_iron->diss = 60000.0;
  P_iron->gamma = c0;
  P_iron->mass = c0;
  P_iron->rho = c0;

  P_CaCO3->sflux_out = c0;
  P_CaCO3->hflux_out = c0;
  P_SiO2->sflux_out = c0;
  P_SiO2->hflux_out = c0;

  if (NET_DUST_IN != c0) {   /* :2081-2087 */
    dust->sflux_out = (c1 - dust->gamma) * NET_DUST_IN;
    dust->hflux_out = dust->gamma * NET_DUST_IN;
  } else {
    dust->sflux_out = c0;
    dust->hflux_out = c0;
  }

  P_iron->sflux_out = c0;
  P_iron->hflux_out = c0;
  POC->sflux_out = c0;
  POC->hflux_out = c0;

  *QA_dust_def = dust->rho * (dust->sflux_out + dust->hflux_out);   /* :2103-2104 */
}

/* BGC_mod.F90:2116-2699 */
static void compute_particulate_terms(const BgcParams *p, int column, int k, int kmax,
                                      sinking_particle *POC, sinking_particle *P_CaCO3,
                                      sinking_particle *P_SiO2, sinking_particle *dust,
                                      sinking_particle *P_iron, double *QA_dust_def,
                                      double TEMP, double O2_loc, double NO3_loc,
                                      double *SED_DENITRIF, double *OTHER_REMIN,
                                      double cell_thickness, double cell_bottom_depth,
                                      double FESEDFLUX_loc, BgcDiagnostics *d, int nL,
                                      long *poc_error_count) {
  double poc_diss, sio2_diss, caco3_diss, dust_diss;
  double work, TfuncS, scalelength = 0.0, DECAY_Hard, DECAY_HardDust;
  double decay_POC_E, decay_SiO2, decay_CaCO3, decay_dust, POC_PROD_avail,
         new_QA_dust_def, flux, flux_alt, dz_loc, dzr_loc;
  int n;
  int poc_error;
  const double dust_to_Fe = 0.035 / 55.847 * 1.0e9;   /* BGC_parms.F90:385-386 */

  /* :2242-2255 */
  P_CaCO3->sflux_in = P_CaCO3->sflux_out;
  P_CaCO3->hflux_in = P_CaCO3->hflux_out;
  P_SiO2->sflux_in = P_SiO2->sflux_out;
  P_SiO2->hflux_in = P_SiO2->hflux_out;
  dust->sflux_in = dust->sflux_out;
  dust->hflux_in = dust->hflux_out;
  POC->sflux_in = POC->sflux_out;
  POC->hflux_in = POC->hflux_out;
  P_iron->sflux_in = P_iron->sflux_out;
  P_iron->hflux_in = P_iron->hflux_out;

  /* :2261-2267 */
  P_iron->sed_loss = c0;
  POC->sed_loss = c0;
  P_CaCO3->sed_loss = c0;
  P_SiO2->sed_loss = c0;
  dust->sed_loss = c0;
  *SED_DENITRIF = c0;
  *OTHER_REMIN = c0;

  /* :2273-2286 */
  if (cell_bottom_depth < p->parm_scalelen_z[0]) {
    scalelength = p->parm_scalelen_vals[0];
  } else if (cell_bottom_depth >= p->parm_scalelen_z[3]) {
    scalelength = p->parm_scalelen_vals[3];
  } else {
    for (n = 2; n <= 4; ++n) {
      if (cell_bottom_depth < p->parm_scalelen_z[n - 1]) {
        scalelength = p->parm_scalelen_vals[n - 2] +
                      (p->parm_scalelen_vals[n - 1] - p->parm_scalelen_vals[n - 2]) *
                          (cell_bottom_depth - p->parm_scalelen_z[n - 2]) /
                          (p->parm_scalelen_z[n - 1] - p->parm_scalelen_z[n - 2]);
        break;
      }
    }
  }

  DECAY_Hard = exp(-cell_thickness / 4.0e6);      /* :2288-2289 */
  DECAY_HardDust = exp(-cell_thickness / 1.2e7);

  TfuncS = pow(1.5, ((TEMP + p->T0_Kelvin_BGC) - (Tref + p->T0_Kelvin_BGC)) / c10);   /* :2295 */

  poc_error = 0;

  dz_loc = cell_thickness;
  dzr_loc = c1 / dz_loc;

  poc_diss = POC->diss;
  sio2_diss = P_SiO2->diss;
  caco3_diss = P_CaCO3->diss;
  dust_diss = dust->diss;

  /* :2311-2315 */
  if ((O2_loc >= 5.0) && (O2_loc < 40.0)) {
    poc_diss = POC->diss * (c1 + (3.3 - c1) * (40.0 - O2_loc) / 35.0);
  } else if (O2_loc < 5.0) {
    poc_diss = POC->diss * 3.3;
  }

  poc_diss = scalelength * poc_diss;     /* :2321-2324 */
  sio2_diss = scalelength * sio2_diss;
  caco3_diss = scalelength * caco3_diss;
  dust_diss = scalelength * dust_diss;

  sio2_diss = sio2_diss / TfuncS;        /* :2330 */

  decay_POC_E = exp(-dz_loc / poc_diss); /* :2336-2339 */
  decay_SiO2 = exp(-dz_loc / sio2_diss);
  decay_CaCO3 = exp(-dz_loc / caco3_diss);
  decay_dust = exp(-dz_loc / dust_diss);

  /* :2349-2365 */
  P_CaCO3->sflux_out = P_CaCO3->sflux_in * decay_CaCO3 +
                       P_CaCO3->prod * ((c1 - P_CaCO3->gamma) * (c1 - decay_CaCO3) * caco3_diss);
  P_CaCO3->hflux_out = P_CaCO3->hflux_in * DECAY_Hard + P_CaCO3->prod * (P_CaCO3->gamma * dz_loc);
  P_SiO2->sflux_out = P_SiO2->sflux_in * decay_SiO2 +
                      P_SiO2->prod * ((c1 - P_SiO2->gamma) * (c1 - decay_SiO2) * sio2_diss);
  P_SiO2->hflux_out = P_SiO2->hflux_in * DECAY_Hard + P_SiO2->prod * (P_SiO2->gamma * dz_loc);
  dust->sflux_out = dust->sflux_in * decay_dust;
  dust->hflux_out = dust->hflux_in * DECAY_HardDust;

  /* :2373-2383 */
  POC_PROD_avail = POC->prod - P_CaCO3->rho * P_CaCO3->prod - P_SiO2->rho * P_SiO2->prod;
  if (POC_PROD_avail < c0) {
    poc_error = 1;
  }

  /* :2390-2396 */
  if (*QA_dust_def > 0) {
    new_QA_dust_def = *QA_dust_def * (dust->sflux_out + dust->hflux_out) /
                      (dust->sflux_in + dust->hflux_in);
  } else {
    new_QA_dust_def = c0;
  }

  /* :2402-2412 */
  if (new_QA_dust_def > c0) {
    new_QA_dust_def = new_QA_dust_def - POC_PROD_avail * dz_loc;
    if (new_QA_dust_def < c0) {
      POC_PROD_avail = -new_QA_dust_def * dzr_loc;
      new_QA_dust_def = c0;
    } else {
      POC_PROD_avail = c0;
    }
  }
  *QA_dust_def = new_QA_dust_def;

  /* :2423-2438 */
  if (POC->hflux_in == c0 && POC->prod == c0) {
    POC->hflux_out = c0;
  } else {
    POC->hflux_out = P_CaCO3->rho * (P_CaCO3->sflux_out + P_CaCO3->hflux_out) +
                     P_SiO2->rho * (P_SiO2->sflux_out + P_SiO2->hflux_out) +
                     dust->rho * (dust->sflux_out + dust->hflux_out) - new_QA_dust_def;
    POC->hflux_out = fmax(POC->hflux_out, 0.0);
  }
  POC->sflux_out = POC->sflux_in * decay_POC_E + POC_PROD_avail * ((c1 - decay_POC_E) * poc_diss);

  /* :2445-2463 */
  P_CaCO3->remin = P_CaCO3->prod + ((P_CaCO3->sflux_in - P_CaCO3->sflux_out) +
                                    (P_CaCO3->hflux_in - P_CaCO3->hflux_out)) * dzr_loc;
  P_SiO2->remin = P_SiO2->prod + ((P_SiO2->sflux_in - P_SiO2->sflux_out) +
                                  (P_SiO2->hflux_in - P_SiO2->hflux_out)) * dzr_loc;
  POC->remin = POC->prod + ((POC->sflux_in - POC->sflux_out) +
                            (POC->hflux_in - POC->hflux_out)) * dzr_loc;
  dust->remin = ((dust->sflux_in - dust->sflux_out) + (dust->hflux_in - dust->hflux_out)) * dzr_loc;

  /* :2469-2486 */
  if (POC->sflux_in + POC->hflux_in == c0) {
    P_iron->remin = (POC->remin * parm_Red_Fe_C);
  } else {
    P_iron->remin = (POC->remin * (P_iron->sflux_in + P_iron->hflux_in) /
                     (POC->sflux_in + POC->hflux_in));
  }
  P_iron->remin = P_iron->remin + (P_iron->sflux_in * 1.5e-5);

  P_iron->sflux_out = P_iron->sflux_in + dz_loc * ((c1 - P_iron->gamma) * P_iron->prod - P_iron->remin);

  if (P_iron->sflux_out < c0) {
    P_iron->sflux_out = c0;
    P_iron->remin = P_iron->sflux_in * dzr_loc + (c1 - P_iron->gamma) * P_iron->prod;
  }

  /* :2497-2501 */
  P_iron->remin = P_iron->remin + dust->remin * dust_to_Fe + (FESEDFLUX_loc * dzr_loc);
  P_iron->hflux_out = P_iron->hflux_in;

  if (k == kmax) {   /* :2522-2631 */
    flux = POC->sflux_out + POC->hflux_out;

    if (flux > c0) {
      flux_alt = flux * mpercm * spd;

      POC->sed_loss = flux * fmin(0.8, p->parm_POMbury *
                                           (0.013 + 0.53 * flux_alt * flux_alt /
                                                        ((7.0 + flux_alt) * (7.0 + flux_alt))));

      *SED_DENITRIF = dzr_loc * flux * (0.06 + 0.19 * pow(0.99, (O2_loc - NO3_loc)));

      if (NO3_loc < 5.0) *SED_DENITRIF = 0.;

      flux_alt = flux * 1.0e-6 * spd * 365.0;
      *OTHER_REMIN = dzr_loc *
                     fmin(fmin(0.1 + flux_alt, 0.5) * (flux - POC->sed_loss),
                          (flux - POC->sed_loss - (*SED_DENITRIF * dz_loc * denitrif_C_N)));

      if (O2_loc < c1) {
        *OTHER_REMIN = dzr_loc * (flux - POC->sed_loss - (*SED_DENITRIF * dz_loc * denitrif_C_N));
      }
    }

    flux = P_SiO2->sflux_out + P_SiO2->hflux_out;
    flux_alt = flux * mpercm * spd;
    if (flux_alt > c2) {
      P_SiO2->sed_loss = 0.2;
    } else {
      P_SiO2->sed_loss = 0.04;
    }
    P_SiO2->sed_loss = flux * p->parm_BSIbury * P_SiO2->sed_loss;

    if (cell_bottom_depth < 3300.0e2) {
      flux = P_CaCO3->sflux_out + P_CaCO3->hflux_out;
      P_CaCO3->sed_loss = flux;
    }

    flux = P_CaCO3->sflux_out + P_CaCO3->hflux_out;
    if (flux > c0) {
      P_CaCO3->remin = P_CaCO3->remin + ((flux - P_CaCO3->sed_loss) * dzr_loc);
    }

    flux = P_SiO2->sflux_out + P_SiO2->hflux_out;
    if (flux > c0) {
      P_SiO2->remin = P_SiO2->remin + ((flux - P_SiO2->sed_loss) * dzr_loc);
    }

    flux = POC->sflux_out + POC->hflux_out;
    if (flux > c0) {
      POC->remin = POC->remin + ((flux - POC->sed_loss) * dzr_loc);
    }

    flux = (P_iron->sflux_out + P_iron->hflux_out);
    if (flux > c0) {
      P_iron->sed_loss = flux;
    }

    dust->sed_loss = dust->sflux_out + dust->hflux_out;

    P_CaCO3->sflux_out = c0;
    P_CaCO3->hflux_out = c0;
    P_SiO2->sflux_out = c0;
    P_SiO2->hflux_out = c0;
    dust->sflux_out = c0;
    dust->hflux_out = c0;
    POC->sflux_out = c0;
    POC->hflux_out = c0;
    P_iron->sflux_out = c0;
    P_iron->hflux_out = c0;
  }

  /* :2637-2694 */
  work = POC->sflux_in + POC->hflux_in;
  A2(d->diag_POC_FLUX_IN, k, column) = work;
  A2(d->diag_POC_PROD, k, column) = POC->prod;
  A2(d->diag_POC_REMIN, k, column) = POC->remin;

  work = P_CaCO3->sflux_in + P_CaCO3->hflux_in;
  A2(d->diag_CaCO3_FLUX_IN, k, column) = work;
  A2(d->diag_CaCO3_PROD, k, column) = P_CaCO3->prod;
  A2(d->diag_CaCO3_REMIN, k, column) = P_CaCO3->remin;

  work = P_SiO2->sflux_in + P_SiO2->hflux_in;
  A2(d->diag_SiO2_FLUX_IN, k, column) = work;
  A2(d->diag_SiO2_PROD, k, column) = P_SiO2->prod;
  A2(d->diag_SiO2_REMIN, k, column) = P_SiO2->remin;

  work = dust->sflux_in + dust->hflux_in;
  A2(d->diag_dust_FLUX_IN, k, column) = work;
  A2(d->diag_dust_REMIN, k, column) = dust->remin;

  work = P_iron->sflux_in + P_iron->hflux_in;
  A2(d->diag_P_iron_FLUX_IN, k, column) = work;
  A2(d->diag_P_iron_PROD, k, column) = P_iron->prod;
  A2(d->diag_P_iron_REMIN, k, column) = P_iron->remin;

  A2(d->diag_calcToSed, k, column) = P_CaCO3->sed_loss;
  A2(d->diag_bsiToSed, k, column) = P_SiO2->sed_loss;
  A2(d->diag_pocToSed, k, column) = POC->sed_loss;

  work = *SED_DENITRIF * cell_thickness;
  A2(d->diag_SedDenitrif, k, column) = work;
  work = *OTHER_REMIN * cell_thickness;
  A2(d->diag_OtherRemin, k, column) = work;
  work = (POC->sed_loss * Q);
  A2(d->diag_ponToSed, k, column) = work;
  work = (POC->sed_loss * Qp_zoo_pom);
  A2(d->diag_popToSed, k, column) = work;
  A2(d->diag_dustToSed, k, column) = dust->sed_loss;
  A2(d->diag_pfeToSed, k, column) = P_iron->sed_loss;

  if (poc_error && poc_error_count) *poc_error_count += 1;   /* reference: computed, never reported */
}

/* whole-array zero fills, BGC_mod.F90:570, :625-727 */
static void zero_fill(double *p, size_t n) {
  if (p) memset(p, 0, n * sizeof(double));
}

/* one column of column_loop, BGC_mod.F90:733-789 (setup for this column) + :799-1970 */
static void source_sink_column(const BgcParams *p, const BgcAutotroph *autotrophs,
                               const BgcIndices *ind, const BgcInput *in,
                               const BgcForcing *forcing, BgcOutput *out, BgcDiagnostics *d,
                               int nL, int nC, int column, int alt_co2_use_eco,
                               double *scratch, OracleSolverStats *st, long *poc_errors) {
  const int autotroph_cnt = 4;
  sinking_particle POC, P_CaCO3, P_SiO2, dust, P_iron;
  double QA_dust_def, dust_flux_in_loc, SED_DENITRIF, OTHER_REMIN, ZSATCALC = 0, ZSATARAG = 0,
         CO3_CALC_ANOM_km1 = 0, CO3_ARAG_ANOM_km1 = 0;
  double *DIC_loc, *DIC_ALT_CO2_loc, *ALK_loc, *PO4_loc, *NO3_loc, *SiO3_loc, *NH4_loc,
         *Fe_loc, *O2_loc, *DOC_loc, *zooC_loc, *DON_loc, *DOFe_loc, *DOP_loc, *DOPr_loc,
         *DONr_loc;
  double *autotrophChl_loc, *autotrophC_loc, *autotrophFe_loc, *autotrophSi_loc,
         *autotrophCaCO3_loc;
  int zero_mask;
  double work1, work2, work3, work4, work5, tmpTopt, tmpTmax;
  double f_loss_thres, ztop, PAR_out, PAR_in, KPARdz, PAR_avg, DOC_prod, DOC_remin,
         DON_remin, DOFe_remin, DOP_remin, NITRIF, DENITRIF, RESTORE;
  double z_umax, C_loss_thres;
  double Tfunc, f_nut, PCmax, light_lim, PCphoto, pChl;
  double f_zoo_detr, Fe_scavenge_rate, Fe_scavenge, Zprime, zoo_loss, zoo_loss_doc,
         zoo_loss_dic;
  double VNC, VPO4, VDOP, VPtot, VFe, VSiO3;
  double thetaC[4], QCaCO3[4], VNO3[4], VNH4[4], VNtot[4], NO3_V[4], NH4_V[4], PO4_V[4],
         DOP_V[4], Qfe[4], gQfe[4], Qsi[4], gQsi[4], Pprime[4], auto_graze[4],
         auto_graze_zoo[4], auto_graze_poc[4], auto_graze_doc[4], auto_graze_dic[4],
         auto_loss[4], auto_loss_poc[4], auto_loss_doc[4], auto_loss_dic[4], auto_agg[4],
         photoC[4], photoFe[4], photoSi[4], CaCO3_PROD[4], photoacc[4], Nfix[4],
         Nexcrete[4];
  double remaining_P;
  double remaining_P_dop[4], remaining_P_dip[4];
  double DON_prod, DOFe_prod, DOP_prod, O2_PRODUCTION, O2_CONSUMPTION, DONr_remin,
         DOPr_remin;
  double partial_thickness_100m, CO3, HCO3, H2CO3, CO3_ALT_CO2, HCO3_ALT_CO2,
         H2CO3_ALT_CO2;
  int k, n, auto_ind, auto_ind2, kmax;
  int i;

  const int po4_ind = ind->po4_ind, no3_ind = ind->no3_ind, sio3_ind = ind->sio3_ind,
            nh4_ind = ind->nh4_ind, fe_ind = ind->fe_ind, o2_ind = ind->o2_ind,
            dic_ind = ind->dic_ind, dic_alt_co2_ind = ind->dic_alt_co2_ind,
            alk_ind = ind->alk_ind, doc_ind = ind->doc_ind, don_ind = ind->don_ind,
            dofe_ind = ind->dofe_ind, dop_ind = ind->dop_ind, dopr_ind = ind->dopr_ind,
            donr_ind = ind->donr_ind, zooC_ind = ind->zooC_ind;

  const double epsC = p->epsC, epsTinv = p->epsTinv, cks = p->cks, cksi = p->cksi,
               dust_fescav_scale = p->dust_fescav_scale;
  const double T0_Kelvin_BGC = p->T0_Kelvin_BGC;

  const double *tr = in->BGC_tracers;
  double *tend = out->BGC_tendencies;

  /* per-column scratch standing in for the reference's allocated (k,col[,auto]) locals;
   * indexed [k-1] and [(k-1) + nL*(auto-1)] */
  DIC_loc = scratch + 0 * nL; DIC_ALT_CO2_loc = scratch + 1 * nL; ALK_loc = scratch + 2 * nL;
  PO4_loc = scratch + 3 * nL; NO3_loc = scratch + 4 * nL; SiO3_loc = scratch + 5 * nL;
  NH4_loc = scratch + 6 * nL; Fe_loc = scratch + 7 * nL; O2_loc = scratch + 8 * nL;
  DOC_loc = scratch + 9 * nL; zooC_loc = scratch + 10 * nL; DON_loc = scratch + 11 * nL;
  DOFe_loc = scratch + 12 * nL; DOP_loc = scratch + 13 * nL; DOPr_loc = scratch + 14 * nL;
  DONr_loc = scratch + 15 * nL;
  autotrophChl_loc = scratch + 16 * nL; autotrophC_loc = scratch + 20 * nL;
  autotrophFe_loc = scratch + 24 * nL; autotrophSi_loc = scratch + 28 * nL;
  autotrophCaCO3_loc = scratch + 32 * nL;
#define L1(p_, k_) ((p_)[(k_)-1])
#define LA(p_, k_, a_) ((p_)[((k_)-1) + nL * ((a_)-1)])

  for (i = 0; i < 4; ++i) {   /* guards in the reference make these never read undefined */
    QCaCO3[i] = 0; Qsi[i] = 0; gQsi[i] = 0; photoSi[i] = 0; CaCO3_PROD[i] = 0; Nfix[i] = 0;
    Nexcrete[i] = 0; remaining_P_dop[i] = 0; remaining_P_dip[i] = 0;
  }
  memset(&POC, 0, sizeof POC); memset(&P_CaCO3, 0, sizeof P_CaCO3);
  memset(&P_SiO2, 0, sizeof P_SiO2); memset(&dust, 0, sizeof dust);
  memset(&P_iron, 0, sizeof P_iron);

  kmax = in->number_of_active_levels[column - 1];
  if (kmax < 1) return;

  /* ---- setup_loop body, :738-787 */
  for (k = 1; k <= kmax; ++k) {
    L1(DIC_loc, k) = fmax(0.0, A3(tr, k, column, dic_ind));
    L1(DIC_ALT_CO2_loc, k) = fmax(0.0, A3(tr, k, column, dic_alt_co2_ind));
    L1(ALK_loc, k) = fmax(0.0, A3(tr, k, column, alk_ind));
    L1(PO4_loc, k) = fmax(0.0, A3(tr, k, column, po4_ind));
    L1(NO3_loc, k) = fmax(0.0, A3(tr, k, column, no3_ind));
    L1(SiO3_loc, k) = fmax(0.0, A3(tr, k, column, sio3_ind));
    L1(NH4_loc, k) = fmax(0.0, A3(tr, k, column, nh4_ind));
    L1(Fe_loc, k) = fmax(0.0, A3(tr, k, column, fe_ind));
    L1(O2_loc, k) = fmax(0.0, A3(tr, k, column, o2_ind));
    L1(DOC_loc, k) = fmax(0.0, A3(tr, k, column, doc_ind));
    L1(zooC_loc, k) = fmax(0.0, A3(tr, k, column, zooC_ind));
    L1(DON_loc, k) = fmax(0.0, A3(tr, k, column, don_ind));
    L1(DOFe_loc, k) = fmax(0.0, A3(tr, k, column, dofe_ind));
    L1(DOP_loc, k) = fmax(0.0, A3(tr, k, column, dop_ind));
    L1(DOPr_loc, k) = fmax(0.0, A3(tr, k, column, dopr_ind));
    L1(DONr_loc, k) = fmax(0.0, A3(tr, k, column, donr_ind));

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      n = at->Chl_ind;
      LA(autotrophChl_loc, k, auto_ind) = fmax(0.0, A3(tr, k, column, n));
      n = at->C_ind;
      LA(autotrophC_loc, k, auto_ind) = fmax(0.0, A3(tr, k, column, n));
      n = at->Fe_ind;
      LA(autotrophFe_loc, k, auto_ind) = fmax(0.0, A3(tr, k, column, n));
      n = at->Si_ind;
      if (n > 0) LA(autotrophSi_loc, k, auto_ind) = fmax(0.0, A3(tr, k, column, n));
      n = at->CaCO3_ind;
      if (n > 0) LA(autotrophCaCO3_loc, k, auto_ind) = fmax(0.0, A3(tr, k, column, n));
    }
  }
  (void)DIC_ALT_CO2_loc;   /* Q1: filled and never read in BGC_SourceSink */

  /* ---- column_loop body, :808-814 */
  dust_flux_in_loc = fmax(0.0, C1(forcing->dust_FLUX_IN, column));

  init_particulate_terms(p, &POC, &P_CaCO3, &P_SiO2, &dust, &P_iron, &QA_dust_def,
                         dust_flux_in_loc);

  PAR_out = fmax(0.0, C1(forcing->ShortWaveFlux_surface, column));
  PAR_out = PAR_out * f_qsw_par;

  for (k = 1; k <= kmax; ++k) {   /* :820 */
    const double TEMP = A2(in->PotentialTemperature, k, column);
    const double SALT = A2(in->Salinity, k, column);
    const double zmid = A2(in->cell_center_depth, k, column);
    const double dz = A2(in->cell_thickness, k, column);
    const double zbot = A2(in->cell_bottom_depth, k, column);
    const double lat = C1(in->cell_latitude, column);

    /* :826-844 */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      zero_mask = LA(autotrophChl_loc, k, auto_ind) == c0 ||
                  LA(autotrophC_loc, k, auto_ind) == c0 ||
                  LA(autotrophFe_loc, k, auto_ind) == c0;
      if (at->Si_ind > 0) zero_mask = zero_mask || LA(autotrophSi_loc, k, auto_ind) == c0;
      if (zero_mask) {
        LA(autotrophChl_loc, k, auto_ind) = c0;
        LA(autotrophC_loc, k, auto_ind) = c0;
        LA(autotrophFe_loc, k, auto_ind) = c0;
      }
      if (at->Si_ind > 0) {
        if (zero_mask) LA(autotrophSi_loc, k, auto_ind) = c0;
      }
      if (at->CaCO3_ind > 0) {
        if (zero_mask) LA(autotrophCaCO3_loc, k, auto_ind) = c0;
      }
    }

    /* :850-856 */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      thetaC[auto_ind - 1] = LA(autotrophChl_loc, k, auto_ind) / (LA(autotrophC_loc, k, auto_ind) + epsC);
      Qfe[auto_ind - 1] = LA(autotrophFe_loc, k, auto_ind) / (LA(autotrophC_loc, k, auto_ind) + epsC);
      if (at->Si_ind > 0) {
        Qsi[auto_ind - 1] = fmin(LA(autotrophSi_loc, k, auto_ind) / (LA(autotrophC_loc, k, auto_ind) + epsC), gQsi_max);
      }
    }

    /* :864-898 */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      gQfe[auto_ind - 1] = at->gQfe_0;
      if (L1(Fe_loc, k) < cks * at->kFe) {
        gQfe[auto_ind - 1] = fmax((gQfe[auto_ind - 1] * L1(Fe_loc, k) / (cks * at->kFe)), at->gQfe_min);
      }

      if (at->Si_ind > 0) {
        gQsi[auto_ind - 1] = gQsi_0;
        if ((L1(Fe_loc, k) < cksi * at->kFe) && (L1(Fe_loc, k) > c0) &&
            (L1(SiO3_loc, k) > (cksi * at->kSiO3))) {
          gQsi[auto_ind - 1] = fmin((gQsi[auto_ind - 1] * cksi * at->kFe / L1(Fe_loc, k)), gQsi_max);
        }
        if (L1(Fe_loc, k) == c0) {
          gQsi[auto_ind - 1] = gQsi_max;
        }
        if (L1(SiO3_loc, k) < (cksi * at->kSiO3)) {
          gQsi[auto_ind - 1] = fmax((gQsi[auto_ind - 1] * L1(SiO3_loc, k) / (cksi * at->kSiO3)), gQsi_min);
        }
      }

      if (at->CaCO3_ind > 0) {
        QCaCO3[auto_ind - 1] = LA(autotrophCaCO3_loc, k, auto_ind) / (LA(autotrophC_loc, k, auto_ind) + epsC);
        if (QCaCO3[auto_ind - 1] > QCaCO3_max) QCaCO3[auto_ind - 1] = QCaCO3_max;
      }
    }

    /* :907-924 */
    PAR_in = PAR_out;

    {
      double s = 0.0;
      for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) s = s + LA(autotrophChl_loc, k, auto_ind);
      work1 = fmax(s, 0.02);
    }
    if (work1 < 0.13224) {
      KPARdz = 0.000919 * pow(work1, 0.3536);
    } else {
      KPARdz = 0.001131 * pow(work1, 0.4562);
    }
    KPARdz = KPARdz * dz;

    PAR_out = PAR_in * exp(-KPARdz);
    PAR_avg = PAR_in * (c1 - exp(-KPARdz)) / KPARdz;

    /* :940-984  (lcalc_co2_terms = lalt_co2_terms = .true.) */
    work3 = c0;
    work4 = c0;
    if (A2(out->PH_PREV_3D, k, column) != c0) {
      work1 = A2(out->PH_PREV_3D, k, column) - del_ph;
      work2 = A2(out->PH_PREV_3D, k, column) + del_ph;
    } else {
      work1 = phlo_3d_init;
      work2 = phhi_3d_init;
    }
    work5 = zmid * 0.01;
    oracle_comp_CO3terms(k, work5, 1, TEMP, SALT, L1(DIC_loc, k), L1(ALK_loc, k),
                         L1(PO4_loc, k), L1(SiO3_loc, k), &work1, &work2, &work3, &H2CO3,
                         &HCO3, &CO3, st);
    A2(out->PH_PREV_3D, k, column) = work3;

    if (A2(out->PH_PREV_ALT_CO2_3D, k, column) != c0) {
      work1 = A2(out->PH_PREV_ALT_CO2_3D, k, column) - del_ph;
      work2 = A2(out->PH_PREV_ALT_CO2_3D, k, column) + del_ph;
    } else {
      work1 = phlo_3d_init;
      work2 = phhi_3d_init;
    }
    work5 = zmid * 0.01;
    /* Q1: DIC_loc, not DIC_ALT_CO2_loc (:975) */
    oracle_comp_CO3terms(k, work5, 1, TEMP, SALT, L1(DIC_loc, k), L1(ALK_loc, k),
                         L1(PO4_loc, k), L1(SiO3_loc, k), &work1, &work2, &work4,
                         &H2CO3_ALT_CO2, &HCO3_ALT_CO2, &CO3_ALT_CO2, st);
    A2(out->PH_PREV_ALT_CO2_3D, k, column) = work4;

    /* :986-1001 */
    A2(d->diag_CO3, k, column) = CO3;
    A2(d->diag_HCO3, k, column) = HCO3;
    A2(d->diag_H2CO3, k, column) = H2CO3;
    A2(d->diag_pH_3D, k, column) = work3;
    A2(d->diag_CO3_ALT_CO2, k, column) = CO3_ALT_CO2;
    A2(d->diag_HCO3_ALT_CO2, k, column) = HCO3_ALT_CO2;
    A2(d->diag_H2CO3_ALT_CO2, k, column) = H2CO3_ALT_CO2;
    A2(d->diag_pH_3D_ALT_CO2, k, column) = work4;

    work5 = zmid * 0.01;
    oracle_comp_co3_sat_vals(k, work5, TEMP, SALT, &work1, &work2);

    A2(d->diag_co3_sat_calc, k, column) = work1;
    A2(d->diag_co3_sat_arag, k, column) = work2;

    /* :1003-1032 */
    if (k == 1) {
      ZSATCALC = (CO3 > work1) ? -c1 : c0;
      ZSATARAG = (CO3 > work2) ? -c1 : c0;
    } else {
      work4 = A2(in->cell_center_depth, k - 1, column) +
              (A2(in->cell_center_depth, k, column) - A2(in->cell_center_depth, k - 1, column));
      if (ZSATCALC == -c1 && CO3 <= work1) {
        ZSATCALC = work4 * CO3_CALC_ANOM_km1 / (CO3_CALC_ANOM_km1 - (CO3 - work1));
      }
      if (ZSATARAG == -c1 && CO3 <= work2) {
        ZSATARAG = work4 * CO3_ARAG_ANOM_km1 / (CO3_ARAG_ANOM_km1 - (CO3 - work2));
      }
      if (ZSATCALC == -c1 && k == kmax) {
        ZSATCALC = zbot;
      }
      if (ZSATARAG == -c1 && k == kmax) {
        ZSATARAG = zbot;
      }
    }

    CO3_CALC_ANOM_km1 = CO3 - work1;
    CO3_ARAG_ANOM_km1 = CO3 - work2;

    if (k == kmax) {
      C1(d->diag_zsatcalc, column) = ZSATCALC;
      C1(d->diag_zsatarag, column) = ZSATARAG;
    }

    /* :1041 */
    Tfunc = pow(Q_10, ((TEMP + T0_Kelvin_BGC) - (Tref + T0_Kelvin_BGC)) / c10);

    /* :1047-1055 */
    if (zmid > thres_z1) {
      if (zmid < thres_z2) {
        f_loss_thres = (thres_z2 - zmid) / (thres_z2 - thres_z1);
      } else {
        f_loss_thres = c0;
      }
    } else {
      f_loss_thres = c1;
    }

    /* :1072-1094 */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      C_loss_thres = f_loss_thres * at->loss_thres;
      switch (at->temp_function) {
        case BGC_TFNC_Q10:
          if (TEMP < at->temp_thres) C_loss_thres = f_loss_thres * at->loss_thres2;
          break;
        case BGC_TFNC_QUASI_MMRT:
          if (lat >= 0.0) {
            tmpTmax = at->temp_thresN;
          } else {
            tmpTmax = at->temp_thresS;
          }
          if (TEMP > tmpTmax) C_loss_thres = f_loss_thres * at->loss_thres2;
          break;
        default:
          break;
      }
      Pprime[auto_ind - 1] = fmax(LA(autotrophC_loc, k, auto_ind) - C_loss_thres, 0.0);
    }

    /* :1107-1388 */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      const int a = auto_ind - 1;

      VNO3[a] = (L1(NO3_loc, k) / at->kNO3) /
                (c1 + (L1(NO3_loc, k) / at->kNO3) + (L1(NH4_loc, k) / at->kNH4));
      VNH4[a] = (L1(NH4_loc, k) / at->kNH4) /
                (c1 + (L1(NO3_loc, k) / at->kNO3) + (L1(NH4_loc, k) / at->kNH4));
      VNtot[a] = VNO3[a] + VNH4[a];
      if (at->Nfixer) VNtot[a] = c1;
      A3(d->diag_N_lim, k, column, auto_ind) = VNtot[a];

      VFe = L1(Fe_loc, k) / (L1(Fe_loc, k) + at->kFe);
      A3(d->diag_Fe_lim, k, column, auto_ind) = VFe;

      f_nut = fmin(VNtot[a], VFe);

      VPO4 = (L1(PO4_loc, k) / at->kPO4) /
             (c1 + (L1(PO4_loc, k) / at->kPO4) + (L1(DOP_loc, k) / at->kDOP));
      VDOP = (L1(DOP_loc, k) / at->kDOP) /
             (c1 + (L1(PO4_loc, k) / at->kPO4) + (L1(DOP_loc, k) / at->kDOP));
      VPtot = VPO4 + VDOP;

      A3(d->diag_P_lim, k, column, auto_ind) = VPtot;

      f_nut = fmin(f_nut, VPtot);

      if (at->kSiO3 > c0) {
        VSiO3 = L1(SiO3_loc, k) / (L1(SiO3_loc, k) + at->kSiO3);
        A3(d->diag_SiO3_lim, k, column, auto_ind) = VSiO3;
        f_nut = fmin(f_nut, VSiO3);
      }

      /* :1146-1173 */
      PCmax = at->PCref * f_nut * Tfunc;
      if (TEMP < at->temp_thres) PCmax = c0;

      switch (at->temp_function) {
        case BGC_TFNC_Q10:
          PCmax = PCmax;
          break;
        case BGC_TFNC_QUASI_MMRT:
          if (lat >= 0.0) {
            tmpTopt = at->temp_optN;
            tmpTmax = at->temp_thresN;
          } else {
            tmpTopt = at->temp_optS;
            tmpTmax = at->temp_thresS;
          }
          PCmax = PCmax * fmin(1.0, ((tmpTmax - TEMP) / (tmpTmax - tmpTopt)));
          if (TEMP > tmpTmax) PCmax = c0;
          break;
        default:
          break;
      }

      /* :1175-1181 */
      light_lim = (c1 - exp((-c1 * at->alphaPI * thetaC[a] * PAR_avg) / (PCmax + epsTinv)));
      PCphoto = PCmax * light_lim;

      A3(d->diag_light_lim, k, column, auto_ind) = light_lim;

      photoC[a] = PCphoto * LA(autotrophC_loc, k, auto_ind);

      /* :1193-1221 */
      if (VNtot[a] > c0) {
        NO3_V[a] = (VNO3[a] / VNtot[a]) * photoC[a] * Q;
        NH4_V[a] = (VNH4[a] / VNtot[a]) * photoC[a] * Q;
        VNC = PCphoto * Q;
      } else {
        NO3_V[a] = c0;
        NH4_V[a] = c0;
        VNC = c0;
      }
      A3(d->diag_photoNO3, k, column, auto_ind) = NO3_V[a];
      A3(d->diag_photoNH4, k, column, auto_ind) = NH4_V[a];

      if (VPtot > c0) {
        PO4_V[a] = (VPO4 / VPtot) * photoC[a] * at->Qp;
        DOP_V[a] = (VDOP / VPtot) * photoC[a] * at->Qp;
      } else {
        PO4_V[a] = c0;
        DOP_V[a] = c0;
      }

      A3(d->diag_PO4_uptake, k, column, auto_ind) = PO4_V[a];
      A3(d->diag_DOP_uptake, k, column, auto_ind) = DOP_V[a];

      photoFe[a] = photoC[a] * gQfe[a];
      A3(d->diag_photoFe, k, column, auto_ind) = photoFe[a];

      /* :1227-1232 */
      if (at->Si_ind > 0) {
        photoSi[a] = photoC[a] * gQsi[a];
        A3(d->diag_bSi_form, k, column, auto_ind) = photoSi[a];
        C1(d->diag_tot_bSi_form, column) = C1(d->diag_tot_bSi_form, column) + photoSi[a];
      }

      /* :1240-1246 */
      work1 = at->alphaPI * thetaC[a] * PAR_avg;
      if (work1 > c0) {
        pChl = at->thetaN_max * PCphoto / work1;
        photoacc[a] = (pChl * VNC / thetaC[a]) * LA(autotrophChl_loc, k, auto_ind);
      } else {
        photoacc[a] = c0;
      }

      /* :1255-1278 */
      if (at->imp_calcifier) {
        CaCO3_PROD[a] = p->parm_f_prod_sp_CaCO3 * photoC[a];
        CaCO3_PROD[a] = CaCO3_PROD[a] * f_nut;

        if (TEMP < CaCO3_temp_thres1)
          CaCO3_PROD[a] = CaCO3_PROD[a] * fmax((TEMP - CaCO3_temp_thres2), 0.0) /
                          (CaCO3_temp_thres1 - CaCO3_temp_thres2);

        if (LA(autotrophC_loc, k, auto_ind) > CaCO3_sp_thres)
          CaCO3_PROD[a] = fmin((CaCO3_PROD[a] * LA(autotrophC_loc, k, auto_ind) / CaCO3_sp_thres),
                               (f_photosp_CaCO3 * photoC[a]));

        A3(d->diag_CaCO3_form, k, column, auto_ind) = CaCO3_PROD[a];
        A2(d->diag_tot_CaCO3_form, k, column) = A2(d->diag_tot_CaCO3_form, k, column) + CaCO3_PROD[a];

        work1 = dz * CaCO3_PROD[a];
        F2(d->diag_CaCO3_form_zint, column, auto_ind) = F2(d->diag_CaCO3_form_zint, column, auto_ind) + work1;
        C1(d->diag_tot_CaCO3_form_zint, column) = C1(d->diag_tot_CaCO3_form_zint, column) + work1;
      }

      /* :1285-1290 */
      auto_loss[a] = at->mort * Pprime[a] * Tfunc;

      auto_agg[a] = fmin((at->agg_rate_max * dps) * Pprime[a], at->mort2 * Pprime[a] * Pprime[a]);
      auto_agg[a] = fmax((at->agg_rate_min * dps) * Pprime[a], auto_agg[a]);

      /* :1297-1324 */
      work1 = c0;
      for (auto_ind2 = 1; auto_ind2 <= autotroph_cnt; ++auto_ind2) {
        if (autotrophs[auto_ind2 - 1].grazee_ind == at->grazee_ind)
          work1 = work1 + Pprime[auto_ind2 - 1];
      }

      z_umax = at->z_umax_0 * Tfunc;

      if (auto_ind == ind->diat_ind) {
        if ((lat >= 0.0) && (TEMP > at->temp_optN)) {
          z_umax = z_umax * fmax((at->temp_thresN - TEMP) / (at->temp_thresN - at->temp_optN), 0.95);
        } else if ((lat <= 0.0) && (TEMP > at->temp_optS)) {
          z_umax = z_umax * fmax((at->temp_thresS - TEMP) / (at->temp_thresS - at->temp_optS), 0.95);
        }
      }

      if (work1 > c0) {
        auto_graze[a] = (Pprime[a] / work1) * z_umax * L1(zooC_loc, k) * (work1 / (work1 + at->z_grz));
      } else {
        auto_graze[a] = c0;
      }

      /* :1331-1338 */
      if (at->Nfixer) {
        work1 = photoC[a] * Q;
        Nfix[a] = (work1 * r_Nfix_photo) - NO3_V[a] - NH4_V[a];
        Nexcrete[a] = Nfix[a] + NO3_V[a] + NH4_V[a] - work1;
        A3(d->diag_Nfix, k, column, auto_ind) = Nfix[a];
        A2(d->diag_tot_Nfix, k, column) = A2(d->diag_tot_Nfix, k, column) + Nfix[a];
      }

      /* :1354-1372 */
      auto_graze_zoo[a] = at->graze_zoo * auto_graze[a];
      if (at->imp_calcifier) {
        auto_graze_poc[a] = auto_graze[a] * fmax((caco3_poc_min * QCaCO3[a]),
                                                 fmin(spc_poc_fac * fmax(1.0, Pprime[a]),
                                                      f_graze_sp_poc_lim));
      } else {
        auto_graze_poc[a] = at->graze_poc * auto_graze[a];
      }
      auto_graze_doc[a] = at->graze_doc * auto_graze[a];
      auto_graze_dic[a] = auto_graze[a] - (auto_graze_zoo[a] + auto_graze_poc[a] + auto_graze_doc[a]);

      if (at->imp_calcifier) {
        auto_loss_poc[a] = QCaCO3[a] * auto_loss[a];
      } else {
        auto_loss_poc[a] = at->loss_poc * auto_loss[a];
      }
      auto_loss_doc[a] = (c1 - p->parm_labile_ratio) * (auto_loss[a] - auto_loss_poc[a]);
      auto_loss_dic[a] = p->parm_labile_ratio * (auto_loss[a] - auto_loss_poc[a]);

      /* :1380-1386 */
      if (at->Qp != Qp_zoo_pom) {
        remaining_P = ((auto_graze[a] + auto_loss[a] + auto_agg[a]) * at->Qp) -
                      ((auto_graze_zoo[a]) * Qp_zoo_pom) -
                      ((auto_graze_poc[a] + auto_loss_poc[a] + auto_agg[a]) * Qp_zoo_pom);
        remaining_P_dop[a] = (c1 - p->parm_labile_ratio) * remaining_P;
        remaining_P_dip[a] = p->parm_labile_ratio * remaining_P;
      }
    }

    /* :1395-1401 */
    work1 = c0;
    work2 = c0;
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      work1 = work1 + autotrophs[auto_ind - 1].f_zoo_detr * (auto_graze[auto_ind - 1] + epsC * epsTinv);
      work2 = work2 + (auto_graze[auto_ind - 1] + epsC * epsTinv);
    }
    f_zoo_detr = work1 / work2;

    /* :1408-1415 */
    C_loss_thres = f_loss_thres * loss_thres_zoo;
    Zprime = fmax(L1(zooC_loc, k) - C_loss_thres, 0.0);
    zoo_loss = (p->parm_z_mort2_0 * pow(Zprime, 1.5) + p->parm_z_mort_0 * Zprime) * Tfunc;
    zoo_loss_doc = (c1 - p->parm_labile_ratio) * (c1 - f_zoo_detr) * zoo_loss;
    zoo_loss_dic = p->parm_labile_ratio * (c1 - f_zoo_detr) * zoo_loss;

    /* :1421-1439 */
    DOC_prod = zoo_loss_doc + sum4(auto_loss_doc) + sum4(auto_graze_doc);
    DON_prod = Q * DOC_prod;
    DOP_prod = Qp_zoo_pom * zoo_loss_doc;
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      if (at->Qp == Qp_zoo_pom) {
        DOP_prod = DOP_prod + at->Qp * (auto_loss_doc[auto_ind - 1] + auto_graze_doc[auto_ind - 1]);
      } else {
        DOP_prod = DOP_prod + remaining_P_dop[auto_ind - 1];
      }
    }
    DOFe_prod = Qfe_zoo * zoo_loss_doc;
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      DOFe_prod = DOFe_prod + Qfe[auto_ind - 1] * (auto_loss_doc[auto_ind - 1] + auto_graze_doc[auto_ind - 1]);
    }

    DOC_remin = L1(DOC_loc, k) * DOC_reminR;
    DON_remin = L1(DON_loc, k) * DON_reminR;
    DOFe_remin = L1(DOFe_loc, k) * DOFe_reminR;
    DOP_remin = L1(DOP_loc, k) * DOP_reminR;

    /* :1451-1461 */
    if (PAR_avg > 1.0) {
      DONr_remin = L1(DONr_loc, k) * DONr_reminR;
      DOPr_remin = L1(DOPr_loc, k) * DOPr_reminR;
    } else {
      DONr_remin = L1(DONr_loc, k) * (c1 / (365.0 * 670.0)) * dps;
      DOPr_remin = L1(DOPr_loc, k) * (c1 / (365.0 * 460.0)) * dps;
      DOC_remin = DOC_remin * 0.0685;
      DON_remin = DON_remin * 0.1;
      DOFe_remin = DOFe_remin * 0.05;
      DOP_remin = DOP_remin * 0.05;
    }

    /* :1467-1468 */
    POC.prod = f_zoo_detr * zoo_loss + sum4(auto_graze_poc) + sum4(auto_agg) + sum4(auto_loss_poc);

    /* :1480-1498  (Q10: last writer wins) */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      if (autotrophs[auto_ind - 1].CaCO3_ind > 0) {
        P_CaCO3.prod = ((c1 - f_graze_CaCO3_remin) * auto_graze[auto_ind - 1] +
                        auto_loss[auto_ind - 1] + auto_agg[auto_ind - 1]) * QCaCO3[auto_ind - 1];
      }
    }
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      if (at->Si_ind > 0) {
        P_SiO2.prod = Qsi[auto_ind - 1] *
                      ((c1 - f_graze_si_remin) * auto_graze[auto_ind - 1] + auto_agg[auto_ind - 1] +
                       at->loss_poc * auto_loss[auto_ind - 1]);
      }
    }

    dust.prod = c0;

    /* :1510-1529 */
    Fe_scavenge_rate = p->parm_fe_scavenge_rate0;

    Fe_scavenge_rate = Fe_scavenge_rate *
                       ((POC.sflux_out + POC.hflux_out) * 120.1 +
                        (P_CaCO3.sflux_out + P_CaCO3.hflux_out) * P_CaCO3.mass +
                        (P_SiO2.sflux_out + P_SiO2.hflux_out) * P_SiO2.mass +
                        (dust.sflux_out + dust.hflux_out) * dust_fescav_scale);

    if (L1(Fe_loc, k) > fe_scavenge_thres1)
      Fe_scavenge_rate = Fe_scavenge_rate + (L1(Fe_loc, k) - fe_scavenge_thres1) * fe_max_scale2;

    Fe_scavenge = yps * L1(Fe_loc, k) * Fe_scavenge_rate;

    P_iron.prod = (zoo_loss * f_zoo_detr * Qfe_zoo) + Fe_scavenge;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      P_iron.prod = P_iron.prod +
                    Qfe[auto_ind - 1] * (auto_agg[auto_ind - 1] + auto_graze_poc[auto_ind - 1] +
                                         auto_loss_poc[auto_ind - 1]);
    }

    /* :1531-1536 */
    compute_particulate_terms(p, column, k, kmax, &POC, &P_CaCO3, &P_SiO2, &dust, &P_iron,
                              &QA_dust_def, TEMP, L1(O2_loc, k), L1(NO3_loc, k), &SED_DENITRIF,
                              &OTHER_REMIN, dz, zbot, A2(forcing->FESEDFLUX, k, column), d, nL,
                              poc_errors);

    /* :1545-1563 */
    if (p->lrest_no3) {
      RESTORE = A2(forcing->NUTR_RESTORE_RTAU, k, column) *
                (A2(forcing->NO3_CLIM, k, column) - L1(NO3_loc, k));
    } else {
      RESTORE = c0;
    }
    A2(d->diag_NO3_RESTORE, k, column) = RESTORE;

    if (PAR_out < p->parm_nitrif_par_lim) {
      NITRIF = p->parm_kappa_nitrif * L1(NH4_loc, k);
      if (PAR_in > p->parm_nitrif_par_lim) {
        NITRIF = NITRIF * log(PAR_out / p->parm_nitrif_par_lim) / (-KPARdz);
      }
    } else {
      NITRIF = c0;
    }
    A2(d->diag_NITRIF, k, column) = NITRIF;

    /* :1569-1577 */
    work1 = ((p->parm_o2_min + p->parm_o2_min_delta) - L1(O2_loc, k)) / p->parm_o2_min_delta;
    work1 = fmin(fmax(work1, 0.0), 1.0);
    work1 = (L1(NO3_loc, k) == c0) ? 0.0 : work1;

    DENITRIF = work1 * ((DOC_remin + POC.remin - OTHER_REMIN) / denitrif_C_N - SED_DENITRIF);
    A2(d->diag_DENITRIF, k, column) = DENITRIF;

    /* :1583-1592 */
    A3(tend, k, column, no3_ind) = RESTORE + NITRIF - DENITRIF - SED_DENITRIF - sum4(NO3_V);

    A3(tend, k, column, nh4_ind) = -sum4(NH4_V) - NITRIF + DON_remin + DONr_remin +
                                   Q * (zoo_loss_dic + sum4(auto_loss_dic) + sum4(auto_graze_dic) +
                                        POC.remin * (c1 - DONrefract));

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      if (autotrophs[auto_ind - 1].Nfixer)
        A3(tend, k, column, nh4_ind) = A3(tend, k, column, nh4_ind) + Nexcrete[auto_ind - 1];
    }

    /* :1598-1605 */
    A3(tend, k, column, fe_ind) = P_iron.remin + (Qfe_zoo * zoo_loss_dic) + DOFe_remin -
                                  sum4(photoFe) - Fe_scavenge;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      A3(tend, k, column, fe_ind) = A3(tend, k, column, fe_ind) +
                                    (Qfe[auto_ind - 1] * (auto_loss_dic[auto_ind - 1] + auto_graze_dic[auto_ind - 1])) +
                                    auto_graze_zoo[auto_ind - 1] * (Qfe[auto_ind - 1] - Qfe_zoo);
    }

    /* :1611-1628 */
    if (p->lrest_sio3) {
      RESTORE = A2(forcing->NUTR_RESTORE_RTAU, k, column) *
                (A2(forcing->SiO3_CLIM, k, column) - L1(SiO3_loc, k));
    } else {
      RESTORE = c0;
    }
    A2(d->diag_SiO3_RESTORE, k, column) = RESTORE;

    A3(tend, k, column, sio3_ind) = RESTORE + P_SiO2.remin;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      if (at->Si_ind > 0) {
        A3(tend, k, column, sio3_ind) = A3(tend, k, column, sio3_ind) - photoSi[auto_ind - 1] +
                                        Qsi[auto_ind - 1] * (f_graze_si_remin * auto_graze[auto_ind - 1] +
                                                             (c1 - at->loss_poc) * auto_loss[auto_ind - 1]);
      }
    }

    /* :1634-1661 */
    if (p->lrest_po4) {
      RESTORE = A2(forcing->NUTR_RESTORE_RTAU, k, column) *
                (A2(forcing->PO4_CLIM, k, column) - L1(PO4_loc, k));
    } else {
      RESTORE = c0;
    }
    A2(d->diag_PO4_RESTORE, k, column) = RESTORE;

    A3(tend, k, column, po4_ind) = RESTORE + DOP_remin + DOPr_remin - sum4(PO4_V) +
                                   Qp_zoo_pom * ((c1 - DOPrefract) * POC.remin + zoo_loss_dic);

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      if (at->Qp == Qp_zoo_pom) {
        A3(tend, k, column, po4_ind) = A3(tend, k, column, po4_ind) +
                                       at->Qp * (auto_loss_dic[auto_ind - 1] + auto_graze_dic[auto_ind - 1]);
      } else {
        A3(tend, k, column, po4_ind) = A3(tend, k, column, po4_ind) + remaining_P_dip[auto_ind - 1];
      }
    }

    /* :1676-1697 */
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const BgcAutotroph *at = &autotrophs[auto_ind - 1];
      const int a = auto_ind - 1;
      work1 = auto_graze[a] + auto_loss[a] + auto_agg[a];

      n = at->C_ind;
      A3(tend, k, column, n) = photoC[a] - work1;
      n = at->Chl_ind;
      A3(tend, k, column, n) = photoacc[a] - thetaC[a] * work1;
      n = at->Fe_ind;
      A3(tend, k, column, n) = photoFe[a] - Qfe[a] * work1;
      n = at->Si_ind;
      if (n > 0) {
        A3(tend, k, column, n) = photoSi[a] - Qsi[a] * work1;
      }
      n = at->CaCO3_ind;
      if (n > 0) {
        A3(tend, k, column, n) = CaCO3_PROD[a] - QCaCO3[a] * work1;
      }
    }

    /* :1703 */
    A3(tend, k, column, zooC_ind) = sum4(auto_graze_zoo) - zoo_loss;

    /* :1710-1723 */
    A3(tend, k, column, doc_ind) = DOC_prod - DOC_remin;
    A3(tend, k, column, don_ind) = (DON_prod * (c1 - DONrefract)) - DON_remin;
    A3(tend, k, column, donr_ind) = (DON_prod * DONrefract) - DONr_remin + (POC.remin * DONrefract * Q);
    A3(tend, k, column, dop_ind) = (DOP_prod * (c1 - DOPrefract)) - DOP_remin - sum4(DOP_V);
    A3(tend, k, column, dopr_ind) = (DOP_prod * DOPrefract) - DOPr_remin + (POC.remin * DOPrefract * Qp_zoo_pom);
    A3(tend, k, column, dofe_ind) = DOFe_prod - DOFe_remin;

    /* :1729-1745 */
    A3(tend, k, column, dic_ind) = sum4(auto_loss_dic) + sum4(auto_graze_dic) - sum4(photoC) +
                                   DOC_remin + POC.remin + zoo_loss_dic + P_CaCO3.remin;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      if (autotrophs[auto_ind - 1].CaCO3_ind > 0)
        A3(tend, k, column, dic_ind) = A3(tend, k, column, dic_ind) +
                                       f_graze_CaCO3_remin * auto_graze[auto_ind - 1] * QCaCO3[auto_ind - 1] -
                                       CaCO3_PROD[auto_ind - 1];
    }

    if (alt_co2_use_eco) {
      A3(tend, k, column, dic_alt_co2_ind) = A3(tend, k, column, dic_ind);
    } else {
      A3(tend, k, column, dic_alt_co2_ind) = 0.0;
    }

    /* :1751-1759 */
    A3(tend, k, column, alk_ind) = -A3(tend, k, column, no3_ind) + A3(tend, k, column, nh4_ind) +
                                   c2 * P_CaCO3.remin;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      if (autotrophs[auto_ind - 1].CaCO3_ind > 0)
        A3(tend, k, column, alk_ind) = A3(tend, k, column, alk_ind) +
                                       c2 * (f_graze_CaCO3_remin * auto_graze[auto_ind - 1] * QCaCO3[auto_ind - 1] -
                                             CaCO3_PROD[auto_ind - 1]);
    }

    /* :1765-1790 */
    O2_PRODUCTION = c0;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const int a = auto_ind - 1;
      if (!autotrophs[a].Nfixer) {
        if (photoC[a] > c0) {
          O2_PRODUCTION = O2_PRODUCTION + photoC[a] *
                          ((NO3_V[a] / (NO3_V[a] + NH4_V[a])) / parm_Red_D_C_O2 +
                           (NH4_V[a] / (NO3_V[a] + NH4_V[a])) / parm_Remin_D_C_O2);
        }
      } else {
        if (photoC[a] > c0) {
          O2_PRODUCTION = O2_PRODUCTION + photoC[a] *
                          ((NO3_V[a] / (NO3_V[a] + NH4_V[a] + Nfix[a])) / parm_Red_D_C_O2 +
                           (NH4_V[a] / (NO3_V[a] + NH4_V[a] + Nfix[a])) / parm_Remin_D_C_O2 +
                           (Nfix[a] / (NO3_V[a] + NH4_V[a] + Nfix[a])) / parm_Red_D_C_O2_diaz);
        }
      }
    }

    work1 = (L1(O2_loc, k) - p->parm_o2_min) / p->parm_o2_min_delta;
    work1 = fmin(fmax(work1, 0.0), 1.0);
    O2_CONSUMPTION = work1 *
                     ((POC.remin + DOC_remin - (SED_DENITRIF * denitrif_C_N) - OTHER_REMIN + zoo_loss_dic +
                       sum4(auto_loss_dic) + sum4(auto_graze_dic)) / parm_Remin_D_C_O2 + (c2 * NITRIF));

    A3(tend, k, column, o2_ind) = O2_PRODUCTION - O2_CONSUMPTION;

    /* :1796-1868 */
    A2(d->diag_O2_PRODUCTION, k, column) = O2_PRODUCTION;
    A2(d->diag_O2_CONSUMPTION, k, column) = O2_CONSUMPTION;

    work1 = oracle_O2SAT_singleValue(TEMP, SALT, T0_Kelvin_BGC);
    work1 = work1 - L1(O2_loc, k);
    A2(d->diag_AOU, k, column) = work1;

    A2(d->diag_PAR_avg, k, column) = PAR_avg;
    A2(d->diag_zoo_loss, k, column) = zoo_loss;

    work1 = sum4(auto_graze);
    A2(d->diag_auto_graze_TOT, k, column) = work1;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const int a = auto_ind - 1;
      A3(d->diag_auto_graze, k, column, auto_ind) = auto_graze[a];
      A3(d->diag_auto_loss, k, column, auto_ind) = auto_loss[a];
      A3(d->diag_auto_agg, k, column, auto_ind) = auto_agg[a];
      A3(d->diag_photoC, k, column, auto_ind) = photoC[a];
      work1 = dz * photoC[a];
      F2(d->diag_photoC_zint, column, auto_ind) = F2(d->diag_photoC_zint, column, auto_ind) + work1;
    }

    work1 = sum4(photoC);
    A2(d->diag_photoC_TOT, k, column) = work1;
    work1 = work1 * dz;
    C1(d->diag_photoC_TOT_zint, column) = C1(d->diag_photoC_TOT_zint, column) + work1;

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      const int a = auto_ind - 1;
      if (VNtot[a] > c0) {
        work1 = (VNO3[a] / VNtot[a]) * photoC[a];
      } else {
        work1 = c0;
      }
      A3(d->diag_photoC_NO3, k, column, auto_ind) = work1;

      work1 = work1 * dz;
      F2(d->diag_photoC_NO3_zint, column, auto_ind) = F2(d->diag_photoC_NO3_zint, column, auto_ind) + work1;

      A2(d->diag_photoC_NO3_TOT, k, column) = A2(d->diag_photoC_NO3_TOT, k, column) +
                                             A3(d->diag_photoC_NO3, k, column, auto_ind);

      /* Q11: adds the RUNNING per-autotroph integral every level */
      C1(d->diag_photoC_NO3_TOT_zint, column) = C1(d->diag_photoC_NO3_TOT_zint, column) +
                                                F2(d->diag_photoC_NO3_zint, column, auto_ind);
    }

    A2(d->diag_DOC_prod, k, column) = DOC_prod;
    A2(d->diag_DOC_remin, k, column) = DOC_remin;
    A2(d->diag_DON_prod, k, column) = DON_prod;
    A2(d->diag_DON_remin, k, column) = DON_remin;
    A2(d->diag_DOP_prod, k, column) = DOP_prod;
    A2(d->diag_DOP_remin, k, column) = DOP_remin;
    A2(d->diag_DOFe_prod, k, column) = DOFe_prod;
    A2(d->diag_DOFe_remin, k, column) = DOFe_remin;
    A2(d->diag_Fe_scavenge, k, column) = Fe_scavenge;
    A2(d->diag_Fe_scavenge_rate, k, column) = Fe_scavenge_rate;

    /* :1870-1945 */
    ztop = c0;
    if (k > 1) ztop = A2(in->cell_bottom_depth, k - 1, column);
    work2 = fmin(100.0e2 - ztop, dz);
    partial_thickness_100m = (work2 > c0) ? work2 : 0.0;

    {
      double s = 0.0;
      for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind)
        s = s + A3(tend, k, column, autotrophs[auto_ind - 1].C_ind);
      work1 = A3(tend, k, column, dic_ind) + A3(tend, k, column, doc_ind) +
              A3(tend, k, column, zooC_ind) + s;
    }
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      n = autotrophs[auto_ind - 1].CaCO3_ind;
      if (n > 0) {
        work1 = work1 + A3(tend, k, column, n);
      }
    }

    C1(d->diag_Jint_Ctot, column) = C1(d->diag_Jint_Ctot, column) + work1 * dz + POC.sed_loss + P_CaCO3.sed_loss;

    C1(d->diag_Jint_100m_Ctot, column) = C1(d->diag_Jint_100m_Ctot, column) + work1 * partial_thickness_100m +
                                         ((zbot <= 100.0e2) ? (POC.sed_loss + P_CaCO3.sed_loss) : 0.0);

    {
      double s = 0.0;
      for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind)
        s = s + A3(tend, k, column, autotrophs[auto_ind - 1].C_ind);
      work1 = A3(tend, k, column, no3_ind) + A3(tend, k, column, nh4_ind) +
              A3(tend, k, column, don_ind) + A3(tend, k, column, donr_ind) +
              Q * A3(tend, k, column, zooC_ind) + Q * s;
    }
    work1 = work1 + DENITRIF + SED_DENITRIF;
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      if (autotrophs[auto_ind - 1].Nfixer) work1 = work1 - Nfix[auto_ind - 1];
    }

    C1(d->diag_Jint_Ntot, column) = C1(d->diag_Jint_Ntot, column) + work1 * dz + POC.sed_loss * Q;

    C1(d->diag_Jint_100m_Ntot, column) = C1(d->diag_Jint_100m_Ntot, column) + work1 * partial_thickness_100m +
                                         ((zbot <= 100.0e2) ? (POC.sed_loss * Q) : 0.0);

    work1 = A3(tend, k, column, po4_ind) + A3(tend, k, column, dop_ind) +
            A3(tend, k, column, dopr_ind) + Qp_zoo_pom * A3(tend, k, column, zooC_ind);
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      n = autotrophs[auto_ind - 1].C_ind;
      work1 = work1 + autotrophs[auto_ind - 1].Qp * A3(tend, k, column, n);
    }

    C1(d->diag_Jint_Ptot, column) = C1(d->diag_Jint_Ptot, column) + work1 * dz + POC.sed_loss * Qp_zoo_pom;

    C1(d->diag_Jint_100m_Ptot, column) = C1(d->diag_Jint_100m_Ptot, column) + work1 * partial_thickness_100m +
                                         ((zbot <= 100.0e2) ? (POC.sed_loss * Qp_zoo_pom) : 0.0);

    work1 = A3(tend, k, column, sio3_ind);
    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      n = autotrophs[auto_ind - 1].Si_ind;
      if (n > 0) {
        work1 = work1 + A3(tend, k, column, n);
      }
    }

    C1(d->diag_Jint_Sitot, column) = C1(d->diag_Jint_Sitot, column) + work1 * dz + P_SiO2.sed_loss;

    C1(d->diag_Jint_100m_Sitot, column) = C1(d->diag_Jint_100m_Sitot, column) + work1 * partial_thickness_100m +
                                          ((zbot <= 100.0e2) ? P_SiO2.sed_loss : 0.0);

    for (auto_ind = 1; auto_ind <= autotroph_cnt; ++auto_ind) {
      C1(d->diag_Chl_TOT_zint_100m, column) = C1(d->diag_Chl_TOT_zint_100m, column) +
                                              LA(autotrophChl_loc, k, auto_ind) * partial_thickness_100m;
    }
  }   /* k loop */

  /* :1954-1968 */
  k = 1;
  work1 = L1(O2_loc, k);
  work2 = work1;
  work3 = A2(in->cell_center_depth, k, column);
  for (k = 2; k <= kmax; ++k) {
    work1 = L1(O2_loc, k);
    if (work1 < work2) {
      work2 = work1;
      work3 = A2(in->cell_center_depth, k, column);
    }
  }
  C1(d->diag_O2_ZMIN, column) = work2;
  C1(d->diag_O2_ZMIN_DEPTH, column) = work3;
#undef L1
#undef LA
}

/* BGC_mod.F90:340-1998 */
void oracle_BGC_SourceSink(const BgcParams *p, const BgcAutotroph autotrophs[4],
                           const BgcIndices *ind, const BgcInput *in,
                           const BgcForcing *forcing, BgcOutput *out, BgcDiagnostics *d,
                           int numLevelsMax, int numColumnsMax, int numColumns,
                           int alt_co2_use_eco, int nthreads, OracleSolverStats *st) {
  const int nL = numLevelsMax, nC = numColumnsMax;
  const size_t n2 = (size_t)nL * nC, nc = (size_t)nC;
  long tr = 0, bg = 0, ni = 0, ncv = 0, pe = 0;
  int column;

  /* :570 */
  zero_fill(out->BGC_tendencies, n2 * BGC_TRACER_CNT);

  /* :625-727 — every diagnostic except diag_POC_ACCUM, diag_DONr_remin,
   * diag_DOPr_remin, which the reference never touches */
  zero_fill(d->diag_tot_CaCO3_form, n2); zero_fill(d->diag_tot_bSi_form, nc);
  zero_fill(d->diag_CaCO3_form_zint, nc * 4); zero_fill(d->diag_tot_CaCO3_form_zint, nc);
  zero_fill(d->diag_photoC_zint, nc * 4); zero_fill(d->diag_photoC_TOT_zint, nc);
  zero_fill(d->diag_tot_Nfix, n2); zero_fill(d->diag_photoC_NO3_zint, nc * 4);
  zero_fill(d->diag_photoC_NO3_TOT, n2); zero_fill(d->diag_photoC_NO3_TOT_zint, nc);
  zero_fill(d->diag_Chl_TOT_zint_100m, nc);
  zero_fill(d->diag_Jint_Ctot, nc); zero_fill(d->diag_Jint_100m_Ctot, nc);
  zero_fill(d->diag_Jint_Ntot, nc); zero_fill(d->diag_Jint_100m_Ntot, nc);
  zero_fill(d->diag_Jint_Ptot, nc); zero_fill(d->diag_Jint_100m_Ptot, nc);
  zero_fill(d->diag_Jint_Sitot, nc); zero_fill(d->diag_Jint_100m_Sitot, nc);

  zero_fill(d->diag_CO3, n2); zero_fill(d->diag_HCO3, n2); zero_fill(d->diag_H2CO3, n2);
  zero_fill(d->diag_pH_3D, n2); zero_fill(d->diag_CO3_ALT_CO2, n2);
  zero_fill(d->diag_HCO3_ALT_CO2, n2); zero_fill(d->diag_H2CO3_ALT_CO2, n2);
  zero_fill(d->diag_pH_3D_ALT_CO2, n2); zero_fill(d->diag_co3_sat_calc, n2);
  zero_fill(d->diag_co3_sat_arag, n2); zero_fill(d->diag_NO3_RESTORE, n2);
  zero_fill(d->diag_NITRIF, n2); zero_fill(d->diag_DENITRIF, n2);
  zero_fill(d->diag_SiO3_RESTORE, n2); zero_fill(d->diag_PO4_RESTORE, n2);
  zero_fill(d->diag_O2_PRODUCTION, n2); zero_fill(d->diag_O2_CONSUMPTION, n2);
  zero_fill(d->diag_AOU, n2); zero_fill(d->diag_PAR_avg, n2); zero_fill(d->diag_zoo_loss, n2);
  zero_fill(d->diag_auto_graze_TOT, n2); zero_fill(d->diag_photoC_TOT, n2);
  zero_fill(d->diag_DOC_prod, n2); zero_fill(d->diag_DOC_remin, n2);
  zero_fill(d->diag_DON_prod, n2); zero_fill(d->diag_DON_remin, n2);
  zero_fill(d->diag_DOP_prod, n2); zero_fill(d->diag_DOP_remin, n2);
  zero_fill(d->diag_DOFe_prod, n2); zero_fill(d->diag_DOFe_remin, n2);
  zero_fill(d->diag_Fe_scavenge, n2); zero_fill(d->diag_Fe_scavenge_rate, n2);
  zero_fill(d->diag_POC_FLUX_IN, n2); zero_fill(d->diag_POC_PROD, n2);
  zero_fill(d->diag_POC_REMIN, n2); zero_fill(d->diag_CaCO3_FLUX_IN, n2);
  zero_fill(d->diag_CaCO3_PROD, n2); zero_fill(d->diag_CaCO3_REMIN, n2);
  zero_fill(d->diag_SiO2_FLUX_IN, n2); zero_fill(d->diag_SiO2_PROD, n2);
  zero_fill(d->diag_SiO2_REMIN, n2); zero_fill(d->diag_dust_FLUX_IN, n2);
  zero_fill(d->diag_dust_REMIN, n2); zero_fill(d->diag_P_iron_FLUX_IN, n2);
  zero_fill(d->diag_P_iron_PROD, n2); zero_fill(d->diag_P_iron_REMIN, n2);
  zero_fill(d->diag_calcToSed, n2); zero_fill(d->diag_bsiToSed, n2);
  zero_fill(d->diag_pocToSed, n2); zero_fill(d->diag_SedDenitrif, n2);
  zero_fill(d->diag_OtherRemin, n2); zero_fill(d->diag_ponToSed, n2);
  zero_fill(d->diag_popToSed, n2); zero_fill(d->diag_dustToSed, n2);
  zero_fill(d->diag_pfeToSed, n2);

  zero_fill(d->diag_zsatcalc, nc); zero_fill(d->diag_zsatarag, nc);
  zero_fill(d->diag_O2_ZMIN, nc); zero_fill(d->diag_O2_ZMIN_DEPTH, nc);

  zero_fill(d->diag_N_lim, n2 * 4); zero_fill(d->diag_Fe_lim, n2 * 4);
  zero_fill(d->diag_P_lim, n2 * 4); zero_fill(d->diag_SiO3_lim, n2 * 4);
  zero_fill(d->diag_light_lim, n2 * 4); zero_fill(d->diag_photoNO3, n2 * 4);
  zero_fill(d->diag_photoNH4, n2 * 4); zero_fill(d->diag_PO4_uptake, n2 * 4);
  zero_fill(d->diag_DOP_uptake, n2 * 4); zero_fill(d->diag_photoFe, n2 * 4);
  zero_fill(d->diag_bSi_form, n2 * 4); zero_fill(d->diag_CaCO3_form, n2 * 4);
  zero_fill(d->diag_Nfix, n2 * 4); zero_fill(d->diag_auto_graze, n2 * 4);
  zero_fill(d->diag_auto_loss, n2 * 4); zero_fill(d->diag_auto_agg, n2 * 4);
  zero_fill(d->diag_photoC, n2 * 4); zero_fill(d->diag_photoC_NO3, n2 * 4);

  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : tr, bg, ni, ncv, pe)
  {
    double *scratch = (double *)malloc(sizeof(double) * 36 * (size_t)nL);
    OracleSolverStats s = {0, 0, 0, 0};
    long poc_errors = 0;
#pragma omp for schedule(dynamic, 64)
    for (column = 1; column <= numColumns; ++column) {
      source_sink_column(p, autotrophs, ind, in, forcing, out, d, nL, nC, column,
                         alt_co2_use_eco, scratch, &s, &poc_errors);
    }
    free(scratch);
    tr += s.talk_row_calls; bg += s.bracket_grow; ni += s.newton_iters; ncv += s.no_convergence;
    pe += poc_errors;
  }
  if (st) {
    st->talk_row_calls += tr; st->bracket_grow += bg; st->newton_iters += ni;
    st->no_convergence += ncv;
  }
  (void)pe;
}

/* BGC_mod.F90:2706-2957 */
void oracle_BGC_SurfaceFluxes(const BgcParams *p, const BgcIndices *ind, const BgcInput *in,
                              BgcForcing *f, BgcFluxDiagnostics *d, int numLevelsMax,
                              int numColumnsMax, int numColumns, int nthreads) {
  const int nL = numLevelsMax, nC = numColumnsMax;
  const size_t nc = (size_t)nC;
  const double *tr = in->BGC_tracers;
  int column;

  /* :2789-2802 */
  zero_fill(d->pistonVel_O2, nc); zero_fill(d->pistonVel_CO2, nc);
  zero_fill(d->SCHMIDT_O2, nc); zero_fill(d->SCHMIDT_CO2, nc); zero_fill(d->O2SAT, nc);
  zero_fill(d->xkw, nc); zero_fill(d->co2star, nc); zero_fill(d->dco2star, nc);
  zero_fill(d->pco2surf, nc); zero_fill(d->dpco2, nc); zero_fill(d->co2star_alt_co2, nc);
  zero_fill(d->dco2star_alt_co2, nc); zero_fill(d->pco2surf_alt_co2, nc);
  zero_fill(d->dpco2_alt_co2, nc);

  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (column = 1; column <= numColumns; ++column) {
    double xkw, xkw_ice, SCHMIDT_O2, O2SAT_1atm, pistonVel_O2, O2SAT, depth, SCHMIDT_CO2,
           pistonVel_CO2, phlo, phhi, ph_new, co2star, dco2star, pco2surf, dpco2;
    double DIC_loc, DIC_ALT_CO2_loc, ALK_loc, PO4_loc, NO3_loc, SiO3_loc, O2_loc;
    int n;

    /* :2816-2822 */
    DIC_loc = fmax(0.0, A3(tr, 1, column, ind->dic_ind));
    DIC_ALT_CO2_loc = fmax(0.0, A3(tr, 1, column, ind->dic_alt_co2_ind));
    ALK_loc = fmax(0.0, A3(tr, 1, column, ind->alk_ind));
    PO4_loc = fmax(0.0, A3(tr, 1, column, ind->po4_ind));
    NO3_loc = fmax(0.0, A3(tr, 1, column, ind->no3_ind));
    SiO3_loc = fmax(0.0, A3(tr, 1, column, ind->sio3_ind));
    O2_loc = fmax(0.0, A3(tr, 1, column, ind->o2_ind));
    (void)NO3_loc;

    /* :2828-2835  (Q14: in-place side effects on the inputs) */
    F2(f->depositionFlux, column, ind->fe_ind) = F2(f->depositionFlux, column, ind->fe_ind) * p->parm_Fe_bioavail;
    F2(f->riverFlux, column, ind->fe_ind) = F2(f->riverFlux, column, ind->fe_ind) * p->parm_Fe_bioavail;
    F2(f->gasFlux, column, ind->fe_ind) = F2(f->gasFlux, column, ind->fe_ind) * p->parm_Fe_bioavail;
    F2(f->seaIceFlux, column, ind->fe_ind) = F2(f->seaIceFlux, column, ind->fe_ind) * p->parm_Fe_bioavail;

    if (C1(f->iceFraction, column) < 0.0) C1(f->iceFraction, column) = 0.0;
    if (C1(f->iceFraction, column) > 1.0) C1(f->iceFraction, column) = 1.0;

    xkw = xkw_coeff * C1(f->windSpeedSquared10m, column);
    xkw_ice = (1.0 - C1(f->iceFraction, column)) * xkw;

    /* :2847-2860 */
    if (f->lcalc_O2_gas_flux) {
      SCHMIDT_O2 = oracle_SCHMIDT_O2_singleValue(C1(f->SST, column));
      O2SAT_1atm = oracle_O2SAT_singleValue(C1(f->SST, column), C1(f->SSS, column), p->T0_Kelvin_BGC);

      pistonVel_O2 = xkw_ice * sqrt(660.0 / SCHMIDT_O2);
      O2SAT = C1(f->surfacePressure, column) * O2SAT_1atm;
      F2(f->gasFlux, column, ind->o2_ind) = pistonVel_O2 * (O2SAT - O2_loc);

      C1(d->pistonVel_O2, column) = pistonVel_O2;
      C1(d->SCHMIDT_O2, column) = SCHMIDT_O2;
      C1(d->O2SAT, column) = O2SAT;
      C1(d->xkw, column) = xkw_ice;
    }

    /* :2866-2923 */
    if (f->lcalc_CO2_gas_flux) {
      SCHMIDT_CO2 = oracle_SCHMIDT_CO2_singleValue(C1(f->SST, column));
      pistonVel_CO2 = xkw_ice * sqrt(660.0 / SCHMIDT_CO2);

      if (C1(f->surface_pH, column) != c0) {
        phlo = C1(f->surface_pH, column) - del_ph;
        phhi = C1(f->surface_pH, column) + del_ph;
      } else {
        phlo = phlo_surf_init;
        phhi = phhi_surf_init;
      }

      depth = C1(f->surfaceDepth, column);
      oracle_co2calc_1point(depth, 1, 1, C1(f->SST, column), C1(f->SSS, column), DIC_loc,
                            ALK_loc, PO4_loc, SiO3_loc, &phlo, &phhi, &ph_new,
                            C1(f->atmCO2, column), C1(f->surfacePressure, column), &co2star,
                            &dco2star, &pco2surf, &dpco2, NULL);

      C1(f->surface_pH, column) = ph_new;
      F2(f->gasFlux, column, ind->dic_ind) = pistonVel_CO2 * dco2star;

      C1(d->co2star, column) = co2star;
      C1(d->dco2star, column) = dco2star;
      C1(d->pco2surf, column) = pco2surf;
      C1(d->dpco2, column) = dpco2;
      C1(d->pistonVel_CO2, column) = pistonVel_CO2;
      C1(d->SCHMIDT_CO2, column) = SCHMIDT_CO2;

      if (C1(f->surface_pH_alt_co2, column) != c0) {
        phlo = C1(f->surface_pH_alt_co2, column) - del_ph;
        phhi = C1(f->surface_pH_alt_co2, column) + del_ph;
      } else {
        phlo = phlo_surf_init;
        phhi = phhi_surf_init;
      }

      oracle_co2calc_1point(depth, 1, 1, C1(f->SST, column), C1(f->SSS, column), DIC_ALT_CO2_loc,
                            ALK_loc, PO4_loc, SiO3_loc, &phlo, &phhi, &ph_new,
                            C1(f->atmCO2_ALT_CO2, column), C1(f->surfacePressure, column),
                            &co2star, &dco2star, &pco2surf, &dpco2, NULL);

      C1(f->surface_pH_alt_co2, column) = ph_new;
      F2(f->gasFlux, column, ind->dic_alt_co2_ind) = pistonVel_CO2 * dco2star;

      C1(d->co2star_alt_co2, column) = co2star;
      C1(d->dco2star_alt_co2, column) = dco2star;
      C1(d->pco2surf_alt_co2, column) = pco2surf;
      C1(d->dpco2_alt_co2, column) = dpco2;
    }

    /* :2929-2942 */
    for (n = 1; n <= BGC_TRACER_CNT; ++n) {
      F2(f->netFlux, column, n) = F2(f->depositionFlux, column, n) + F2(f->gasFlux, column, n) +
                                  F2(f->riverFlux, column, n) + F2(f->seaIceFlux, column, n);
    }

    F2(f->netFlux, column, ind->alk_ind) = F2(f->netFlux, column, ind->alk_ind) +
                                           F2(f->netFlux, column, ind->nh4_ind) -
                                           F2(f->netFlux, column, ind->no3_ind);
  }
  (void)nL;
}
