/* dms_macros_oracle.c — restatement of DMS_SourceSink, DMS_SurfaceFluxes
 * (DMS_mod.F90) and MACROS_SourceSink (MACROS_mod.F90).
 * TEST INFRASTRUCTURE ONLY (see bgc_oracle.h).  Pinned against the translated reference. */
#include "bgc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A2(p, k, c) ((p)[(size_t)((k)-1) + (size_t)nL * (size_t)((c)-1)])
#define A3(p, k, c, n) ((p)[(size_t)((k)-1) + (size_t)nL * ((size_t)((c)-1) + (size_t)nC * (size_t)((n)-1))])
#define F2(p, c, n) ((p)[(size_t)((c)-1) + (size_t)nC * (size_t)((n)-1)])
#define C1(p, c) ((p)[(size_t)((c)-1)])

static const double epsC = 1.00e-8;   /* DMS_parms.F90:194-195 (carries _r8: exact double) */

/* DMS_mod.F90:915-959 */
double oracle_SCHMIDT_DMS_singleValue(double SST) {
  const double a = 2674.0, b = 147.12, c = 3.726, d = 0.038;
  return a + SST * (-b + SST * (c + SST * (-d)));
}

/* DMS_mod.F90:966-1008 */
static double DMSSAT_singleValue(double SST, double SSS) {
  (void)SST; (void)SSS;
  return 0.0;
}

/* DMS_mod.F90:156-770 */
void oracle_DMS_SourceSink(const DmsParams *p, const DmsIndices *ind, const DmsInput *in,
                           const DmsForcing *forcing, DmsOutput *out, DmsDiagnostics *d,
                           int numLevelsMax, int numColumnsMax, int numColumns, int nthreads) {
  const int nL = numLevelsMax, nC = numColumnsMax;
  const double *tr = in->DMS_tracers;
  double *tend = out->DMS_tendencies;
  int column;

  const int no3_ind = ind->no3_ind, doc_ind = ind->doc_ind, zooC_ind = ind->zooC_ind,
            spC_ind = ind->spC_ind, diatC_ind = ind->diatC_ind, diazC_ind = ind->diazC_ind,
            phaeoC_ind = ind->phaeoC_ind, spChl_ind = ind->spChl_ind,
            diatChl_ind = ind->diatChl_ind, diazChl_ind = ind->diazChl_ind,
            phaeoChl_ind = ind->phaeoChl_ind, spCaCO3_ind = ind->spCaCO3_ind,
            dms_ind = ind->dms_ind, dmsp_ind = ind->dmsp_ind;

  /* :413 */
  memset(tend, 0, sizeof(double) * (size_t)nL * nC * DMS_TRACER_CNT);

  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
  for (column = 1; column <= numColumns; ++column) {
    double totalChl, PAR_out, PAR_in, KPARdz, PAR_avg, UV_out, UV_in, KUVdz, UV_avg;
    double Fcocco;
    double diatN_loc, phaeoN_loc, coccoN_loc, cyanoN_loc, eukarN_loc, diazN_loc, phytoN_loc,
           zooN_loc;
    double diatS_loc, phaeoS_loc, coccoS_loc, cyanoS_loc, eukarS_loc, diazS_loc, phytoS_loc,
           zooS_loc;
    double k_S_p, yield, B_diagnosed, j_dms, T_ind, Cocco_frac, Cyano_frac, Eukar_frac,
           Sp_dec, Stress_fac, SST_loc, Rs2n_zoo;
    double dms_s_dmsp, dms_s, dms_r_B, dms_r_phot, dms_r_bkgnd, dms_r, dmsp_s_phaeo,
           dmsp_s_nonphaeo, dmsp_s_zoo, dmsp_s, dmsp_r_B, dmsp_r_bkgnd, dmsp_r, work;
    int kmax, k;

    kmax = in->number_of_active_levels[column - 1];
    if (kmax < 1) continue;

    /* :504-510 */
    SST_loc = C1(forcing->SST, column);
    PAR_out = fmax(0.0, C1(forcing->ShortWaveFlux_surface, column));
    PAR_out = PAR_out * p->f_qsw_par_DMS;
    UV_out = PAR_out * 0.01;

    for (k = 1; k <= kmax; ++k) {
      /* setup_loop, :471-485 */
      const double NO3_loc = fmax(0.0, A3(tr, k, column, no3_ind));
      const double DOC_loc = fmax(0.0, A3(tr, k, column, doc_ind));
      const double zooC_loc = fmax(0.0, A3(tr, k, column, zooC_ind));
      const double spC_loc = fmax(0.0, A3(tr, k, column, spC_ind));
      const double diatC_loc = fmax(0.0, A3(tr, k, column, diatC_ind));
      const double diazC_loc = fmax(0.0, A3(tr, k, column, diazC_ind));
      const double phaeoC_loc = fmax(0.0, A3(tr, k, column, phaeoC_ind));
      const double spChl_loc = fmax(0.0, A3(tr, k, column, spChl_ind));
      const double diatChl_loc = fmax(0.0, A3(tr, k, column, diatChl_ind));
      const double diazChl_loc = fmax(0.0, A3(tr, k, column, diazChl_ind));
      const double phaeoChl_loc = fmax(0.0, A3(tr, k, column, phaeoChl_ind));
      const double spCaCO3_loc = fmax(0.0, A3(tr, k, column, spCaCO3_ind));
      const double DMS_loc = fmax(0.0, A3(tr, k, column, dms_ind));
      const double DMSP_loc = fmax(0.0, A3(tr, k, column, dmsp_ind));
      const double dz = A2(in->cell_thickness, k, column);
      (void)NO3_loc;

      /* :529  (Q13: literal 0.3, not zooC_avg) */
      k_S_p = p->k_S_p_base * (p->mort + (zooC_loc / 0.3));

      /* :531-536  (dead: UV never reaches an output) */
      UV_in = UV_out;
      KUVdz = (0.01e-2 * DOC_loc + 0.04e-4) * dz;
      UV_out = UV_in * exp(-KUVdz);
      UV_avg = UV_in * (1.0 - exp(-KUVdz)) / KUVdz;
      (void)UV_avg;

      /* :538-551 */
      PAR_in = PAR_out;
      totalChl = spChl_loc + diatChl_loc + diazChl_loc + phaeoChl_loc;
      work = fmax(totalChl, 0.02);
      if (work < 0.13224) {
        KPARdz = 0.000919 * pow(work, 0.3536);
      } else {
        KPARdz = 0.001131 * pow(work, 0.4562);
      }
      KPARdz = KPARdz * dz;

      PAR_out = PAR_in * exp(-KPARdz);
      PAR_avg = PAR_in * (1.0 - exp(-KPARdz)) / KPARdz;

      j_dms = p->j_dms_perI * PAR_avg;   /* :562 */

      /* :570-573 */
      Fcocco = spCaCO3_loc / (spC_loc + epsC);
      if (Fcocco > 0.4) Fcocco = 0.4;
      Cocco_frac = Fcocco;

      /* :584-592 */
      T_ind = (SST_loc - p->T_lo) / (p->T_hi - p->T_lo);
      if (T_ind <= 0.0) T_ind = 0.0;
      if (T_ind >= 1.0) T_ind = 1.0;

      Cyano_frac = (T_ind * (p->Max_cyano_frac - p->Min_cyano_frac)) + p->Min_cyano_frac;
      Cyano_frac = (1.0 - Cocco_frac) * Cyano_frac;
      Eukar_frac = 1.0 - Cocco_frac - Cyano_frac;

      /* :598-612 */
      diatN_loc = p->R * diatC_loc;
      phaeoN_loc = p->R * phaeoC_loc;
      coccoN_loc = Cocco_frac * p->R * spC_loc;
      cyanoN_loc = Cyano_frac * p->R * spC_loc;
      eukarN_loc = Eukar_frac * p->R * spC_loc;
      diazN_loc = p->R * diazC_loc;
      zooN_loc = p->R * zooC_loc;

      phytoN_loc = diatN_loc + coccoN_loc + cyanoN_loc + eukarN_loc + diazN_loc + phaeoN_loc;

      /* :621-628 */
      Sp_dec = (p->Sp_ref - spChl_loc) / p->Sp_ref;
      if (Sp_dec <= 0.0) Sp_dec = 0.0;
      if (Sp_dec >= 1.0) Sp_dec = 1.0;
      Stress_fac = 1.0 + p->Stress_mult * Sp_dec * Sp_dec;
      if (Stress_fac >= 10.0) Stress_fac = 10.0;

      /* :637-640 */
      yield = (T_ind * (p->Max_yld - p->Min_yld)) + p->Min_yld;
      if (SST_loc < p->T_cryo_hi && SST_loc > p->T_cryo_lo) yield = 0.5;
      if (SST_loc < -1.0) yield = 0.25;

      /* :647-660 */
      diatS_loc = p->Rs2n_diat * diatN_loc;
      phaeoS_loc = p->Rs2n_phaeo * phaeoN_loc;
      coccoS_loc = p->Rs2n_cocco * coccoN_loc;
      cyanoS_loc = p->Rs2n_cyano * cyanoN_loc;
      eukarS_loc = p->Rs2n_eukar * eukarN_loc * Stress_fac;
      diazS_loc = p->Rs2n_diaz * diazN_loc;

      phytoS_loc = diatS_loc + coccoS_loc + cyanoS_loc + eukarS_loc + diazS_loc +
                   p->G_phaeo_S * phaeoS_loc;

      /* :671-684 */
      if (phytoN_loc > 0.0) {
        Rs2n_zoo = (p->Rs2n_diat * diatN_loc +
                    p->G_phaeo_S * p->Rs2n_phaeo * phaeoN_loc +
                    p->Rs2n_cocco * coccoN_loc +
                    p->Rs2n_cyano * cyanoN_loc +
                    p->Rs2n_eukar * eukarN_loc * Stress_fac +
                    p->Rs2n_diaz * diazN_loc) / phytoN_loc;
      } else {
        Rs2n_zoo = (p->Rs2n_diat + p->Rs2n_cocco + p->Rs2n_cyano + p->Rs2n_eukar + p->Rs2n_diaz +
                    p->Rs2n_phaeo) / 6.0;
      }
      zooS_loc = Rs2n_zoo * zooN_loc;

      B_diagnosed = p->B_preexp * pow(phytoN_loc, p->B_exp);   /* :695 */

      /* :701-719 */
      dms_s_dmsp = yield * p->k_conv * DMSP_loc;
      dms_s = dms_s_dmsp;

      dms_r_B = p->k_S_B * B_diagnosed * DMS_loc;
      dms_r_phot = j_dms * DMS_loc;
      dms_r_bkgnd = p->k_bkgnd * DMS_loc;
      dms_r = dms_r_B + dms_r_phot + dms_r_bkgnd;

      dmsp_s_phaeo = p->inject_scale * p->k_S_p_base * phaeoS_loc;
      dmsp_s_nonphaeo = p->inject_scale * k_S_p * phytoS_loc;
      dmsp_s_zoo = p->inject_scale * p->k_S_z * zooS_loc;
      dmsp_s = dmsp_s_phaeo + dmsp_s_nonphaeo + dmsp_s_zoo;

      dmsp_r_B = p->k_conv * DMSP_loc;
      dmsp_r_bkgnd = p->k_bkgnd * DMSP_loc;
      dmsp_r = dmsp_r_B + dmsp_r_bkgnd;

      A3(tend, k, column, dms_ind) = dms_s - dms_r;
      A3(tend, k, column, dmsp_ind) = dmsp_s - dmsp_r;

      /* :723-761 */
      A2(d->diag_DMS_S_DMSP, k, column) = dms_s_dmsp;
      A2(d->diag_DMS_S_TOTAL, k, column) = dms_s;
      A2(d->diag_DMS_R_B, k, column) = dms_r_B;
      A2(d->diag_DMS_R_PHOT, k, column) = dms_r_phot;
      A2(d->diag_DMS_R_BKGND, k, column) = dms_r_bkgnd;
      A2(d->diag_DMS_R_TOTAL, k, column) = dms_r;
      A2(d->diag_DMSP_S_PHAEO, k, column) = dmsp_s_phaeo;
      A2(d->diag_DMSP_S_NONPHAEO, k, column) = dmsp_s_nonphaeo;
      A2(d->diag_DMSP_S_ZOO, k, column) = dmsp_s_zoo;
      A2(d->diag_DMSP_S_TOTAL, k, column) = dmsp_s;
      A2(d->diag_DMSP_R_B, k, column) = dmsp_r_B;
      A2(d->diag_DMSP_R_BKGND, k, column) = dmsp_r_bkgnd;
      A2(d->diag_DMSP_R_TOTAL, k, column) = dmsp_r;
      A2(d->diag_Cyano_frac, k, column) = Cyano_frac;
      A2(d->diag_Cocco_frac, k, column) = Cocco_frac;
      A2(d->diag_Eukar_frac, k, column) = Eukar_frac;
      A2(d->diag_diatS, k, column) = diatS_loc;
      A2(d->diag_diatN, k, column) = diatN_loc;
      A2(d->diag_phytoN, k, column) = phytoN_loc;
      A2(d->diag_coccoS, k, column) = coccoS_loc;
      A2(d->diag_cyanoS, k, column) = cyanoS_loc;
      A2(d->diag_eukarS, k, column) = eukarS_loc;
      A2(d->diag_diazS, k, column) = diazS_loc;
      A2(d->diag_phaeoS, k, column) = phaeoS_loc;
      A2(d->diag_zooS, k, column) = zooS_loc;
      A2(d->diag_zooCC, k, column) = zooC_loc;
      A2(d->diag_RSNzoo, k, column) = Rs2n_zoo;
    }
  }
}

/* DMS_mod.F90:778-908 */
void oracle_DMS_SurfaceFluxes(const DmsParams *p, const DmsIndices *ind, const DmsInput *in,
                              DmsForcing *f, DmsFluxDiagnostics *d, int numLevelsMax,
                              int numColumnsMax, int numColumns) {
  const int nL = numLevelsMax, nC = numColumnsMax;
  const double *tr = in->DMS_tracers;
  const double a = 0.31, e2 = 2.85, e3 = 0.612;   /* :831-838 (e1,e4,e5,e6 unused) */
  int column;
  (void)p;

  if (f->lcalc_DMS_gas_flux) {
    for (column = 1; column <= numColumns; ++column) {
      double seaSurfaceTemp, seaSurfaceSalt, seaSurfaceDMS, xkw = 0.0, xkw_ice, SCHMIDT_DMS,
             DMSSAT_1atm, pistonVel_DMS, DMSSAT, WIND_SPEED, FW92, FLM86, XKW_W92, XKW_LM86;

      seaSurfaceDMS = fmax(0.0, A3(tr, 1, column, ind->dms_ind));
      seaSurfaceTemp = C1(f->SST, column);
      seaSurfaceSalt = C1(f->SSS, column);

      if (C1(f->iceFraction, column) < 0.0) C1(f->iceFraction, column) = 0.0;
      if (C1(f->iceFraction, column) > 1.0) C1(f->iceFraction, column) = 1.0;

      SCHMIDT_DMS = oracle_SCHMIDT_DMS_singleValue(seaSurfaceTemp);

      WIND_SPEED = sqrt(fabs(C1(f->windSpeedSquared10m, column))) * 0.01;   /* :866 */

      XKW_W92 = a * (pow((660.0 / SCHMIDT_DMS), 0.500)) * WIND_SPEED * WIND_SPEED;
      XKW_LM86 = e2 * (pow((600.0 / SCHMIDT_DMS), 0.500)) * (WIND_SPEED - 3.6) +
                 e3 * (pow((600.0 / SCHMIDT_DMS), 0.667));

      if (WIND_SPEED < 3.6) xkw = XKW_W92;
      if ((WIND_SPEED >= 3.6) && (WIND_SPEED < 5.6)) {
        FLM86 = 0.5 * (WIND_SPEED - 3.6);
        FW92 = 1.0 - FLM86;
        xkw = FW92 * XKW_W92 + FLM86 * XKW_LM86;
      }
      if (WIND_SPEED >= 5.6) xkw = XKW_LM86;

      xkw = xkw / 3600.0;
      xkw_ice = (1.0 - C1(f->iceFraction, column)) * xkw;

      DMSSAT_1atm = DMSSAT_singleValue(seaSurfaceTemp, seaSurfaceSalt);

      pistonVel_DMS = xkw_ice * sqrt(660.0 / SCHMIDT_DMS);
      DMSSAT = C1(f->surfacePressure, column) * DMSSAT_1atm;
      F2(f->netFlux, column, ind->dms_ind) = pistonVel_DMS * (DMSSAT - seaSurfaceDMS);
      F2(f->netFlux, column, ind->dmsp_ind) = 0.0;

      C1(d->diag_DMS_IFRAC, column) = C1(f->iceFraction, column);
      C1(d->diag_DMS_XKW, column) = xkw_ice;
      C1(d->diag_DMS_ATM_PRESS, column) = C1(f->surfacePressure, column);
      C1(d->diag_DMS_PV, column) = pistonVel_DMS;
      C1(d->diag_DMS_SCHMIDT, column) = SCHMIDT_DMS;
      C1(d->diag_DMS_SAT, column) = DMSSAT;
      C1(d->diag_DMS_SURF, column) = seaSurfaceDMS;
      C1(d->diag_DMS_WS, column) = WIND_SPEED;
    }
  }
}

/* MACROS_mod.F90:137-411 */
void oracle_MACROS_SourceSink(const MacrosParams *p, const MacrosIndices *ind,
                              const MacrosInput *in, MacrosOutput *out, MacrosDiagnostics *d,
                              int numLevelsMax, int numColumnsMax, int numColumns,
                              int nthreads) {
  const int nL = numLevelsMax, nC = numColumnsMax;
  const double *tr = in->MACROS_tracers;
  double *tend = out->MACROS_tendencies;
  int column;

  memset(tend, 0, sizeof(double) * (size_t)nL * nC * MACROS_TRACER_CNT);   /* :267 */

  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
  for (column = 1; column <= numColumns; ++column) {
    int kmax, k;
    kmax = in->number_of_active_levels[column - 1];
    if (kmax < 1) continue;

    for (k = 1; k <= kmax; ++k) {
      /* :309-319 */
      const double zooC_loc = fmax(0.0, A3(tr, k, column, ind->zooC_ind));
      const double spC_loc = fmax(0.0, A3(tr, k, column, ind->spC_ind));
      const double diatC_loc = fmax(0.0, A3(tr, k, column, ind->diatC_ind));
      const double diazC_loc = fmax(0.0, A3(tr, k, column, ind->diazC_ind));
      const double phaeoC_loc = fmax(0.0, A3(tr, k, column, ind->phaeoC_ind));
      const double prot_loc = fmax(0.0, A3(tr, k, column, ind->prot_ind));
      const double poly_loc = fmax(0.0, A3(tr, k, column, ind->poly_ind));
      const double lip_loc = fmax(0.0, A3(tr, k, column, ind->lip_ind));
      double k_C_p, spCk_loc, diatCk_loc, diazCk_loc, phaeoCk_loc, phytoC_loc;
      double prot_s_disr, poly_s_disr, lip_s_disr, prot_r_bac, poly_r_bac, lip_r_bac, prot_s,
             poly_s, lip_s, prot_r, poly_r, lip_r;

      k_C_p = p->k_C_p_base * (p->mort + (zooC_loc / p->zooC_avg));   /* :349 */

      spCk_loc = spC_loc;
      diatCk_loc = diatC_loc;
      phaeoCk_loc = phaeoC_loc;
      diazCk_loc = diazC_loc;

      phytoC_loc = diatCk_loc + phaeoCk_loc + spCk_loc + diazCk_loc;   /* :366 */

      /* :372-390 */
      prot_s_disr = p->inject_scale * p->f_prot * k_C_p * phytoC_loc;
      poly_s_disr = p->inject_scale * p->f_poly * k_C_p * phytoC_loc;
      lip_s_disr = p->inject_scale * p->f_lip * k_C_p * phytoC_loc;

      prot_r_bac = p->k_prot_bac * prot_loc;
      poly_r_bac = p->k_poly_bac * poly_loc;
      lip_r_bac = p->k_lip_bac * lip_loc;

      prot_s = prot_s_disr;
      poly_s = poly_s_disr;
      lip_s = lip_s_disr;
      prot_r = prot_r_bac;
      poly_r = poly_r_bac;
      lip_r = lip_r_bac;

      A3(tend, k, column, ind->prot_ind) = prot_s - prot_r;
      A3(tend, k, column, ind->poly_ind) = poly_s - poly_r;
      A3(tend, k, column, ind->lip_ind) = lip_s - lip_r;

      /* :396-402 */
      A2(d->diag_PROT_S_TOTAL, k, column) = prot_s;
      A2(d->diag_POLY_S_TOTAL, k, column) = poly_s;
      A2(d->diag_LIP_S_TOTAL, k, column) = lip_s;
      A2(d->diag_PROT_R_TOTAL, k, column) = prot_r;
      A2(d->diag_POLY_R_TOTAL, k, column) = poly_r;
      A2(d->diag_LIP_R_TOTAL, k, column) = lip_r;
    }
  }
}
