/* bgc_host_parms.c — host-side defaults of the parameter tables and the index
 * wiring, as the reference initialisers produce them.  No device needed.
 *
 *   bgc_parms_init     <- BGC_parms_init      BGC_parms.F90:497-699
 *   bgc_init           <- BGC_init (autotroph index wiring only) BGC_mod.F90:271-321
 *   dms_parms_init     <- DMS_parms_init      DMS_parms.F90:203-241
 *   macros_parms_init  <- MACROS_parms_init   MACROS_parms.F90:143-162
 *
 * The tracer slots of BGC_indices_type / DMS_indices_type / MACROS_indices_type
 * are chosen by the HOST model (they are not set anywhere in the reference);
 * *_default_tracer_indices fill them 1..N in declaration order for callers
 * that have no preference.
 */
#include "bgc_b200.h"
#include <string.h>

#define SPD 86400.0
#define DPS (1.0 / SPD)   /* BGC_parms.F90:37-40 */

/* One row per functional group, columns in the order of BgcAutotroph's double
 * members.  Rates quoted per day in the reference are multiplied by dps below. */
typedef struct AutoRow {
  int Nfixer, imp_calcifier, exp_calcifier, grazee, temp_function;
  double kFe, kPO4, kDOP, kNO3, kNH4, kSiO3, Qp, gQfe_0, gQfe_min, alphaPI_pd, PCref_pd,
         thetaN_max, loss_thres, loss_thres2, temp_thres, temp_thresS, temp_thresN,
         temp_optN, temp_optS, mort_pd, mort2_pd, agg_rate_max, agg_rate_min, z_umax_0_pd,
         z_grz, graze_zoo, graze_poc, graze_doc, loss_poc, f_zoo_detr;
} AutoRow;

/* BGC_parms.F90:543-697; order sp, diat, diaz, phaeo (:515-518).  phaeo shares
 * the diatom grazee class (:666) and is the only quasi-MMRT group (:684). */
static const AutoRow k_rows[BGC_AUTOTROPH_CNT] = {
  /* sp */
  {0, 1, 0, 1, BGC_TFNC_Q10,
   0.04e-3, 0.01, 0.26, 0.1, 0.01, 0.0, 0.00855, 20.0e-6, 3.0e-6, 0.6, 5.5,
   2.5, 0.04, 0.0, -20.0, -20.0, -20.0, 50.0, 50.0, 0.12, 0.001, 0.9, 0.01, 3.3,
   1.05, 0.3, 0.0, 0.15, 0.0, 0.15},
  /* diat */
  {0, 0, 0, 2, BGC_TFNC_Q10,
   0.06e-3, 0.05, 0.9, 0.5, 0.05, 0.8, 0.00855, 20.0e-6, 3.0e-6, 0.465, 5.5,
   4.0, 0.04, 0.0, -20.0, 10.0, 35.0, 16.3, 5.0, 0.12, 0.001, 0.9, 0.02, 3.23,
   1.0, 0.3, 0.42, 0.15, 0.0, 0.2},
  /* diaz */
  {1, 0, 0, 3, BGC_TFNC_Q10,
   0.04e-3, 0.02, 0.09, 1.0, 0.15, 0.0, 0.002735, 60.0e-6, 12.0e-6, 0.4, 0.7,
   2.5, 0.022, 0.001, 14.0, -20.0, -20.0, 50.0, 50.0, 0.15, 0.0, 0.0, 0.0, 0.6,
   1.2, 0.3, 0.05, 0.15, 0.0, 0.15},
  /* phaeo */
  {0, 0, 0, 2, BGC_TFNC_QUASI_MMRT,
   0.075e-3, 0.05, 0.9, 0.7, 0.05, 0.0, 0.00855, 20.0e-6, 3.0e-6, 0.77, 5.5,
   2.5, 0.04, 0.0, -20.0, 10.0, 35.0, 16.3, 5.0, 0.12, 0.001, 0.9, 0.02, 3.23,
   1.0, 0.3, 0.42, 0.15, 0.0, 0.2},
};

int bgc_parms_init(BgcParams *p, BgcAutotroph autotrophs[BGC_AUTOTROPH_CNT], BgcIndices *ind) {
  static const double scalelen_z[4] = {130.0e2, 290.0e2, 670.0e2, 1700.0e2};
  static const double scalelen_v[4] = {1.0, 3.0, 5.0, 9.0};
  int g, i;
  if (!p || !autotrophs || !ind) return BGC_ERR_ARG;

  ind->sp_ind = 1; ind->diat_ind = 2; ind->diaz_ind = 3; ind->phaeo_ind = 4;

  memset(p, 0, sizeof *p);
  p->parm_Fe_bioavail = 1.0;
  p->parm_o2_min = 4.0;
  p->parm_o2_min_delta = 2.0;
  p->parm_kappa_nitrif = 0.06 * DPS;
  p->parm_nitrif_par_lim = 1.0;
  p->parm_z_mort_0 = 0.1 * DPS;
  p->parm_z_mort2_0 = 0.4 * DPS;
  p->parm_labile_ratio = 0.85;
  p->parm_POMbury = 1.4;
  p->parm_BSIbury = 0.65;
  p->parm_fe_scavenge_rate0 = 3.0;
  p->parm_f_prod_sp_CaCO3 = 0.055;
  p->parm_POC_diss = 88.0e2;
  p->parm_SiO2_diss = 250.0e2;
  p->parm_CaCO3_diss = 150.0e2;
  for (i = 0; i < 4; ++i) {
    p->parm_scalelen_z[i] = scalelen_z[i];
    p->parm_scalelen_vals[i] = scalelen_v[i];
  }

  /* Host-set in the reference (BGC_parms.F90:45, never assigned there). */
  p->T0_Kelvin_BGC = 273.15;

  /* BGC_parms.F90:373,480-486: literals WITHOUT a kind suffix.  Under plain
   * `gfortran -O2` (the stated CPU baseline) they are REAL(4) constants widened
   * to REAL(8); a host built with -fdefault-real-8 overwrites these six members
   * with the exact doubles 1e-8, 3.17e-8, 1e-6, 1e9, 9, 5 before bgc_set_params. */
  p->epsC = (double)1.00e-8f;
  p->epsTinv = (double)3.17e-8f;
  p->epsnondim = (double)1.00e-6f;
  p->dust_fescav_scale = (double)1.0e9f;
  p->cks = 9.0;
  p->cksi = 5.0;

  p->lrest_po4 = p->lrest_no3 = p->lrest_sio3 = 0;   /* BGC_mod.F90:131-134: never set */

  for (g = 0; g < BGC_AUTOTROPH_CNT; ++g) {
    const AutoRow *r = &k_rows[g];
    BgcAutotroph *a = &autotrophs[g];
    memset(a, 0, sizeof *a);
    a->Nfixer = r->Nfixer;
    a->imp_calcifier = r->imp_calcifier;
    a->exp_calcifier = r->exp_calcifier;
    a->grazee_ind = r->grazee;
    a->temp_function = r->temp_function;
    a->kFe = r->kFe; a->kPO4 = r->kPO4; a->kDOP = r->kDOP; a->kNO3 = r->kNO3;
    a->kNH4 = r->kNH4; a->kSiO3 = r->kSiO3; a->Qp = r->Qp;
    a->gQfe_0 = r->gQfe_0; a->gQfe_min = r->gQfe_min;
    a->alphaPI = r->alphaPI_pd * DPS;
    a->PCref = r->PCref_pd * DPS;
    a->thetaN_max = r->thetaN_max;
    a->loss_thres = r->loss_thres; a->loss_thres2 = r->loss_thres2;
    a->temp_thres = r->temp_thres; a->temp_thresS = r->temp_thresS;
    a->temp_thresN = r->temp_thresN; a->temp_optN = r->temp_optN; a->temp_optS = r->temp_optS;
    a->mort = r->mort_pd * DPS;
    a->mort2 = r->mort2_pd * DPS;
    a->agg_rate_max = r->agg_rate_max; a->agg_rate_min = r->agg_rate_min;
    a->z_umax_0 = r->z_umax_0_pd * DPS;
    a->z_grz = r->z_grz;
    a->graze_zoo = r->graze_zoo; a->graze_poc = r->graze_poc; a->graze_doc = r->graze_doc;
    a->loss_poc = r->loss_poc; a->f_zoo_detr = r->f_zoo_detr;
  }
  return BGC_OK;
}

int bgc_default_tracer_indices(BgcIndices *ind) {
  int *slot;
  int i;
  if (!ind) return BGC_ERR_ARG;
  slot = &ind->po4_ind;   /* the 30 tracer members are contiguous ints, po4_ind first */
  for (i = 0; i < BGC_TRACER_CNT; ++i) slot[i] = i + 1;
  return BGC_OK;
}

/* BGC_mod.F90:271-321: each group gets its Chl/C/Fe slots; the silicifier
 * (kSiO3 > 0) gets diatSi, calcifiers get spCaCO3, everyone else 0. */
int bgc_init(const BgcIndices *ind, BgcAutotroph autotrophs[BGC_AUTOTROPH_CNT]) {
  int g;
  if (!ind || !autotrophs) return BGC_ERR_ARG;
  for (g = 1; g <= BGC_AUTOTROPH_CNT; ++g) {
    BgcAutotroph *a = &autotrophs[g - 1];
    if (g == ind->sp_ind) {
      a->Chl_ind = ind->spChl_ind; a->C_ind = ind->spC_ind; a->Fe_ind = ind->spFe_ind;
    } else if (g == ind->diat_ind) {
      a->Chl_ind = ind->diatChl_ind; a->C_ind = ind->diatC_ind; a->Fe_ind = ind->diatFe_ind;
    } else if (g == ind->diaz_ind) {
      a->Chl_ind = ind->diazChl_ind; a->C_ind = ind->diazC_ind; a->Fe_ind = ind->diazFe_ind;
    } else if (g == ind->phaeo_ind) {
      a->Chl_ind = ind->phaeoChl_ind; a->C_ind = ind->phaeoC_ind; a->Fe_ind = ind->phaeoFe_ind;
    } else {
      return BGC_ERR_ARG;
    }
    a->Si_ind = (a->kSiO3 > 0.0) ? ind->diatSi_ind : 0;
    a->CaCO3_ind = (a->imp_calcifier || a->exp_calcifier) ? ind->spCaCO3_ind : 0;
  }
  return BGC_OK;
}

int dms_parms_init(DmsParams *p) {
  if (!p) return BGC_ERR_ARG;
  p->k_S_p_base = 0.1 * DPS;
  p->zooC_avg = 0.3;
  p->mort = 0.0;
  p->k_conv = 1.0 * DPS;
  p->k_S_z = 0.1 * DPS;
  p->B_preexp = 0.1;
  p->B_exp = 0.5;
  p->k_S_B = 30.0 * DPS;
  p->k_bkgnd = 0.01 * DPS;
  p->j_dms_perI = 0.005 * DPS;
  p->inject_scale = 1.00;
  p->T_cryo_hi = 1.0;
  p->T_cryo_lo = -1.0;
  p->T_lo = 15.0;
  p->T_hi = 20.0;
  p->Min_cyano_frac = 0.0;
  p->Max_cyano_frac = 0.5;
  p->Min_yld = 0.2;
  p->Max_yld = 0.7;
  p->G_phaeo_S = 0.4;
  p->Sp_ref = 0.1;
  p->Stress_mult = 10.0;
  p->R = 0.137;
  p->Rs2n_diat = 0.01;
  p->Rs2n_phaeo = 0.3;
  p->Rs2n_cocco = 0.1;
  p->Rs2n_cyano = 0.0;
  p->Rs2n_eukar = 0.1;
  p->Rs2n_diaz = 0.0;
  p->f_qsw_par_DMS = 0.45;   /* DMS_parms.F90:191-192 */
  return BGC_OK;
}

int dms_default_tracer_indices(DmsIndices *ind) {
  int *slot;
  int i;
  if (!ind) return BGC_ERR_ARG;
  slot = &ind->dms_ind;
  for (i = 0; i < DMS_TRACER_CNT; ++i) slot[i] = i + 1;
  return BGC_OK;
}

int macros_parms_init(MacrosParams *p) {
  if (!p) return BGC_ERR_ARG;
  p->f_prot = 0.6;
  p->f_poly = 0.2;
  p->f_lip = 0.2;
  p->k_C_p_base = DPS * 0.1;
  p->zooC_avg = 0.3;
  p->mort = 0.0;
  p->k_prot_bac = DPS * 0.1;
  p->k_poly_bac = DPS * 0.01;
  p->k_lip_bac = DPS * 1.0;
  p->inject_scale = 1.0;
  return BGC_OK;
}

int macros_default_tracer_indices(MacrosIndices *ind) {
  int *slot;
  int i;
  if (!ind) return BGC_ERR_ARG;
  slot = &ind->prot_ind;
  for (i = 0; i < MACROS_TRACER_CNT; ++i) slot[i] = i + 1;
  return BGC_OK;
}
