/* parms_oracle.c — restatement of BGC_parms_init / DMS_parms_init /
 * MACROS_parms_init and the index wiring of BGC_init.
 * TEST INFRASTRUCTURE ONLY (see bgc_oracle.h).  Pinned against the translated reference. */
#include "bgc_oracle.h"
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* BGC_parms.F90:37-40 */
static const double spd = 86400.0;
#define DPS (1.0 / spd)

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* BGC_parms.F90:385-386 */
double oracle_dust_to_Fe(void) { return 0.035 / 55.847 * 1.0e9; }

/* BGC_parms.F90:497-699.  `default_real_8` selects how the reference was
 * compiled: 0 = plain `gfortran -O2` (north_star's baseline): the un-suffixed
 * literals at BGC_parms.F90:373,480-486 are REAL(4) values widened to REAL(8);
 * 1 = `-fdefault-real-8` (MPAS's usual flags): they are exact doubles. */
void oracle_BGC_parms_init(BgcParams *p, BgcAutotroph a[4], BgcIndices *ind,
                           int default_real_8) {
  const double dps = DPS;
  int i;
  memset(p, 0, sizeof(*p));
  memset(a, 0, 4 * sizeof(a[0]));

  ind->sp_ind = 1;    /* :515-518 */
  ind->diat_ind = 2;
  ind->diaz_ind = 3;
  ind->phaeo_ind = 4;

  p->parm_Fe_bioavail = 1.0;            /* :524-538 */
  p->parm_o2_min = 4.0;
  p->parm_o2_min_delta = 2.0;
  p->parm_kappa_nitrif = 0.06 * dps;
  p->parm_nitrif_par_lim = 1.0;
  p->parm_z_mort_0 = 0.1 * dps;
  p->parm_z_mort2_0 = 0.4 * dps;
  p->parm_labile_ratio = 0.85;
  p->parm_POMbury = 1.4;
  p->parm_BSIbury = 0.65;
  p->parm_fe_scavenge_rate0 = 3.0;
  p->parm_f_prod_sp_CaCO3 = 0.055;
  p->parm_POC_diss = 88.0e2;
  p->parm_SiO2_diss = 250.0e2;
  p->parm_CaCO3_diss = 150.0e2;
  {
    const double z[4] = {130.0e2, 290.0e2, 670.0e2, 1700.0e2};   /* :540-541 */
    const double v[4] = {1.0, 3.0, 5.0, 9.0};
    for (i = 0; i < 4; ++i) { p->parm_scalelen_z[i] = z[i]; p->parm_scalelen_vals[i] = v[i]; }
  }
  p->T0_Kelvin_BGC = 273.15;            /* :45 — host-set; never assigned in the reference */
  if (default_real_8) {                 /* :373, :480-486 */
    p->epsC = 1.00e-8; p->epsTinv = 3.17e-8; p->epsnondim = 1.00e-6;
  } else {
    p->epsC = (double)1.00e-8f; p->epsTinv = (double)3.17e-8f; p->epsnondim = (double)1.00e-6f;
  }
  p->dust_fescav_scale = (double)1.0e9f;  /* exactly representable */
  p->cks = (double)9.f;
  p->cksi = (double)5.f;
  p->lrest_po4 = p->lrest_no3 = p->lrest_sio3 = 0;   /* BGC_mod.F90:131-134 */

  /* sp :543-580 */
  i = ind->sp_ind - 1;
  a[i].Nfixer = 0; a[i].imp_calcifier = 1; a[i].exp_calcifier = 0;
  a[i].grazee_ind = ind->sp_ind;
  a[i].kFe = 0.04e-3; a[i].kPO4 = 0.01; a[i].kDOP = 0.26; a[i].kNO3 = 0.1;
  a[i].kNH4 = 0.01; a[i].kSiO3 = 0.0; a[i].Qp = 0.00855; a[i].gQfe_0 = 20.0e-6;
  a[i].gQfe_min = 3.0e-6; a[i].alphaPI = 0.6 * dps; a[i].PCref = 5.5 * dps;
  a[i].thetaN_max = 2.5; a[i].loss_thres = 0.04; a[i].loss_thres2 = 0.0;
  a[i].temp_thres = -20.0; a[i].temp_thresN = -20.0; a[i].temp_thresS = -20.0;
  a[i].temp_function = BGC_TFNC_Q10; a[i].temp_optN = 50.0; a[i].temp_optS = 50.0;
  a[i].mort = 0.12 * dps; a[i].mort2 = 0.001 * dps; a[i].agg_rate_max = 0.9;
  a[i].agg_rate_min = 0.01; a[i].z_umax_0 = 3.3 * dps; a[i].z_grz = 1.05;
  a[i].graze_zoo = 0.3; a[i].graze_poc = 0.0; a[i].graze_doc = 0.15;
  a[i].loss_poc = 0.0; a[i].f_zoo_detr = 0.15;

  /* diat :582-619 */
  i = ind->diat_ind - 1;
  a[i].Nfixer = 0; a[i].imp_calcifier = 0; a[i].exp_calcifier = 0;
  a[i].grazee_ind = ind->diat_ind;
  a[i].kFe = 0.06e-3; a[i].kPO4 = 0.05; a[i].kDOP = 0.9; a[i].kNO3 = 0.5;
  a[i].kNH4 = 0.05; a[i].kSiO3 = 0.8; a[i].Qp = 0.00855; a[i].gQfe_0 = 20.0e-6;
  a[i].gQfe_min = 3.0e-6; a[i].alphaPI = 0.465 * dps; a[i].PCref = 5.5 * dps;
  a[i].thetaN_max = 4.0; a[i].loss_thres = 0.04; a[i].loss_thres2 = 0.0;
  a[i].temp_thres = -20.0; a[i].temp_thresN = 35.0; a[i].temp_thresS = 10.0;
  a[i].temp_function = BGC_TFNC_Q10; a[i].temp_optN = 16.3; a[i].temp_optS = 5.0;
  a[i].mort = 0.12 * dps; a[i].mort2 = 0.001 * dps; a[i].agg_rate_max = 0.9;
  a[i].agg_rate_min = 0.02; a[i].z_umax_0 = 3.23 * dps; a[i].z_grz = 1.0;
  a[i].graze_zoo = 0.3; a[i].graze_poc = 0.42; a[i].graze_doc = 0.15;
  a[i].loss_poc = 0.0; a[i].f_zoo_detr = 0.2;

  /* diaz :621-658 */
  i = ind->diaz_ind - 1;
  a[i].Nfixer = 1; a[i].imp_calcifier = 0; a[i].exp_calcifier = 0;
  a[i].grazee_ind = ind->diaz_ind;
  a[i].kFe = 0.04e-3; a[i].kPO4 = 0.02; a[i].kDOP = 0.09; a[i].kNO3 = 1.0;
  a[i].kNH4 = 0.15; a[i].kSiO3 = 0.0; a[i].Qp = 0.002735; a[i].gQfe_0 = 60.0e-6;
  a[i].gQfe_min = 12.0e-6; a[i].alphaPI = 0.4 * dps; a[i].PCref = 0.7 * dps;
  a[i].thetaN_max = 2.5; a[i].loss_thres = 0.022; a[i].loss_thres2 = 0.001;
  a[i].temp_thres = 14.0; a[i].temp_thresN = -20.0; a[i].temp_thresS = -20.0;
  a[i].temp_function = BGC_TFNC_Q10; a[i].temp_optN = 50.0; a[i].temp_optS = 50.0;
  a[i].mort = 0.15 * dps; a[i].mort2 = 0.0; a[i].agg_rate_max = 0.0;
  a[i].agg_rate_min = 0.0; a[i].z_umax_0 = 0.6 * dps; a[i].z_grz = 1.2;
  a[i].graze_zoo = 0.3; a[i].graze_poc = 0.05; a[i].graze_doc = 0.15;
  a[i].loss_poc = 0.0; a[i].f_zoo_detr = 0.15;

  /* phaeo :660-697 */
  i = ind->phaeo_ind - 1;
  a[i].Nfixer = 0; a[i].imp_calcifier = 0; a[i].exp_calcifier = 0;
  a[i].grazee_ind = ind->diat_ind;
  a[i].kFe = 0.075e-3; a[i].kPO4 = 0.05; a[i].kDOP = 0.9; a[i].kNO3 = 0.7;
  a[i].kNH4 = 0.05; a[i].kSiO3 = 0.0; a[i].Qp = 0.00855; a[i].gQfe_0 = 20.0e-6;
  a[i].gQfe_min = 3.0e-6; a[i].alphaPI = 0.77 * dps; a[i].PCref = 5.5 * dps;
  a[i].thetaN_max = 2.5; a[i].loss_thres = 0.04; a[i].loss_thres2 = 0.0;
  a[i].temp_thres = -20.0; a[i].temp_thresN = 35.0; a[i].temp_thresS = 10.0;
  a[i].temp_function = BGC_TFNC_QUASI_MMRT; a[i].temp_optN = 16.3; a[i].temp_optS = 5.0;
  a[i].mort = 0.12 * dps; a[i].mort2 = 0.001 * dps; a[i].agg_rate_max = 0.9;
  a[i].agg_rate_min = 0.02; a[i].z_umax_0 = 3.23 * dps; a[i].z_grz = 1.0;
  a[i].graze_zoo = 0.3; a[i].graze_poc = 0.42; a[i].graze_doc = 0.15;
  a[i].loss_poc = 0.0; a[i].f_zoo_detr = 0.2;
}

/* BGC_mod.F90:271-321 — only the index wiring; the name strings stay in the
 * Fortran shim. */
void oracle_BGC_init(const BgcIndices *ind, BgcAutotroph a[4]) {
  int auto_ind;
  for (auto_ind = 1; auto_ind <= 4; ++auto_ind) {
    BgcAutotroph *at = &a[auto_ind - 1];
    int Chl_ind = 0, C_ind = 0, Fe_ind = 0;
    if (auto_ind == ind->sp_ind) {
      Chl_ind = ind->spChl_ind; C_ind = ind->spC_ind; Fe_ind = ind->spFe_ind;
    } else if (auto_ind == ind->diat_ind) {
      Chl_ind = ind->diatChl_ind; C_ind = ind->diatC_ind; Fe_ind = ind->diatFe_ind;
    } else if (auto_ind == ind->diaz_ind) {
      Chl_ind = ind->diazChl_ind; C_ind = ind->diazC_ind; Fe_ind = ind->diazFe_ind;
    } else if (auto_ind == ind->phaeo_ind) {
      Chl_ind = ind->phaeoChl_ind; C_ind = ind->phaeoC_ind; Fe_ind = ind->phaeoFe_ind;
    }
    at->Chl_ind = Chl_ind; at->C_ind = C_ind; at->Fe_ind = Fe_ind;
    at->Si_ind = (at->kSiO3 > 0.0) ? ind->diatSi_ind : 0;                         /* :303-310 */
    at->CaCO3_ind = (at->imp_calcifier || at->exp_calcifier) ? ind->spCaCO3_ind : 0; /* :312-320 */
  }
}

/* DMS_parms.F90:203-241 (+ :191-192) */
void oracle_DMS_parms_init(DmsParams *p) {
  const double dps = DPS;
  p->k_S_p_base = 0.1 * dps; p->zooC_avg = 0.3; p->mort = 0.0; p->k_conv = 1.0 * dps;
  p->k_S_z = 0.1 * dps; p->B_preexp = 0.1; p->B_exp = 0.5; p->k_S_B = 30.0 * dps;
  p->k_bkgnd = 0.01 * dps; p->j_dms_perI = 0.005 * dps; p->inject_scale = 1.00;
  p->T_cryo_hi = 1.0; p->T_cryo_lo = -1.0; p->T_lo = 15.0; p->T_hi = 20.0;
  p->Min_cyano_frac = 0.0; p->Max_cyano_frac = 0.5; p->Min_yld = 0.2; p->Max_yld = 0.7;
  p->G_phaeo_S = 0.4; p->Sp_ref = 0.1; p->Stress_mult = 10.0; p->R = 0.137;
  p->Rs2n_diat = 0.01; p->Rs2n_phaeo = 0.3; p->Rs2n_cocco = 0.1; p->Rs2n_cyano = 0.0;
  p->Rs2n_eukar = 0.1; p->Rs2n_diaz = 0.0;
  p->f_qsw_par_DMS = 0.45;
}

/* MACROS_parms.F90:143-162 */
void oracle_MACROS_parms_init(MacrosParams *p) {
  const double dps = DPS;
  p->f_prot = 0.6; p->f_poly = 0.2; p->f_lip = 0.2; p->k_C_p_base = dps * 0.1;
  p->zooC_avg = 0.3; p->mort = 0.0; p->k_prot_bac = dps * 0.1; p->k_poly_bac = dps * 0.01;
  p->k_lip_bac = dps * 1.0; p->inject_scale = 1.0;
}
