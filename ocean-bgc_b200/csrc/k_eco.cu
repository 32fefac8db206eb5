// k_eco.cu — the ecosystem + sinking-particle column sweep for sm_100a.
//
// Replaces the column_loop of BGC_SourceSink (BGC_mod.F90:799-1970) together
// with init_particulate_terms (:2006-2109) and compute_particulate_terms
// (:2116-2699).  One thread owns one ocean column and walks it top to bottom;
// everything the level loop carries (PAR, the ten particle fluxes, the QA dust
// deficit, saturation-depth scan state, the column integrals) lives in that
// thread's registers.  The setup_loop clamp pass (:733-789) and the whole-array
// zero fills (:570, :625-727) are folded into the same sweep: a cell is read
// once and every output element is written exactly once.
//
// The carbonate solve of each cell has no vertical coupling and runs in the
// cell-parallel kernel of k_co3.cu; this kernel only consumes CO3 and the two
// saturation concentrations for the saturation-depth scan (:1003-1032).
#include "bgc_kernels.cuh"

namespace bgc {

__constant__ BgcTables c_eco;

cudaError_t upload_bgc_tables_eco(const BgcTables &t, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_eco, &t, sizeof(BgcTables), 0, cudaMemcpyHostToDevice, s);
}

namespace {

// BGC_parms.F90:37-40
constexpr double spd = 86400.0;
constexpr double dps = 1.0 / spd;
constexpr double yps = 1.0 / (365.0 * spd);
// BGC_parms.F90:327-339
constexpr double parm_Red_D_C_P = 117.0;
constexpr double parm_Red_D_C_O2 = parm_Red_D_C_P / 170.0;
constexpr double parm_Remin_D_C_O2 = parm_Red_D_C_P / 138.0;
constexpr double parm_Red_Fe_C = 3.0e-6;
constexpr double parm_Red_D_C_O2_diaz = parm_Red_D_C_P / 150.0;
// :371-386
constexpr double fe_scavenge_thres1 = 0.8e-3;
constexpr double fe_max_scale2 = 1200.0;
constexpr double dust_to_Fe = 0.035 / 55.847 * 1.0e9;
// :394-429
constexpr double caco3_poc_min = 0.4;
constexpr double spc_poc_fac = 0.11;
constexpr double f_graze_sp_poc_lim = 0.3;
constexpr double f_photosp_CaCO3 = 0.4;
constexpr double f_graze_CaCO3_remin = 0.33;
constexpr double f_graze_si_remin = 0.35;
constexpr double r_Nfix_photo = 1.25;
constexpr double Qn = 0.137;            // "Q", N/C
constexpr double Qp_zoo_pom = 0.00855;
constexpr double Qfe_zoo = 3.0e-6;
constexpr double gQsi_0 = 0.137;
constexpr double gQsi_max = 0.685;
constexpr double gQsi_min = 0.0457;
constexpr double QCaCO3_max = 0.4;
constexpr double denitrif_C_N = parm_Red_D_C_P / 136.0;
// :435-477
constexpr double thres_z1 = 100.0e2;
constexpr double thres_z2 = 150.0e2;
constexpr double loss_thres_zoo = 0.005;
constexpr double CaCO3_temp_thres1 = 6.0;
constexpr double CaCO3_temp_thres2 = -2.0;
constexpr double CaCO3_sp_thres = 4.0;
constexpr double f_qsw_par = 0.45;
constexpr double Tref = 30.0;
constexpr double Q_10 = 1.5;
constexpr double DOC_reminR = (1.0 / 250.0) * dps;
constexpr double DON_reminR = (1.0 / 160.0) * dps;
constexpr double DOFe_reminR = (1.0 / 160.0) * dps;
constexpr double DOP_reminR = (1.0 / 160.0) * dps;
constexpr double DONr_reminR = (1.0 / (365.0 * 2.5)) * dps;
constexpr double DOPr_reminR = (1.0 / (365.0 * 2.5)) * dps;
constexpr double DONrefract = 0.08;
constexpr double DOPrefract = 0.03;
constexpr double mpercm = 0.01;

// sinking_particle class constants, init_particulate_terms (BGC_mod.F90:2046-2069)
constexpr double POC_mass = 12.01;
constexpr double CaCO3_gamma = 0.30, CaCO3_mass = 100.09, CaCO3_rho = 0.05 * CaCO3_mass / POC_mass;
constexpr double SiO2_gamma = 0.030, SiO2_mass = 60.08, SiO2_rho = 0.05 * SiO2_mass / POC_mass;
constexpr double dust_diss0 = 20000.0, dust_gamma = 0.97, dust_mass = 1.0e9,
                 dust_rho = 0.05 * dust_mass / POC_mass;
constexpr double P_iron_gamma = 0.0;

constexpr int NA = BGC_AUTOTROPH_CNT;

#define ST2(name, val) do { if (A.d.name) A.d.name[i2] = (val); } while (0)
#define STA(name, a, val) do { if (A.d.name) A.d.name[i2 + (size_t)(a) * nLnC] = (val); } while (0)
#define STC(name, val) do { if (A.d.name) A.d.name[col] = (val); } while (0)
#define STCA(name, a, val) do { if (A.d.name) A.d.name[col + (size_t)(a) * (size_t)nC] = (val); } while (0)

#define ZERO_K2(name) ST2(name, 0.0);
#define ZERO_KA(name) { STA(name, 0, 0.0); STA(name, 1, 0.0); STA(name, 2, 0.0); STA(name, 3, 0.0); }
#define ZERO_CA(name) { STCA(name, 0, 0.0); STCA(name, 1, 0.0); STCA(name, 2, 0.0); STCA(name, 3, 0.0); }
#define ZERO_C1(name) STC(name, 0.0);

template <bool DIAG>
__global__ void __launch_bounds__(128)
eco_columns_kernel(const __grid_constant__ EcoArgs A) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int nL = A.nL, nC = A.nC;
  if (col >= nC) return;
  const size_t nLnC = (size_t)nL * (size_t)nC;

  int kmax = (col < A.nColumns) ? A.kmax[col] : 0;
  if (kmax > nL) kmax = nL;
  if (kmax < 0) kmax = 0;

  const BgcParams &P = c_eco.p;
  const BgcIndices &I = c_eco.ind;
  const double epsC = P.epsC, epsTinv = P.epsTinv;
  const double T0K = P.T0_Kelvin_BGC;

  // ---- various k==1 initialisations (BGC_mod.F90:808-814, :2046-2104)
  double lat = 0.0, PAR_out = 0.0;
  double POC_s = 0.0, POC_h = 0.0, Ca_s = 0.0, Ca_h = 0.0, Si_s = 0.0, Si_h = 0.0,
         du_s = 0.0, du_h = 0.0, Fe_s = 0.0, Fe_h = 0.0, QA_dust_def = 0.0;
  if (kmax > 0) {
    lat = A.lat[col];
    const double dust_in = fmax(0.0, A.dust_flux_in[col]);
    if (dust_in != 0.0) {
      du_s = (1.0 - dust_gamma) * dust_in;
      du_h = dust_gamma * dust_in;
    }
    QA_dust_def = dust_rho * (du_s + du_h);
    PAR_out = fmax(0.0, A.sw_flux[col]);
    PAR_out = PAR_out * f_qsw_par;
  }
  const bool north = lat >= 0.0;

  // ---- column integrals / scan state (diagnostics only)
  double ZSATCALC = 0.0, ZSATARAG = 0.0, CALC_ANOM_km1 = 0.0, ARAG_ANOM_km1 = 0.0;
  double zmid_km1 = 0.0, zbot_km1 = 0.0;
  double tot_bSi_form = 0.0, tot_CaCO3_form_zint = 0.0, photoC_TOT_zint = 0.0,
         photoC_NO3_TOT_zint = 0.0, Chl_TOT_zint_100m = 0.0;
  double CaCO3_form_zint[NA] = {0.0, 0.0, 0.0, 0.0}, photoC_zint[NA] = {0.0, 0.0, 0.0, 0.0},
         photoC_NO3_zint[NA] = {0.0, 0.0, 0.0, 0.0};
  double JC = 0.0, JC100 = 0.0, JN = 0.0, JN100 = 0.0, JP = 0.0, JP100 = 0.0, JSi = 0.0, JSi100 = 0.0;
  double O2_min = 0.0, O2_min_depth = 0.0;
  unsigned poc_errors = 0;

  const double *trc = A.tracers + col;
  double *tnd = A.tend + col;

  for (int k = 0; k < nL; ++k) {
    const size_t i2 = (size_t)col + (size_t)nC * (size_t)k;
    const size_t o2 = (size_t)nC * (size_t)k;   // offset of level k within one tracer slab

    if (k >= kmax) {
      // ---- inactive cell: the reference's whole-array zero fills
#pragma unroll
      for (int n = 0; n < BGC_TRACER_CNT; ++n) tnd[o2 + (size_t)n * nLnC] = 0.0;
      if (DIAG) {
        BGC_DIAG_K2_LIST(ZERO_K2)
        BGC_DIAG_KA_LIST(ZERO_KA)
      }
      continue;
    }

#define TR(ind_) fmax(0.0, trc[o2 + (size_t)((ind_) - 1) * nLnC])
#define TEND(ind_) tnd[o2 + (size_t)((ind_) - 1) * nLnC]

    // ---- this level's inputs (setup_loop clamp folded in, :747-783)
    const double TEMP = A.T[i2];
    const double zmid = A.zmid[i2];
    const double dz = A.dz[i2];
    const double zbot = A.zbot[i2];
    const double PO4_loc = TR(I.po4_ind), NO3_loc = TR(I.no3_ind), SiO3_loc = TR(I.sio3_ind),
                 NH4_loc = TR(I.nh4_ind), Fe_loc = TR(I.fe_ind), O2_loc = TR(I.o2_ind),
                 DOC_loc = TR(I.doc_ind), DON_loc = TR(I.don_ind), DOFe_loc = TR(I.dofe_ind),
                 DOP_loc = TR(I.dop_ind), DOPr_loc = TR(I.dopr_ind), DONr_loc = TR(I.donr_ind),
                 zooC_loc = TR(I.zooC_ind);
    double aChl[NA], aC[NA], aFe[NA], aSi[NA], aCaCO3[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      aChl[a] = TR(at.Chl_ind);
      aC[a] = TR(at.C_ind);
      aFe[a] = TR(at.Fe_ind);
      aSi[a] = (at.Si_ind > 0) ? TR(at.Si_ind) : 0.0;
      aCaCO3[a] = (at.CaCO3_ind > 0) ? TR(at.CaCO3_ind) : 0.0;
    }

    // ---- zero mask (:826-844)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      bool zero_mask = aChl[a] == 0.0 || aC[a] == 0.0 || aFe[a] == 0.0;
      if (at.Si_ind > 0) zero_mask = zero_mask || aSi[a] == 0.0;
      if (zero_mask) {
        aChl[a] = 0.0; aC[a] = 0.0; aFe[a] = 0.0; aSi[a] = 0.0; aCaCO3[a] = 0.0;
      }
    }

    // ---- incoming quotas and growth quotas (:850-898)
    double thetaC[NA], Qfe[NA], Qsi[NA], gQfe[NA], gQsi[NA], QCaCO3[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      const double Cden = aC[a] + epsC;
      thetaC[a] = aChl[a] / Cden;
      Qfe[a] = aFe[a] / Cden;
      Qsi[a] = 0.0; gQsi[a] = 0.0; QCaCO3[a] = 0.0;
      if (at.Si_ind > 0) Qsi[a] = fmin(aSi[a] / Cden, gQsi_max);

      gQfe[a] = at.gQfe_0;
      if (Fe_loc < P.cks * at.kFe) {
        gQfe[a] = fmax((gQfe[a] * Fe_loc / (P.cks * at.kFe)), at.gQfe_min);
      }
      if (at.Si_ind > 0) {
        double g = gQsi_0;
        if ((Fe_loc < P.cksi * at.kFe) && (Fe_loc > 0.0) && (SiO3_loc > (P.cksi * at.kSiO3))) {
          g = fmin((g * P.cksi * at.kFe / Fe_loc), gQsi_max);
        }
        if (Fe_loc == 0.0) g = gQsi_max;
        if (SiO3_loc < (P.cksi * at.kSiO3)) {
          g = fmax((g * SiO3_loc / (P.cksi * at.kSiO3)), gQsi_min);
        }
        gQsi[a] = g;
      }
      if (at.CaCO3_ind > 0) {
        QCaCO3[a] = aCaCO3[a] / Cden;
        if (QCaCO3[a] > QCaCO3_max) QCaCO3[a] = QCaCO3_max;
      }
    }

    // ---- PAR (Morel & Maritorena 2001), :907-924
    const double PAR_in = PAR_out;
    double KPARdz;
    {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < NA; ++a) s = s + aChl[a];
      const double w = fmax(s, 0.02);
      if (w < 0.13224) KPARdz = 0.000919 * pow(w, 0.3536);
      else             KPARdz = 0.001131 * pow(w, 0.4562);
    }
    KPARdz = KPARdz * dz;
    const double eKPAR = exp(-KPARdz);
    PAR_out = PAR_in * eKPAR;
    const double PAR_avg = PAR_in * (1.0 - eKPAR) / KPARdz;

    // ---- saturation-depth scan (:1003-1032); CO3 & saturation values come from k_co3
    if (DIAG) {
      const double CO3 = A.co3[i2], sat_c = A.sat_calc[i2], sat_a = A.sat_arag[i2];
      if (k == 0) {
        ZSATCALC = (CO3 > sat_c) ? -1.0 : 0.0;
        ZSATARAG = (CO3 > sat_a) ? -1.0 : 0.0;
      } else {
        const double w4 = zmid_km1 + (zmid - zmid_km1);
        if (ZSATCALC == -1.0 && CO3 <= sat_c)
          ZSATCALC = w4 * CALC_ANOM_km1 / (CALC_ANOM_km1 - (CO3 - sat_c));
        if (ZSATARAG == -1.0 && CO3 <= sat_a)
          ZSATARAG = w4 * ARAG_ANOM_km1 / (ARAG_ANOM_km1 - (CO3 - sat_a));
        if (ZSATCALC == -1.0 && k == kmax - 1) ZSATCALC = zbot;
        if (ZSATARAG == -1.0 && k == kmax - 1) ZSATARAG = zbot;
      }
      CALC_ANOM_km1 = CO3 - sat_c;
      ARAG_ANOM_km1 = CO3 - sat_a;
    }

    // ---- temperature function, loss thresholds (:1041-1094)
    const double Tfunc = pow(Q_10, ((TEMP + T0K) - (Tref + T0K)) / 10.0);

    double f_loss_thres;
    if (zmid > thres_z1) {
      if (zmid < thres_z2) f_loss_thres = (thres_z2 - zmid) / (thres_z2 - thres_z1);
      else                 f_loss_thres = 0.0;
    } else {
      f_loss_thres = 1.0;
    }

    double Pprime[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      double C_loss_thres = f_loss_thres * at.loss_thres;
      if (at.temp_function == BGC_TFNC_Q10) {
        if (TEMP < at.temp_thres) C_loss_thres = f_loss_thres * at.loss_thres2;
      } else if (at.temp_function == BGC_TFNC_QUASI_MMRT) {
        const double tmpTmax = north ? at.temp_thresN : at.temp_thresS;
        if (TEMP > tmpTmax) C_loss_thres = f_loss_thres * at.loss_thres2;
      }
      Pprime[a] = fmax(aC[a] - C_loss_thres, 0.0);
    }

    // ---- per functional group: uptake, photosynthesis, losses, grazing, routing (:1107-1388)
    double NO3_V[NA], NH4_V[NA], PO4_V[NA], DOP_V[NA], auto_graze[NA], auto_graze_zoo[NA],
           auto_graze_poc[NA], auto_graze_doc[NA], auto_graze_dic[NA], auto_loss[NA],
           auto_loss_poc[NA], auto_loss_doc[NA], auto_loss_dic[NA], auto_agg[NA], photoC[NA],
           photoFe[NA], photoSi[NA], CaCO3_PROD[NA], photoacc[NA], Nfix[NA], Nexcrete[NA],
           remaining_P_dop[NA], remaining_P_dip[NA], photoC_NO3[NA];
    double tot_CaCO3_form = 0.0, tot_Nfix = 0.0;

#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];

      const double rNO3 = NO3_loc / at.kNO3, rNH4 = NH4_loc / at.kNH4;
      const double VNO3 = rNO3 / (1.0 + rNO3 + rNH4);
      const double VNH4 = rNH4 / (1.0 + rNO3 + rNH4);
      double VNtot = VNO3 + VNH4;
      if (at.Nfixer) VNtot = 1.0;

      const double VFe = Fe_loc / (Fe_loc + at.kFe);
      double f_nut = fmin(VNtot, VFe);

      const double rPO4 = PO4_loc / at.kPO4, rDOP = DOP_loc / at.kDOP;
      const double VPO4 = rPO4 / (1.0 + rPO4 + rDOP);
      const double VDOP = rDOP / (1.0 + rPO4 + rDOP);
      const double VPtot = VPO4 + VDOP;
      f_nut = fmin(f_nut, VPtot);

      double VSiO3 = 0.0;
      if (at.kSiO3 > 0.0) {
        VSiO3 = SiO3_loc / (SiO3_loc + at.kSiO3);
        f_nut = fmin(f_nut, VSiO3);
      }
      if (DIAG) {
        STA(diag_N_lim, a, VNtot);
        STA(diag_Fe_lim, a, VFe);
        STA(diag_P_lim, a, VPtot);
        STA(diag_SiO3_lim, a, VSiO3);
      }

      double PCmax = at.PCref * f_nut * Tfunc;
      if (TEMP < at.temp_thres) PCmax = 0.0;
      if (at.temp_function == BGC_TFNC_QUASI_MMRT) {
        const double tmpTopt = north ? at.temp_optN : at.temp_optS;
        const double tmpTmax = north ? at.temp_thresN : at.temp_thresS;
        PCmax = PCmax * fmin(1.0, ((tmpTmax - TEMP) / (tmpTmax - tmpTopt)));
        if (TEMP > tmpTmax) PCmax = 0.0;
      }

      const double light_lim =
          (1.0 - exp((-1.0 * at.alphaPI * thetaC[a] * PAR_avg) / (PCmax + epsTinv)));
      const double PCphoto = PCmax * light_lim;
      if (DIAG) STA(diag_light_lim, a, light_lim);

      photoC[a] = PCphoto * aC[a];

      double VNC;
      if (VNtot > 0.0) {
        NO3_V[a] = (VNO3 / VNtot) * photoC[a] * Qn;
        NH4_V[a] = (VNH4 / VNtot) * photoC[a] * Qn;
        VNC = PCphoto * Qn;
        photoC_NO3[a] = (VNO3 / VNtot) * photoC[a];
      } else {
        NO3_V[a] = 0.0; NH4_V[a] = 0.0; VNC = 0.0; photoC_NO3[a] = 0.0;
      }
      if (VPtot > 0.0) {
        PO4_V[a] = (VPO4 / VPtot) * photoC[a] * at.Qp;
        DOP_V[a] = (VDOP / VPtot) * photoC[a] * at.Qp;
      } else {
        PO4_V[a] = 0.0; DOP_V[a] = 0.0;
      }
      photoFe[a] = photoC[a] * gQfe[a];

      photoSi[a] = 0.0;
      if (at.Si_ind > 0) {
        photoSi[a] = photoC[a] * gQsi[a];
        tot_bSi_form = tot_bSi_form + photoSi[a];   // (:1230-1231, no dz)
      }
      if (DIAG) {
        STA(diag_photoNO3, a, NO3_V[a]);
        STA(diag_photoNH4, a, NH4_V[a]);
        STA(diag_PO4_uptake, a, PO4_V[a]);
        STA(diag_DOP_uptake, a, DOP_V[a]);
        STA(diag_photoFe, a, photoFe[a]);
        STA(diag_bSi_form, a, photoSi[a]);
      }

      // Chl synthesis, GD98 (:1240-1246)
      {
        const double w = at.alphaPI * thetaC[a] * PAR_avg;
        if (w > 0.0) {
          const double pChl = at.thetaN_max * PCphoto / w;
          photoacc[a] = (pChl * VNC / thetaC[a]) * aChl[a];
        } else {
          photoacc[a] = 0.0;
        }
      }

      // implicit calcification (:1255-1278)
      CaCO3_PROD[a] = 0.0;
      if (at.imp_calcifier) {
        double cp = P.parm_f_prod_sp_CaCO3 * photoC[a];
        cp = cp * f_nut;
        if (TEMP < CaCO3_temp_thres1)
          cp = cp * fmax((TEMP - CaCO3_temp_thres2), 0.0) / (CaCO3_temp_thres1 - CaCO3_temp_thres2);
        if (aC[a] > CaCO3_sp_thres)
          cp = fmin((cp * aC[a] / CaCO3_sp_thres), (f_photosp_CaCO3 * photoC[a]));
        CaCO3_PROD[a] = cp;
        tot_CaCO3_form = tot_CaCO3_form + cp;
        if (DIAG) {
          const double w = dz * cp;
          CaCO3_form_zint[a] = CaCO3_form_zint[a] + w;
          tot_CaCO3_form_zint = tot_CaCO3_form_zint + w;
        }
      }
      if (DIAG) STA(diag_CaCO3_form, a, CaCO3_PROD[a]);

      // losses and aggregation (:1285-1290)
      auto_loss[a] = at.mort * Pprime[a] * Tfunc;
      auto_agg[a] = fmin((at.agg_rate_max * dps) * Pprime[a], at.mort2 * Pprime[a] * Pprime[a]);
      auto_agg[a] = fmax((at.agg_rate_min * dps) * Pprime[a], auto_agg[a]);

      // grazing (:1297-1324)
      double grazee_C = 0.0;
#pragma unroll
      for (int b = 0; b < NA; ++b)
        if (c_eco.same_grazee[a][b]) grazee_C = grazee_C + Pprime[b];

      double z_umax = at.z_umax_0 * Tfunc;
      if (a + 1 == I.diat_ind) {
        if (north && (TEMP > at.temp_optN)) {
          z_umax = z_umax * fmax((at.temp_thresN - TEMP) / (at.temp_thresN - at.temp_optN), 0.95);
        } else if ((lat <= 0.0) && (TEMP > at.temp_optS)) {
          z_umax = z_umax * fmax((at.temp_thresS - TEMP) / (at.temp_thresS - at.temp_optS), 0.95);
        }
      }
      if (grazee_C > 0.0) {
        auto_graze[a] = (Pprime[a] / grazee_C) * z_umax * zooC_loc * (grazee_C / (grazee_C + at.z_grz));
      } else {
        auto_graze[a] = 0.0;
      }

      // N fixation (:1331-1338)
      Nfix[a] = 0.0; Nexcrete[a] = 0.0;
      if (at.Nfixer) {
        const double w = photoC[a] * Qn;
        Nfix[a] = (w * r_Nfix_photo) - NO3_V[a] - NH4_V[a];
        Nexcrete[a] = Nfix[a] + NO3_V[a] + NH4_V[a] - w;
        tot_Nfix = tot_Nfix + Nfix[a];
      }
      if (DIAG) STA(diag_Nfix, a, Nfix[a]);

      // routing (:1354-1372)
      auto_graze_zoo[a] = at.graze_zoo * auto_graze[a];
      if (at.imp_calcifier) {
        auto_graze_poc[a] = auto_graze[a] * fmax((caco3_poc_min * QCaCO3[a]),
                                                 fmin(spc_poc_fac * fmax(1.0, Pprime[a]), f_graze_sp_poc_lim));
      } else {
        auto_graze_poc[a] = at.graze_poc * auto_graze[a];
      }
      auto_graze_doc[a] = at.graze_doc * auto_graze[a];
      auto_graze_dic[a] = auto_graze[a] - (auto_graze_zoo[a] + auto_graze_poc[a] + auto_graze_doc[a]);

      if (at.imp_calcifier) auto_loss_poc[a] = QCaCO3[a] * auto_loss[a];
      else                  auto_loss_poc[a] = at.loss_poc * auto_loss[a];
      auto_loss_doc[a] = (1.0 - P.parm_labile_ratio) * (auto_loss[a] - auto_loss_poc[a]);
      auto_loss_dic[a] = P.parm_labile_ratio * (auto_loss[a] - auto_loss_poc[a]);

      // P routing for groups whose Qp differs from Qp_zoo_pom (:1380-1386)
      remaining_P_dop[a] = 0.0; remaining_P_dip[a] = 0.0;
      if (at.Qp != Qp_zoo_pom) {
        const double remaining_P = ((auto_graze[a] + auto_loss[a] + auto_agg[a]) * at.Qp)
                                 - ((auto_graze_zoo[a]) * Qp_zoo_pom)
                                 - ((auto_graze_poc[a] + auto_loss_poc[a] + auto_agg[a]) * Qp_zoo_pom);
        remaining_P_dop[a] = (1.0 - P.parm_labile_ratio) * remaining_P;
        remaining_P_dip[a] = P.parm_labile_ratio * remaining_P;
      }
    }

    // sequential sums over the functional groups, in the reference's SUM order
#define SUM4(x) ((((0.0 + x[0]) + x[1]) + x[2]) + x[3])
    const double s_auto_loss_doc = SUM4(auto_loss_doc), s_auto_graze_doc = SUM4(auto_graze_doc),
                 s_auto_graze_poc = SUM4(auto_graze_poc), s_auto_agg = SUM4(auto_agg),
                 s_auto_loss_poc = SUM4(auto_loss_poc), s_NO3_V = SUM4(NO3_V), s_NH4_V = SUM4(NH4_V),
                 s_auto_loss_dic = SUM4(auto_loss_dic), s_auto_graze_dic = SUM4(auto_graze_dic),
                 s_photoFe = SUM4(photoFe), s_PO4_V = SUM4(PO4_V), s_auto_graze_zoo = SUM4(auto_graze_zoo),
                 s_DOP_V = SUM4(DOP_V), s_photoC = SUM4(photoC);

    // ---- zooplankton routing (:1395-1415)
    double f_zoo_detr;
    {
      double w1 = 0.0, w2 = 0.0;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        w1 = w1 + c_eco.a[a].f_zoo_detr * (auto_graze[a] + epsC * epsTinv);
        w2 = w2 + (auto_graze[a] + epsC * epsTinv);
      }
      f_zoo_detr = w1 / w2;
    }
    const double Zprime = fmax(zooC_loc - f_loss_thres * loss_thres_zoo, 0.0);
    const double zoo_loss = (P.parm_z_mort2_0 * pow(Zprime, 1.5) + P.parm_z_mort_0 * Zprime) * Tfunc;
    const double zoo_loss_doc = (1.0 - P.parm_labile_ratio) * (1.0 - f_zoo_detr) * zoo_loss;
    const double zoo_loss_dic = P.parm_labile_ratio * (1.0 - f_zoo_detr) * zoo_loss;

    // ---- DOM (:1421-1461)
    const double DOC_prod = zoo_loss_doc + s_auto_loss_doc + s_auto_graze_doc;
    const double DON_prod = Qn * DOC_prod;
    double DOP_prod = Qp_zoo_pom * zoo_loss_doc;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      if (at.Qp == Qp_zoo_pom) DOP_prod = DOP_prod + at.Qp * (auto_loss_doc[a] + auto_graze_doc[a]);
      else                     DOP_prod = DOP_prod + remaining_P_dop[a];
    }
    double DOFe_prod = Qfe_zoo * zoo_loss_doc;
#pragma unroll
    for (int a = 0; a < NA; ++a) DOFe_prod = DOFe_prod + Qfe[a] * (auto_loss_doc[a] + auto_graze_doc[a]);

    double DOC_remin = DOC_loc * DOC_reminR;
    double DON_remin = DON_loc * DON_reminR;
    double DOFe_remin = DOFe_loc * DOFe_reminR;
    double DOP_remin = DOP_loc * DOP_reminR;
    double DONr_remin, DOPr_remin;
    if (PAR_avg > 1.0) {
      DONr_remin = DONr_loc * DONr_reminR;
      DOPr_remin = DOPr_loc * DOPr_reminR;
    } else {
      DONr_remin = DONr_loc * (1.0 / (365.0 * 670.0)) * dps;
      DOPr_remin = DOPr_loc * (1.0 / (365.0 * 460.0)) * dps;
      DOC_remin = DOC_remin * 0.0685;
      DON_remin = DON_remin * 0.1;
      DOFe_remin = DOFe_remin * 0.05;
      DOP_remin = DOP_remin * 0.05;
    }

    // ---- particle production (:1467-1529)
    const double POC_prod = f_zoo_detr * zoo_loss + s_auto_graze_poc + s_auto_agg + s_auto_loss_poc;
    double Ca_prod = 0.0, Si_prod = 0.0;   // last writer wins among qualifying groups (:1480-1498)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (c_eco.a[a].CaCO3_ind > 0)
        Ca_prod = ((1.0 - f_graze_CaCO3_remin) * auto_graze[a] + auto_loss[a] + auto_agg[a]) * QCaCO3[a];
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (c_eco.a[a].Si_ind > 0)
        Si_prod = Qsi[a] * ((1.0 - f_graze_si_remin) * auto_graze[a] + auto_agg[a] +
                            c_eco.a[a].loss_poc * auto_loss[a]);
    }

    double Fe_scavenge_rate = P.parm_fe_scavenge_rate0;
    Fe_scavenge_rate = Fe_scavenge_rate *
        ((POC_s + POC_h) * 120.1 +
         (Ca_s + Ca_h) * CaCO3_mass +
         (Si_s + Si_h) * SiO2_mass +
         (du_s + du_h) * P.dust_fescav_scale);
    if (Fe_loc > fe_scavenge_thres1)
      Fe_scavenge_rate = Fe_scavenge_rate + (Fe_loc - fe_scavenge_thres1) * fe_max_scale2;
    const double Fe_scavenge = yps * Fe_loc * Fe_scavenge_rate;

    double Fe_prod = (zoo_loss * f_zoo_detr * Qfe_zoo) + Fe_scavenge;
#pragma unroll
    for (int a = 0; a < NA; ++a)
      Fe_prod = Fe_prod + Qfe[a] * (auto_agg[a] + auto_graze_poc[a] + auto_loss_poc[a]);

    // =====================================================================
    // compute_particulate_terms (BGC_mod.F90:2116-2699) for this level
    // =====================================================================
    const double Ca_s_in = Ca_s, Ca_h_in = Ca_h, Si_s_in = Si_s, Si_h_in = Si_h,
                 du_s_in = du_s, du_h_in = du_h, POC_s_in = POC_s, POC_h_in = POC_h,
                 Fe_s_in = Fe_s, Fe_h_in = Fe_h;
    double POC_sed = 0.0, Ca_sed = 0.0, Si_sed = 0.0, du_sed = 0.0, Fe_sed = 0.0;
    double SED_DENITRIF = 0.0, OTHER_REMIN = 0.0;
    double POC_remin, Ca_remin, Si_remin, du_remin, Fe_remin;
    {
      double scalelength;   // piecewise-linear in zbot, :2273-2286
      if (zbot < P.parm_scalelen_z[0]) {
        scalelength = P.parm_scalelen_vals[0];
      } else if (zbot >= P.parm_scalelen_z[3]) {
        scalelength = P.parm_scalelen_vals[3];
      } else {
        scalelength = 0.0;
#pragma unroll
        for (int n = 3; n >= 1; --n) {   // first n (ascending) with zbot < z[n]  <=>  last assignment descending
          if (zbot < P.parm_scalelen_z[n])
            scalelength = P.parm_scalelen_vals[n - 1] +
                          (P.parm_scalelen_vals[n] - P.parm_scalelen_vals[n - 1]) *
                              (zbot - P.parm_scalelen_z[n - 1]) /
                              (P.parm_scalelen_z[n] - P.parm_scalelen_z[n - 1]);
        }
      }

      const double DECAY_Hard = exp(-dz / 4.0e6);
      const double DECAY_HardDust = exp(-dz / 1.2e7);
      const double TfuncS = Tfunc;   // 1.5**(same exponent) (:2295) is bit-identical to Tfunc (:1041)

      const double dzr = 1.0 / dz;

      double poc_diss = P.parm_POC_diss;
      if ((O2_loc >= 5.0) && (O2_loc < 40.0)) {
        poc_diss = P.parm_POC_diss * (1.0 + (3.3 - 1.0) * (40.0 - O2_loc) / 35.0);
      } else if (O2_loc < 5.0) {
        poc_diss = P.parm_POC_diss * 3.3;
      }
      poc_diss = scalelength * poc_diss;
      double sio2_diss = scalelength * P.parm_SiO2_diss;
      const double caco3_diss = scalelength * P.parm_CaCO3_diss;
      const double dust_diss = scalelength * dust_diss0;
      sio2_diss = sio2_diss / TfuncS;

      const double decay_POC_E = exp(-dz / poc_diss);
      const double decay_SiO2 = exp(-dz / sio2_diss);
      const double decay_CaCO3 = exp(-dz / caco3_diss);
      const double decay_dust = exp(-dz / dust_diss);

      Ca_s = Ca_s_in * decay_CaCO3 + Ca_prod * ((1.0 - CaCO3_gamma) * (1.0 - decay_CaCO3) * caco3_diss);
      Ca_h = Ca_h_in * DECAY_Hard + Ca_prod * (CaCO3_gamma * dz);
      Si_s = Si_s_in * decay_SiO2 + Si_prod * ((1.0 - SiO2_gamma) * (1.0 - decay_SiO2) * sio2_diss);
      Si_h = Si_h_in * DECAY_Hard + Si_prod * (SiO2_gamma * dz);
      du_s = du_s_in * decay_dust;
      du_h = du_h_in * DECAY_HardDust;

      double POC_PROD_avail = POC_prod - CaCO3_rho * Ca_prod - SiO2_rho * Si_prod;
      if (POC_PROD_avail < 0.0) poc_errors++;   // computed and never reported by the reference (:2381-2383)

      double new_QA_dust_def;
      if (QA_dust_def > 0.0) {
        new_QA_dust_def = QA_dust_def * (du_s + du_h) / (du_s_in + du_h_in);
      } else {
        new_QA_dust_def = 0.0;
      }
      if (new_QA_dust_def > 0.0) {
        new_QA_dust_def = new_QA_dust_def - POC_PROD_avail * dz;
        if (new_QA_dust_def < 0.0) {
          POC_PROD_avail = -new_QA_dust_def * dzr;
          new_QA_dust_def = 0.0;
        } else {
          POC_PROD_avail = 0.0;
        }
      }
      QA_dust_def = new_QA_dust_def;

      if (POC_h_in == 0.0 && POC_prod == 0.0) {
        POC_h = 0.0;
      } else {
        POC_h = CaCO3_rho * (Ca_s + Ca_h) + SiO2_rho * (Si_s + Si_h) + dust_rho * (du_s + du_h) -
                new_QA_dust_def;
        POC_h = fmax(POC_h, 0.0);
      }
      POC_s = POC_s_in * decay_POC_E + POC_PROD_avail * ((1.0 - decay_POC_E) * poc_diss);

      Ca_remin = Ca_prod + ((Ca_s_in - Ca_s) + (Ca_h_in - Ca_h)) * dzr;
      Si_remin = Si_prod + ((Si_s_in - Si_s) + (Si_h_in - Si_h)) * dzr;
      POC_remin = POC_prod + ((POC_s_in - POC_s) + (POC_h_in - POC_h)) * dzr;
      du_remin = ((du_s_in - du_s) + (du_h_in - du_h)) * dzr;

      if (POC_s_in + POC_h_in == 0.0) {
        Fe_remin = (POC_remin * parm_Red_Fe_C);
      } else {
        Fe_remin = (POC_remin * (Fe_s_in + Fe_h_in) / (POC_s_in + POC_h_in));
      }
      Fe_remin = Fe_remin + (Fe_s_in * 1.5e-5);
      Fe_s = Fe_s_in + dz * ((1.0 - P_iron_gamma) * Fe_prod - Fe_remin);
      if (Fe_s < 0.0) {
        Fe_s = 0.0;
        Fe_remin = Fe_s_in * dzr + (1.0 - P_iron_gamma) * Fe_prod;
      }
      Fe_remin = Fe_remin + du_remin * dust_to_Fe + (A.fesedflux[i2] * dzr);
      Fe_h = Fe_h_in;

      if (k == kmax - 1) {   // bottom cell: burial, sediment denitrification (:2522-2631)
        double flux = POC_s + POC_h;
        if (flux > 0.0) {
          double flux_alt = flux * mpercm * spd;
          POC_sed = flux * fmin(0.8, P.parm_POMbury *
                                         (0.013 + 0.53 * flux_alt * flux_alt /
                                                      ((7.0 + flux_alt) * (7.0 + flux_alt))));
          SED_DENITRIF = dzr * flux * (0.06 + 0.19 * pow(0.99, (O2_loc - NO3_loc)));
          if (NO3_loc < 5.0) SED_DENITRIF = 0.0;
          flux_alt = flux * 1.0e-6 * spd * 365.0;
          OTHER_REMIN = dzr * fmin(fmin(0.1 + flux_alt, 0.5) * (flux - POC_sed),
                                   (flux - POC_sed - (SED_DENITRIF * dz * denitrif_C_N)));
          if (O2_loc < 1.0) OTHER_REMIN = dzr * (flux - POC_sed - (SED_DENITRIF * dz * denitrif_C_N));
        }

        flux = Si_s + Si_h;
        {
          const double flux_alt = flux * mpercm * spd;
          Si_sed = (flux_alt > 2.0) ? 0.2 : 0.04;
          Si_sed = flux * P.parm_BSIbury * Si_sed;
        }
        if (zbot < 3300.0e2) Ca_sed = Ca_s + Ca_h;

        flux = Ca_s + Ca_h;
        if (flux > 0.0) Ca_remin = Ca_remin + ((flux - Ca_sed) * dzr);
        flux = Si_s + Si_h;
        if (flux > 0.0) Si_remin = Si_remin + ((flux - Si_sed) * dzr);
        flux = POC_s + POC_h;
        if (flux > 0.0) POC_remin = POC_remin + ((flux - POC_sed) * dzr);

        flux = (Fe_s + Fe_h);
        if (flux > 0.0) Fe_sed = flux;
        du_sed = du_s + du_h;

        Ca_s = 0.0; Ca_h = 0.0; Si_s = 0.0; Si_h = 0.0; du_s = 0.0; du_h = 0.0;
        POC_s = 0.0; POC_h = 0.0; Fe_s = 0.0; Fe_h = 0.0;
      }

      if (DIAG) {   // :2637-2694
        ST2(diag_POC_FLUX_IN, POC_s_in + POC_h_in);
        ST2(diag_POC_PROD, POC_prod);
        ST2(diag_POC_REMIN, POC_remin);
        ST2(diag_CaCO3_FLUX_IN, Ca_s_in + Ca_h_in);
        ST2(diag_CaCO3_PROD, Ca_prod);
        ST2(diag_CaCO3_REMIN, Ca_remin);
        ST2(diag_SiO2_FLUX_IN, Si_s_in + Si_h_in);
        ST2(diag_SiO2_PROD, Si_prod);
        ST2(diag_SiO2_REMIN, Si_remin);
        ST2(diag_dust_FLUX_IN, du_s_in + du_h_in);
        ST2(diag_dust_REMIN, du_remin);
        ST2(diag_P_iron_FLUX_IN, Fe_s_in + Fe_h_in);
        ST2(diag_P_iron_PROD, Fe_prod);
        ST2(diag_P_iron_REMIN, Fe_remin);
        ST2(diag_calcToSed, Ca_sed);
        ST2(diag_bsiToSed, Si_sed);
        ST2(diag_pocToSed, POC_sed);
        ST2(diag_SedDenitrif, SED_DENITRIF * dz);
        ST2(diag_OtherRemin, OTHER_REMIN * dz);
        ST2(diag_ponToSed, (POC_sed * Qn));
        ST2(diag_popToSed, (POC_sed * Qp_zoo_pom));
        ST2(diag_dustToSed, du_sed);
        ST2(diag_pfeToSed, Fe_sed);
      }
    }

    // ---- nitrification / denitrification (:1545-1577)
    double RESTORE_NO3 = 0.0, RESTORE_SiO3 = 0.0, RESTORE_PO4 = 0.0;
    if (P.lrest_no3) RESTORE_NO3 = A.rtau[i2] * (A.no3_clim[i2] - NO3_loc);
    if (P.lrest_sio3) RESTORE_SiO3 = A.rtau[i2] * (A.sio3_clim[i2] - SiO3_loc);
    if (P.lrest_po4) RESTORE_PO4 = A.rtau[i2] * (A.po4_clim[i2] - PO4_loc);

    double NITRIF;
    if (PAR_out < P.parm_nitrif_par_lim) {
      NITRIF = P.parm_kappa_nitrif * NH4_loc;
      if (PAR_in > P.parm_nitrif_par_lim)
        NITRIF = NITRIF * log(PAR_out / P.parm_nitrif_par_lim) / (-KPARdz);
    } else {
      NITRIF = 0.0;
    }

    double DENITRIF;
    {
      double w = ((P.parm_o2_min + P.parm_o2_min_delta) - O2_loc) / P.parm_o2_min_delta;
      w = fmin(fmax(w, 0.0), 1.0);
      if (NO3_loc == 0.0) w = 0.0;
      DENITRIF = w * ((DOC_remin + POC_remin - OTHER_REMIN) / denitrif_C_N - SED_DENITRIF);
    }

    // ---- tendencies (:1583-1790)
    const double t_no3 = RESTORE_NO3 + NITRIF - DENITRIF - SED_DENITRIF - s_NO3_V;

    double t_nh4 = -s_NH4_V - NITRIF + DON_remin + DONr_remin +
                   Qn * (zoo_loss_dic + s_auto_loss_dic + s_auto_graze_dic + POC_remin * (1.0 - DONrefract));
#pragma unroll
    for (int a = 0; a < NA; ++a)
      if (c_eco.a[a].Nfixer) t_nh4 = t_nh4 + Nexcrete[a];

    double t_fe = Fe_remin + (Qfe_zoo * zoo_loss_dic) + DOFe_remin - s_photoFe - Fe_scavenge;
#pragma unroll
    for (int a = 0; a < NA; ++a)
      t_fe = t_fe + (Qfe[a] * (auto_loss_dic[a] + auto_graze_dic[a])) + auto_graze_zoo[a] * (Qfe[a] - Qfe_zoo);

    double t_sio3 = RESTORE_SiO3 + Si_remin;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (c_eco.a[a].Si_ind > 0)
        t_sio3 = t_sio3 - photoSi[a] +
                 Qsi[a] * (f_graze_si_remin * auto_graze[a] + (1.0 - c_eco.a[a].loss_poc) * auto_loss[a]);
    }

    double t_po4 = RESTORE_PO4 + DOP_remin + DOPr_remin - s_PO4_V +
                   Qp_zoo_pom * ((1.0 - DOPrefract) * POC_remin + zoo_loss_dic);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      if (at.Qp == Qp_zoo_pom) t_po4 = t_po4 + at.Qp * (auto_loss_dic[a] + auto_graze_dic[a]);
      else                     t_po4 = t_po4 + remaining_P_dip[a];
    }

    double t_autoC[NA], t_autoSi[NA], t_autoCaCO3[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      const double w = auto_graze[a] + auto_loss[a] + auto_agg[a];
      t_autoC[a] = photoC[a] - w;
      t_autoSi[a] = photoSi[a] - Qsi[a] * w;
      t_autoCaCO3[a] = CaCO3_PROD[a] - QCaCO3[a] * w;
      TEND(at.C_ind) = t_autoC[a];
      TEND(at.Chl_ind) = photoacc[a] - thetaC[a] * w;
      TEND(at.Fe_ind) = photoFe[a] - Qfe[a] * w;
      if (at.Si_ind > 0) TEND(at.Si_ind) = t_autoSi[a];
      if (at.CaCO3_ind > 0) TEND(at.CaCO3_ind) = t_autoCaCO3[a];
    }

    const double t_zooC = s_auto_graze_zoo - zoo_loss;
    const double t_doc = DOC_prod - DOC_remin;
    const double t_don = (DON_prod * (1.0 - DONrefract)) - DON_remin;
    const double t_donr = (DON_prod * DONrefract) - DONr_remin + (POC_remin * DONrefract * Qn);
    const double t_dop = (DOP_prod * (1.0 - DOPrefract)) - DOP_remin - s_DOP_V;
    const double t_dopr = (DOP_prod * DOPrefract) - DOPr_remin + (POC_remin * DOPrefract * Qp_zoo_pom);
    const double t_dofe = DOFe_prod - DOFe_remin;

    double t_dic = s_auto_loss_dic + s_auto_graze_dic - s_photoC + DOC_remin + POC_remin + zoo_loss_dic + Ca_remin;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (c_eco.a[a].CaCO3_ind > 0)
        t_dic = t_dic + f_graze_CaCO3_remin * auto_graze[a] * QCaCO3[a] - CaCO3_PROD[a];
    }
    const double t_dic_alt = A.alt_co2_use_eco ? t_dic : 0.0;

    double t_alk = -t_no3 + t_nh4 + 2.0 * Ca_remin;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (c_eco.a[a].CaCO3_ind > 0)
        t_alk = t_alk + 2.0 * (f_graze_CaCO3_remin * auto_graze[a] * QCaCO3[a] - CaCO3_PROD[a]);
    }

    double O2_PRODUCTION = 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (photoC[a] > 0.0) {
        if (!c_eco.a[a].Nfixer) {
          const double den = NO3_V[a] + NH4_V[a];
          O2_PRODUCTION = O2_PRODUCTION + photoC[a] *
              ((NO3_V[a] / den) / parm_Red_D_C_O2 + (NH4_V[a] / den) / parm_Remin_D_C_O2);
        } else {
          const double den = NO3_V[a] + NH4_V[a] + Nfix[a];
          O2_PRODUCTION = O2_PRODUCTION + photoC[a] *
              ((NO3_V[a] / den) / parm_Red_D_C_O2 + (NH4_V[a] / den) / parm_Remin_D_C_O2 +
               (Nfix[a] / den) / parm_Red_D_C_O2_diaz);
        }
      }
    }
    double O2_CONSUMPTION;
    {
      double w = (O2_loc - P.parm_o2_min) / P.parm_o2_min_delta;
      w = fmin(fmax(w, 0.0), 1.0);
      O2_CONSUMPTION = w * ((POC_remin + DOC_remin - (SED_DENITRIF * denitrif_C_N) - OTHER_REMIN +
                             zoo_loss_dic + s_auto_loss_dic + s_auto_graze_dic) / parm_Remin_D_C_O2 +
                            (2.0 * NITRIF));
    }
    const double t_o2 = O2_PRODUCTION - O2_CONSUMPTION;

    TEND(I.no3_ind) = t_no3;
    TEND(I.nh4_ind) = t_nh4;
    TEND(I.fe_ind) = t_fe;
    TEND(I.sio3_ind) = t_sio3;
    TEND(I.po4_ind) = t_po4;
    TEND(I.zooC_ind) = t_zooC;
    TEND(I.doc_ind) = t_doc;
    TEND(I.don_ind) = t_don;
    TEND(I.donr_ind) = t_donr;
    TEND(I.dop_ind) = t_dop;
    TEND(I.dopr_ind) = t_dopr;
    TEND(I.dofe_ind) = t_dofe;
    TEND(I.dic_ind) = t_dic;
    TEND(I.dic_alt_co2_ind) = t_dic_alt;
    TEND(I.alk_ind) = t_alk;
    TEND(I.o2_ind) = t_o2;

    // ---- diagnostics and column integrals (:1796-1945)
    if (DIAG) {
      ST2(diag_tot_Nfix, tot_Nfix);
      ST2(diag_tot_CaCO3_form, tot_CaCO3_form);
      ST2(diag_NO3_RESTORE, RESTORE_NO3);
      ST2(diag_SiO3_RESTORE, RESTORE_SiO3);
      ST2(diag_PO4_RESTORE, RESTORE_PO4);
      ST2(diag_NITRIF, NITRIF);
      ST2(diag_DENITRIF, DENITRIF);
      ST2(diag_O2_PRODUCTION, O2_PRODUCTION);
      ST2(diag_O2_CONSUMPTION, O2_CONSUMPTION);
      if (A.d.diag_AOU) {   // O2SAT_singleValue, Garcia & Gordon 1992 (:3012-3083)
        const double SALT = A.S[i2];
        const double TS = log(((T0K + 25.0) - TEMP) / (T0K + TEMP));
        double o2sat = exp(2.00907 + TS * (3.22014 + TS * (4.05010 + TS * (4.94457 + TS * (-2.56847E-1 + TS * 3.88767)))) +
                           SALT * ((-6.24523E-3 + TS * (-7.37614E-3 + TS * (-1.03410E-2 + TS * -8.17083E-3))) +
                                   SALT * -4.88682E-7));
        o2sat = o2sat / 0.0223916;
        A.d.diag_AOU[i2] = o2sat - O2_loc;
      }
      ST2(diag_PAR_avg, PAR_avg);
      ST2(diag_zoo_loss, zoo_loss);
      ST2(diag_auto_graze_TOT, SUM4(auto_graze));
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        STA(diag_auto_graze, a, auto_graze[a]);
        STA(diag_auto_loss, a, auto_loss[a]);
        STA(diag_auto_agg, a, auto_agg[a]);
        STA(diag_photoC, a, photoC[a]);
        photoC_zint[a] = photoC_zint[a] + dz * photoC[a];
      }
      ST2(diag_photoC_TOT, s_photoC);
      photoC_TOT_zint = photoC_TOT_zint + s_photoC * dz;

      double photoC_NO3_TOT = 0.0;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        STA(diag_photoC_NO3, a, photoC_NO3[a]);
        photoC_NO3_zint[a] = photoC_NO3_zint[a] + photoC_NO3[a] * dz;
        photoC_NO3_TOT = photoC_NO3_TOT + photoC_NO3[a];
        // adds the RUNNING per-group integral every level (:1844-1846)
        photoC_NO3_TOT_zint = photoC_NO3_TOT_zint + photoC_NO3_zint[a];
      }
      ST2(diag_photoC_NO3_TOT, photoC_NO3_TOT);

      ST2(diag_DOC_prod, DOC_prod);
      ST2(diag_DOC_remin, DOC_remin);
      ST2(diag_DON_prod, DON_prod);
      ST2(diag_DON_remin, DON_remin);
      ST2(diag_DOP_prod, DOP_prod);
      ST2(diag_DOP_remin, DOP_remin);
      ST2(diag_DOFe_prod, DOFe_prod);
      ST2(diag_DOFe_remin, DOFe_remin);
      ST2(diag_Fe_scavenge, Fe_scavenge);
      ST2(diag_Fe_scavenge_rate, Fe_scavenge_rate);

      const double ztop = (k > 0) ? zbot_km1 : 0.0;
      const double w2 = fmin(100.0e2 - ztop, dz);
      const double pt100 = (w2 > 0.0) ? w2 : 0.0;
      const bool shallow = zbot <= 100.0e2;

      const double s_tC = SUM4(t_autoC);
      double w1 = t_dic + t_doc + t_zooC + s_tC;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (c_eco.a[a].CaCO3_ind > 0) w1 = w1 + t_autoCaCO3[a];
      JC = JC + w1 * dz + POC_sed + Ca_sed;
      JC100 = JC100 + w1 * pt100 + (shallow ? (POC_sed + Ca_sed) : 0.0);

      w1 = t_no3 + t_nh4 + t_don + t_donr + Qn * t_zooC + Qn * s_tC;
      w1 = w1 + DENITRIF + SED_DENITRIF;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (c_eco.a[a].Nfixer) w1 = w1 - Nfix[a];
      JN = JN + w1 * dz + POC_sed * Qn;
      JN100 = JN100 + w1 * pt100 + (shallow ? (POC_sed * Qn) : 0.0);

      w1 = t_po4 + t_dop + t_dopr + Qp_zoo_pom * t_zooC;
#pragma unroll
      for (int a = 0; a < NA; ++a) w1 = w1 + c_eco.a[a].Qp * t_autoC[a];
      JP = JP + w1 * dz + POC_sed * Qp_zoo_pom;
      JP100 = JP100 + w1 * pt100 + (shallow ? (POC_sed * Qp_zoo_pom) : 0.0);

      w1 = t_sio3;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (c_eco.a[a].Si_ind > 0) w1 = w1 + t_autoSi[a];
      JSi = JSi + w1 * dz + Si_sed;
      JSi100 = JSi100 + w1 * pt100 + (shallow ? Si_sed : 0.0);

#pragma unroll
      for (int a = 0; a < NA; ++a) Chl_TOT_zint_100m = Chl_TOT_zint_100m + aChl[a] * pt100;

      // O2 minimum scan (:1954-1968)
      if (k == 0 || O2_loc < O2_min) { O2_min = O2_loc; O2_min_depth = zmid; }

      zmid_km1 = zmid;
      zbot_km1 = zbot;
    }
#undef TR
#undef TEND
  }   // level loop

  // ---- per-column diagnostics
  if (DIAG) {
    if (kmax > 0) {
      STC(diag_photoC_TOT_zint, photoC_TOT_zint);
      STC(diag_photoC_NO3_TOT_zint, photoC_NO3_TOT_zint);
      STC(diag_Jint_Ctot, JC);       STC(diag_Jint_100m_Ctot, JC100);
      STC(diag_Jint_Ntot, JN);       STC(diag_Jint_100m_Ntot, JN100);
      STC(diag_Jint_Ptot, JP);       STC(diag_Jint_100m_Ptot, JP100);
      STC(diag_Jint_Sitot, JSi);     STC(diag_Jint_100m_Sitot, JSi100);
      STC(diag_Chl_TOT_zint_100m, Chl_TOT_zint_100m);
      STC(diag_tot_CaCO3_form_zint, tot_CaCO3_form_zint);
      STC(diag_tot_bSi_form, tot_bSi_form);
      STC(diag_zsatcalc, ZSATCALC);
      STC(diag_zsatarag, ZSATARAG);
      STC(diag_O2_ZMIN, O2_min);
      STC(diag_O2_ZMIN_DEPTH, O2_min_depth);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        STCA(diag_photoC_zint, a, photoC_zint[a]);
        STCA(diag_photoC_NO3_zint, a, photoC_NO3_zint[a]);
        STCA(diag_CaCO3_form_zint, a, CaCO3_form_zint[a]);
      }
    } else {
      BGC_DIAG_C1_LIST(ZERO_C1)
      BGC_DIAG_CA_LIST(ZERO_CA)
    }
  }
  if (poc_errors && A.status) atomicAdd(&A.status[2], (unsigned long long)poc_errors);
}

}  // namespace

cudaError_t launch_eco_columns(const EcoArgs &a, bool any_diag, cudaStream_t s) {
  if (a.nC <= 0 || a.nL <= 0) return cudaSuccess;
  const int block = 128;
  const int grid = (a.nC + block - 1) / block;
  if (any_diag) eco_columns_kernel<true><<<grid, block, 0, s>>>(a);
  else          eco_columns_kernel<false><<<grid, block, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace bgc
