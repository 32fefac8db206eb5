/* bgc_b200.h — C ABI of the B200-native Ocean-BGC column hot path.
 *
 * Every entry point replaces one public procedure of the reference Fortran
 * library (E3SM-Project/Ocean-BGC, libBGC.a).  A thin ISO_C_BINDING shim
 * (ocean-bgc_b200/fortran/) keeps the Fortran module / procedure names and
 * derived types unchanged and forwards to these functions; see INTEGRATION.md.
 *
 *   bgc_parms_init        <- BGC_parms_init        BGC_parms.F90:497-699
 *   bgc_init              <- BGC_init (index wiring only) BGC_mod.F90:184-333
 *   bgc_source_sink       <- BGC_SourceSink        BGC_mod.F90:340-1998
 *                            (+ init/compute_particulate_terms :2006-2699,
 *                             comp_CO3terms / comp_co3_sat_vals co2calc.F90:214,1096)
 *   bgc_surface_fluxes    <- BGC_SurfaceFluxes     BGC_mod.F90:2706-2957
 *   bgc_co2calc_points    <- co2calc_1point        co2calc.F90:75-210 (batched)
 *   dms_parms_init        <- DMS_parms_init        DMS_parms.F90:203-241
 *   dms_source_sink       <- DMS_SourceSink        DMS_mod.F90:156-770
 *   dms_surface_fluxes    <- DMS_SurfaceFluxes     DMS_mod.F90:778-908
 *   macros_parms_init     <- MACROS_parms_init     MACROS_parms.F90:143-162
 *   macros_source_sink    <- MACROS_SourceSink     MACROS_mod.F90:137-411
 *
 * Conventions
 *   - plain C: pointers, ints, doubles; no C++ / torch types cross the boundary.
 *   - every function returns BGC_OK (0) or a negative BGC_ERR_* code; nothing
 *     throws.  bgc_last_error() returns a static message for the last failure
 *     on the calling thread.
 *   - tracer / autotroph indices are 1-based exactly as the Fortran host
 *     chooses them (BGC_parms.F90:82-118).
 *   - two memory spaces (argument `mem_space`):
 *       BGC_MEM_HOST_FORTRAN  host pointers, reference layout: level fastest,
 *                             A(k,col[,n]) at  k + nLevelsMax*(col + nColumnsMax*n)
 *                             surface/flux arrays F(col[,n]) at col + nColumnsMax*n.
 *                             The call is synchronous (Fortran semantics).
 *       BGC_MEM_DEVICE_SOA    device pointers, column fastest:
 *                             A(k,col[,n]) at  col + nColumnsMax*(k + nLevelsMax*n)
 *                             F(col[,n])   at  col + nColumnsMax*n   (unchanged).
 *                             The call is stream-ordered on the ctx stream and
 *                             returns without synchronising.  An even nColumnsMax and
 *                             16-byte aligned arrays let the column sweep stage its
 *                             inputs with bulk (TMA) copies; otherwise it falls back to
 *                             per-thread loads (same results up to the last bit of a few
 *                             tendencies, about half the speed).
 *   - the caller owns every array; the library never frees caller memory.
 */
#ifndef BGC_B200_H
#define BGC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGC_TRACER_CNT     30   /* BGC_mod.F90:117-118   */
#define DMS_TRACER_CNT     14   /* DMS_mod.F90:61-62     */
#define MACROS_TRACER_CNT   8   /* MACROS_mod.F90:60-61  */
#define BGC_AUTOTROPH_CNT   4   /* BGC_parms.F90:42-43   */

#define BGC_TFNC_Q10        1   /* BGC_parms.F90:447-449 */
#define BGC_TFNC_QUASI_MMRT 2

enum {
  BGC_OK = 0,
  BGC_ERR_ARG = -1,         /* null / inconsistent argument                     */
  BGC_ERR_CUDA = -2,        /* a CUDA runtime call failed (see bgc_last_error)  */
  BGC_ERR_NO_DEVICE = -3,   /* no CUDA device: there is NO CPU fallback         */
  BGC_ERR_PARAMS = -4,      /* *_set_params not called before a compute call    */
  BGC_ERR_NCCL = -5
};

enum { BGC_MEM_HOST_FORTRAN = 0, BGC_MEM_DEVICE_SOA = 1 };

/* ------------------------------------------------------------------ parameters */

/* Run-time tunables of module BGC_parms (BGC_parms.F90:346-365) plus the few
 * compile-time constants whose value depends on how the reference is compiled
 * (single-precision literals, BGC_parms.F90:373,480-486) and the host-set
 * T0_Kelvin_BGC (BGC_parms.F90:45). */
typedef struct BgcParams {
  double parm_Fe_bioavail, parm_o2_min, parm_o2_min_delta, parm_kappa_nitrif,
         parm_nitrif_par_lim, parm_z_mort_0, parm_z_mort2_0, parm_labile_ratio,
         parm_POMbury, parm_BSIbury, parm_fe_scavenge_rate0, parm_f_prod_sp_CaCO3,
         parm_POC_diss, parm_SiO2_diss, parm_CaCO3_diss;
  double parm_scalelen_z[4], parm_scalelen_vals[4];
  double T0_Kelvin_BGC;
  double epsC, epsTinv, epsnondim, dust_fescav_scale, cks, cksi;
  int    lrest_po4, lrest_no3, lrest_sio3;   /* BGC_mod.F90:131-134 (never set => 0) */
  int    reserved;
} BgcParams;

/* autotroph_type minus its two CHARACTER(256) names (BGC_parms.F90:51-79). */
typedef struct BgcAutotroph {
  int    Nfixer, imp_calcifier, exp_calcifier;
  int    grazee_ind, temp_function, Chl_ind, C_ind, Fe_ind, Si_ind, CaCO3_ind;
  double kFe, kPO4, kDOP, kNO3, kNH4, kSiO3, Qp, gQfe_0, gQfe_min, alphaPI, PCref,
         thetaN_max, loss_thres, loss_thres2, temp_thres, temp_thresS, temp_thresN,
         temp_optN, temp_optS, mort, mort2, agg_rate_max, agg_rate_min, z_umax_0,
         z_grz, graze_zoo, graze_poc, graze_doc, loss_poc, f_zoo_detr;
} BgcAutotroph;

/* BGC_indices_type minus the name strings (BGC_parms.F90:81-125). 1-based. */
typedef struct BgcIndices {
  int po4_ind, no3_ind, sio3_ind, nh4_ind, fe_ind, o2_ind, dic_ind, dic_alt_co2_ind,
      alk_ind, doc_ind, don_ind, dofe_ind, dop_ind, dopr_ind, donr_ind, zooC_ind,
      spC_ind, spChl_ind, spFe_ind, spCaCO3_ind, diatC_ind, diatChl_ind, diatFe_ind,
      diatSi_ind, phaeoC_ind, phaeoChl_ind, phaeoFe_ind, diazC_ind, diazChl_ind,
      diazFe_ind;
  int sp_ind, diat_ind, diaz_ind, phaeo_ind;
} BgcIndices;

/* DMS_parms.F90:160-192 */
typedef struct DmsParams {
  double k_S_p_base, zooC_avg, mort, k_conv, k_S_z, B_preexp, B_exp, k_S_B, k_bkgnd,
         j_dms_perI, inject_scale, T_cryo_hi, T_cryo_lo, T_lo, T_hi, Min_cyano_frac,
         Max_cyano_frac, Min_yld, Max_yld, G_phaeo_S, Sp_ref, Stress_mult, R,
         Rs2n_diat, Rs2n_phaeo, Rs2n_cocco, Rs2n_cyano, Rs2n_eukar, Rs2n_diaz,
         f_qsw_par_DMS;
} DmsParams;

/* DMS_parms.F90:62-77 */
typedef struct DmsIndices {
  int dms_ind, dmsp_ind, no3_ind, doc_ind, zooC_ind, spC_ind, spCaCO3_ind, diatC_ind,
      diazC_ind, phaeoC_ind, spChl_ind, diatChl_ind, diazChl_ind, phaeoChl_ind;
} DmsIndices;

/* MACROS_parms.F90:122-132 */
typedef struct MacrosParams {
  double f_prot, f_poly, f_lip, k_C_p_base, zooC_avg, mort, k_prot_bac, k_poly_bac,
         k_lip_bac, inject_scale;
} MacrosParams;

/* MACROS_parms.F90:62-71 */
typedef struct MacrosIndices {
  int prot_ind, poly_ind, lip_ind, zooC_ind, spC_ind, diatC_ind, diazC_ind, phaeoC_ind;
} MacrosIndices;

/* ------------------------------------------------------------ argument blocks */

/* BGC_input_type (BGC_parms.F90:127-137) */
typedef struct BgcInput {
  const double *BGC_tracers;                 /* (nLevelsMax, nColumnsMax, 30) */
  const double *PotentialTemperature, *Salinity, *cell_center_depth,
               *cell_thickness, *cell_bottom_depth;   /* (nLevelsMax, nColumnsMax) */
  const double *cell_latitude;               /* (nColumnsMax) */
  const int    *number_of_active_levels;     /* (nColumnsMax) */
} BgcInput;

/* BGC_forcing_type (BGC_parms.F90:139-165).  intent(inout) in BGC_SurfaceFluxes. */
typedef struct BgcForcing {
  double *FESEDFLUX, *NUTR_RESTORE_RTAU, *NO3_CLIM, *PO4_CLIM, *SiO3_CLIM; /* (k,col) */
  double *dust_FLUX_IN, *ShortWaveFlux_surface, *surfacePressure, *iceFraction,
         *windSpeedSquared10m, *atmCO2, *atmCO2_ALT_CO2, *surface_pH,
         *surface_pH_alt_co2, *surfaceDepth, *SST, *SSS;                    /* (col)  */
  double *depositionFlux, *riverFlux, *gasFlux, *seaIceFlux, *netFlux;     /* (col,30) */
  int lcalc_O2_gas_flux, lcalc_CO2_gas_flux;
} BgcForcing;

/* BGC_output_type (BGC_parms.F90:167-172) */
typedef struct BgcOutput {
  double *BGC_tendencies;                    /* (nLevelsMax, nColumnsMax, 30) */
  double *PH_PREV_3D, *PH_PREV_ALT_CO2_3D;   /* (nLevelsMax, nColumnsMax) RMW  */
} BgcOutput;

/* BGC_flux_diagnostics_type (BGC_parms.F90:174-190), all (nColumnsMax) */
#define BGC_FLUX_DIAG_LIST(X) \
  X(pistonVel_O2) X(SCHMIDT_O2) X(O2SAT) X(xkw) X(co2star) X(dco2star) X(pco2surf) \
  X(dpco2) X(pistonVel_CO2) X(SCHMIDT_CO2) X(co2star_alt_co2) X(dco2star_alt_co2) \
  X(pco2surf_alt_co2) X(dpco2_alt_co2)

/* BGC_diagnostics_type (BGC_parms.F90:192-321), in declaration order.
 * K2 : (nLevelsMax, nColumnsMax)        61 arrays
 * KA : (nLevelsMax, nColumnsMax, 4)     18 arrays
 * CA : (nColumnsMax, 4)                  3 arrays
 * C1 : (nColumnsMax)                    17 arrays */
#define BGC_DIAG_K2_LIST(X) \
  X(diag_tot_Nfix) X(diag_O2_PRODUCTION) X(diag_O2_CONSUMPTION) X(diag_AOU) \
  X(diag_PO4_RESTORE) X(diag_NO3_RESTORE) X(diag_SiO3_RESTORE) X(diag_PAR_avg) \
  X(diag_POC_FLUX_IN) X(diag_POC_PROD) X(diag_POC_REMIN) X(diag_POC_ACCUM) \
  X(diag_CaCO3_FLUX_IN) X(diag_CaCO3_PROD) X(diag_CaCO3_REMIN) X(diag_SiO2_FLUX_IN) \
  X(diag_SiO2_PROD) X(diag_SiO2_REMIN) X(diag_dust_FLUX_IN) X(diag_dust_REMIN) \
  X(diag_P_iron_FLUX_IN) X(diag_P_iron_PROD) X(diag_P_iron_REMIN) X(diag_auto_graze_TOT) \
  X(diag_zoo_loss) X(diag_photoC_TOT) X(diag_photoC_NO3_TOT) X(diag_DOC_prod) \
  X(diag_DOC_remin) X(diag_DON_prod) X(diag_DON_remin) X(diag_DOFe_prod) \
  X(diag_DOFe_remin) X(diag_DOP_prod) X(diag_DOP_remin) X(diag_Fe_scavenge) \
  X(diag_Fe_scavenge_rate) X(diag_NITRIF) X(diag_DENITRIF) X(diag_DONr_remin) \
  X(diag_DOPr_remin) \
  X(diag_CO3) X(diag_HCO3) X(diag_H2CO3) X(diag_pH_3D) X(diag_CO3_ALT_CO2) \
  X(diag_HCO3_ALT_CO2) X(diag_H2CO3_ALT_CO2) X(diag_pH_3D_ALT_CO2) X(diag_co3_sat_calc) \
  X(diag_co3_sat_arag) X(diag_calcToSed) X(diag_pocToSed) X(diag_ponToSed) \
  X(diag_popToSed) X(diag_bsiToSed) X(diag_dustToSed) X(diag_pfeToSed) \
  X(diag_SedDenitrif) X(diag_OtherRemin) X(diag_tot_CaCO3_form)
#define BGC_DIAG_KA_LIST(X) \
  X(diag_N_lim) X(diag_P_lim) X(diag_Fe_lim) X(diag_SiO3_lim) X(diag_light_lim) \
  X(diag_photoC) X(diag_photoC_NO3) X(diag_photoFe) X(diag_photoNO3) X(diag_photoNH4) \
  X(diag_DOP_uptake) X(diag_PO4_uptake) X(diag_auto_graze) X(diag_auto_loss) \
  X(diag_auto_agg) X(diag_bSi_form) X(diag_CaCO3_form) X(diag_Nfix)
#define BGC_DIAG_CA_LIST(X) \
  X(diag_photoC_zint) X(diag_photoC_NO3_zint) X(diag_CaCO3_form_zint)
#define BGC_DIAG_C1_LIST(X) \
  X(diag_photoC_TOT_zint) X(diag_photoC_NO3_TOT_zint) X(diag_Jint_Ctot) \
  X(diag_Jint_100m_Ctot) X(diag_Jint_Ntot) X(diag_Jint_100m_Ntot) X(diag_Jint_Ptot) \
  X(diag_Jint_100m_Ptot) X(diag_Jint_Sitot) X(diag_Jint_100m_Sitot) \
  X(diag_Chl_TOT_zint_100m) X(diag_tot_CaCO3_form_zint) X(diag_tot_bSi_form) \
  X(diag_zsatcalc) X(diag_zsatarag) X(diag_O2_ZMIN) X(diag_O2_ZMIN_DEPTH)

#define BGC_DECL_PTR(name) double *name;

typedef struct BgcFluxDiagnostics { BGC_FLUX_DIAG_LIST(BGC_DECL_PTR) } BgcFluxDiagnostics;

/* A NULL member is an extension of the reference contract: that diagnostic is
 * simply not produced ("diagnostics off"). */
typedef struct BgcDiagnostics {
  BGC_DIAG_K2_LIST(BGC_DECL_PTR)
  BGC_DIAG_KA_LIST(BGC_DECL_PTR)
  BGC_DIAG_CA_LIST(BGC_DECL_PTR)
  BGC_DIAG_C1_LIST(BGC_DECL_PTR)
} BgcDiagnostics;

/* DMS_input_type / DMS_forcing_type / DMS_output_type (DMS_parms.F90:85-111) */
typedef struct DmsInput {
  const double *DMS_tracers;                 /* (nLevelsMax, nColumnsMax, 14) */
  const double *cell_thickness;              /* (nLevelsMax, nColumnsMax) */
  const int    *number_of_active_levels;
} DmsInput;
typedef struct DmsForcing {
  double *ShortWaveFlux_surface, *surfacePressure, *iceFraction, *windSpeedSquared10m,
         *SST, *SSS;                         /* (col) */
  double *netFlux;                           /* (col,14) */
  int lcalc_DMS_gas_flux;
} DmsForcing;
typedef struct DmsOutput { double *DMS_tendencies; } DmsOutput;

/* DMS_flux_diagnostics_type (DMS_parms.F90:113-123), (nColumnsMax) */
#define DMS_FLUX_DIAG_LIST(X) \
  X(diag_DMS_IFRAC) X(diag_DMS_XKW) X(diag_DMS_ATM_PRESS) X(diag_DMS_PV) \
  X(diag_DMS_SCHMIDT) X(diag_DMS_SAT) X(diag_DMS_SURF) X(diag_DMS_WS)
/* DMS_diagnostics_type (DMS_parms.F90:125-154), (nLevelsMax, nColumnsMax) */
#define DMS_DIAG_LIST(X) \
  X(diag_DMS_S_DMSP) X(diag_DMS_S_TOTAL) X(diag_DMS_R_B) X(diag_DMS_R_PHOT) \
  X(diag_DMS_R_BKGND) X(diag_DMS_R_TOTAL) X(diag_DMSP_S_PHAEO) X(diag_DMSP_S_NONPHAEO) \
  X(diag_DMSP_S_ZOO) X(diag_DMSP_S_TOTAL) X(diag_DMSP_R_B) X(diag_DMSP_R_BKGND) \
  X(diag_DMSP_R_TOTAL) X(diag_Cyano_frac) X(diag_Cocco_frac) X(diag_Eukar_frac) \
  X(diag_diatS) X(diag_diatN) X(diag_phytoN) X(diag_coccoS) X(diag_cyanoS) \
  X(diag_eukarS) X(diag_diazS) X(diag_phaeoS) X(diag_zooS) X(diag_zooCC) X(diag_RSNzoo)
typedef struct DmsFluxDiagnostics { DMS_FLUX_DIAG_LIST(BGC_DECL_PTR) } DmsFluxDiagnostics;
typedef struct DmsDiagnostics { DMS_DIAG_LIST(BGC_DECL_PTR) } DmsDiagnostics;

/* MACROS types (MACROS_parms.F90:79-113) */
typedef struct MacrosInput {
  const double *MACROS_tracers;              /* (nLevelsMax, nColumnsMax, 8) */
  const double *cell_thickness;              /* never read (MACROS_mod.F90) */
  const int    *number_of_active_levels;
} MacrosInput;
typedef struct MacrosOutput { double *MACROS_tendencies; } MacrosOutput;
#define MACROS_DIAG_LIST(X) \
  X(diag_PROT_S_TOTAL) X(diag_POLY_S_TOTAL) X(diag_LIP_S_TOTAL) X(diag_PROT_R_TOTAL) \
  X(diag_POLY_R_TOTAL) X(diag_LIP_R_TOTAL)
typedef struct MacrosDiagnostics { MACROS_DIAG_LIST(BGC_DECL_PTR) } MacrosDiagnostics;

/* Device-side status counters (the reference has no error reporting at all:
 * both solver aborts are commented out, co2calc.F90:931-933,993-995). */
typedef struct BgcStatus {
  unsigned long long no_bracket;      /* bracket growth hit the iteration cap      */
  unsigned long long no_convergence;  /* drtsafe fell through maxit=100            */
  unsigned long long poc_error;       /* POC_PROD_avail < 0 (BGC_mod.F90:2381-2383) */
  unsigned long long nonfinite;       /* a NaN/Inf tendency was produced           */
} BgcStatus;

/* Global inventory / conservation vector produced by the source-sink kernels
 * (sum over this rank's columns), reduced across ranks by bgc_inventory_allreduce. */
#define BGC_INVENTORY_LEN 64
/*   [0..29]  sum_col sum_k BGC_tendency(n) * dz          (n = tracer slot, 0-based)
 *   [30..43] same for the 14 DMS tendencies
 *   [44..51] same for the 8 MACROS tendencies
 *   [52..59] sum_col of diag_Jint_{C,100m_C,N,100m_N,P,100m_P,Si,100m_Si}tot
 *   [60]     number of active cells, [61] number of active columns, [62..63] spare */

typedef struct bgc_ctx bgc_ctx;

/* ------------------------------------------------------------------ functions */

const char *bgc_last_error(void);
const char *bgc_version(void);

/* Host-side defaults; mirror the reference initialisers. No device needed. */
int bgc_parms_init(BgcParams *p, BgcAutotroph autotrophs[BGC_AUTOTROPH_CNT], BgcIndices *ind);
int bgc_default_tracer_indices(BgcIndices *ind);  /* fills po4_ind..diazFe_ind = 1..30 in declaration order */
int bgc_init(const BgcIndices *ind, BgcAutotroph autotrophs[BGC_AUTOTROPH_CNT]);
int dms_parms_init(DmsParams *p);
int dms_default_tracer_indices(DmsIndices *ind);
int macros_parms_init(MacrosParams *p);
int macros_default_tracer_indices(MacrosIndices *ind);

/* Context = one GPU + persistent device arena sized for (nLevelsMax, nColumnsMax). */
int bgc_ctx_create(int device, int nLevelsMax, int nColumnsMax, bgc_ctx **out);
int bgc_ctx_destroy(bgc_ctx *ctx);
int bgc_ctx_set_stream(bgc_ctx *ctx, void *cuda_stream);   /* cudaStream_t; NULL = ctx-owned stream */
int bgc_ctx_synchronize(bgc_ctx *ctx);
/* bgc_source_sink (BGC_MEM_DEVICE_SOA) runs the carbonate kernel on an internal side stream
 * beside the column sweep and, by default, joins it to the ctx stream before returning, so
 * the call is stream-ordered as a whole.  With the deferred join enabled the join moves to
 * the next JOIN POINT: bgc_carbonate_join (stream-ordered, no host wait), bgc_ctx_synchronize,
 * bgc_get_status, bgc_inventory_get / bgc_inventory_allreduce, bgc_ctx_set_stream, the next
 * bgc_source_sink or any BGC_MEM_HOST_FORTRAN call.  Until then the caller must neither
 * overwrite the inputs of that bgc_source_sink call nor read PH_PREV_*, the ten carbonate
 * diagnostics and diag_zsatcalc / diag_zsatarag on the ctx stream; in exchange the FP64-bound
 * carbonate solve overlaps the HBM-bound DMS / MACROS / surface-flux kernels that follow. */
int bgc_ctx_set_deferred_join(bgc_ctx *ctx, int enable);
/* enable = 0 launches the carbonate kernel on the ctx stream after the sweep (no side stream):
 * for per-kernel timing and debugging.  Default: 1 (or the environment BGC_CONCURRENT_CO3) = side
 * stream, placement chosen by the size of the sweep: beside the sweep when it has a partial last wave
 * to fill; when the sweep is less than one wave (a GPU's slab in an 8-way split) the carbonate work is
 * confined to the SMs the sweep cannot use, the remainder following behind the sweep.  2 and 3 pin
 * "behind the sweep" and "confined" for experiments.  Results do not depend on the placement. */
int bgc_ctx_set_concurrency(bgc_ctx *ctx, int enable);
/* Zero-biomass shortcut of the column sweep (default on).  The reference zeroes a functional
 * group whose Chl, C or Fe is exactly zero in a cell (BGC_mod.F90:826-844) and then spends
 * the whole group body producing zeros.  Where that holds for all 32 columns of a warp at a
 * level the sweep skips the body and writes the zeros directly; results are bit-identical.
 * enable = 0 executes the full body everywhere (for measuring the data-independent cost). */
int bgc_ctx_set_zero_shortcut(bgc_ctx *ctx, int enable);
/* Bytes the BGC_MEM_HOST_FORTRAN calls of this ctx have copied across PCIe since it was created or the counters
 * were last reset: bytes[0] host -> device, bytes[1] device -> host.  (Outputs the reference assigns the constant
 * zero whatever the inputs are do not cross PCIe: the library zero-fills those host ranges itself.) */
int bgc_transfer_bytes(bgc_ctx *ctx, unsigned long long bytes[2], int reset);
int bgc_carbonate_join(bgc_ctx *ctx);
int bgc_get_status(bgc_ctx *ctx, BgcStatus *out, int reset);

int bgc_set_params(bgc_ctx *ctx, const BgcParams *p,
                   const BgcAutotroph autotrophs[BGC_AUTOTROPH_CNT], const BgcIndices *ind);
int dms_set_params(bgc_ctx *ctx, const DmsParams *p, const DmsIndices *ind);
int macros_set_params(bgc_ctx *ctx, const MacrosParams *p, const MacrosIndices *ind);

int bgc_source_sink(bgc_ctx *ctx, const BgcInput *in, const BgcForcing *forcing,
                    BgcOutput *out, BgcDiagnostics *diag,
                    int nLevelsMax, int nColumnsMax, int nColumns,
                    int alt_co2_use_eco, int mem_space);

int bgc_surface_fluxes(bgc_ctx *ctx, const BgcInput *in, BgcForcing *forcing,
                       BgcFluxDiagnostics *diag, int nLevelsMax,
                       int nColumnsMax, int nColumns, int mem_space);

/* Batched co2calc_1point (k = 1).  All arrays length n.  phlo/phhi are the pH
 * bracket (not modified); outputs ph, co2star, dco2star, pco2surf, dpco2. */
int bgc_co2calc_points(bgc_ctx *ctx, int n, const double *depth, const double *temp,
                       const double *salt, const double *dic, const double *ta,
                       const double *pt, const double *sit, const double *phlo,
                       const double *phhi, const double *xco2, const double *atmpres,
                       double *ph, double *co2star, double *dco2star, double *pco2surf,
                       double *dpco2, int mem_space);

/* The other two public procedures of the reference's co2calc module (co2calc.F90:24), batched:
 * comp_CO3terms (co2calc.F90:214-316) and comp_co3_sat_vals (:1096-1238), n points per call.
 * k_level[i] is the reference's 1-based level index of point i (the pressure correction of the
 * equilibrium constants is keyed on k > 1, not on depth); k_level == NULL means k_all for
 * every point.  depth in metres as in the reference.  phlo/phhi are not modified (the reference
 * declares them INOUT and never writes them: comp_htotal hands copies to the solver).  Every call
 * computes its own coefficients (the reference's lcomp_co3_coeffs = .true.; the .false. form reuses
 * module state of the previous scalar call, which a batch does not have). */
int bgc_comp_co3terms(bgc_ctx *ctx, int n, const int *k_level, int k_all, const double *depth,
                      const double *temp, const double *salt, const double *dic, const double *ta,
                      const double *pt, const double *sit, const double *phlo, const double *phhi,
                      double *ph, double *h2co3, double *hco3, double *co3, int mem_space);
int bgc_comp_co3_sat_vals(bgc_ctx *ctx, int n, const int *k_level, int k_all, const double *depth,
                          const double *temp, const double *salt, double *co3_sat_calc,
                          double *co3_sat_arag, int mem_space);

int dms_source_sink(bgc_ctx *ctx, const DmsInput *in, const DmsForcing *forcing,
                    DmsOutput *out, DmsDiagnostics *diag,
                    int nLevelsMax, int nColumnsMax, int nColumns, int mem_space);
int dms_surface_fluxes(bgc_ctx *ctx, const DmsInput *in, DmsForcing *forcing,
                       DmsFluxDiagnostics *diag, int nLevelsMax,
                       int nColumnsMax, int nColumns, int mem_space);
int macros_source_sink(bgc_ctx *ctx, const MacrosInput *in, MacrosOutput *out,
                       MacrosDiagnostics *diag,
                       int nLevelsMax, int nColumnsMax, int nColumns, int mem_space);

/* Launch accounting and optional per-kernel timing.  The library counts every
 * kernel it launches (per kernel id below); with timing enabled each launch is
 * also bracketed by CUDA events on the ctx stream and bgc_timing_get returns the
 * accumulated device time.  (The reference has no profiling hooks at all.) */
enum {
  BGC_K_CO3_CELLS = 0, BGC_K_ECO_COLUMNS = 1, BGC_K_DMS_COLUMNS = 2, BGC_K_MACROS_CELLS = 3,
  BGC_K_SURFACE_FLUXES = 4, BGC_K_DMS_SURFACE = 5, BGC_K_CO2CALC_POINTS = 6, BGC_K_INVENTORY = 7,
  BGC_K_TRANSPOSE = 8, BGC_K_ZSAT_COLUMNS = 9, BGC_K_ACCUMULATE = 10, BGC_KERNEL_ID_COUNT = 11
};
int bgc_timing_enable(bgc_ctx *ctx, int enable);
int bgc_timing_reset(bgc_ctx *ctx);   /* zeroes times AND launch counters */
int bgc_timing_get(bgc_ctx *ctx, int kernel_id, double *total_ms,
                   unsigned long long *timed_launches, unsigned long long *launches);
const char *bgc_kernel_name(int kernel_id);

/* Inventory: device vector accumulated by the *_source_sink calls since the
 * last bgc_inventory_reset.  bgc_inventory_get copies this rank's vector to the
 * host; bgc_inventory_device_ptr exposes it for a caller-side collective. */
int bgc_inventory_enable(bgc_ctx *ctx, int enable);   /* default: disabled (no extra pass) */
int bgc_inventory_reset(bgc_ctx *ctx);
int bgc_inventory_get(bgc_ctx *ctx, double out[BGC_INVENTORY_LEN]);
int bgc_inventory_device_ptr(bgc_ctx *ctx, double **dev_ptr);

/* Multi-GPU (one process per GPU): NCCL communicator owned by the ctx.
 * The 128-byte unique id is produced on rank 0 and broadcast by the host's own
 * transport (MPI_Bcast in MPAS; torch.distributed in bench.py). */
int bgc_comm_unique_id(unsigned char id[128]);
int bgc_comm_init_rank(bgc_ctx *ctx, int nranks, int rank, const unsigned char id[128]);
int bgc_inventory_allreduce(bgc_ctx *ctx, double out[BGC_INVENTORY_LEN]);   /* begin + end */
/* Stream-ordered form: _begin enqueues the all-reduce and the copy of its 512-byte result to a
 * page-locked buffer and returns at once; _end waits for that copy (of the latest _begin) only. */
int bgc_inventory_allreduce_begin(bgc_ctx *ctx);
int bgc_inventory_allreduce_end(bgc_ctx *ctx, double out[BGC_INVENTORY_LEN]);

/* Diagnostics accumulation for BGC_MEM_HOST_FORTRAN callers (extension).
 * The host model time-averages the ~160 diagnostic arrays for its history files, yet they
 * are 3/4 of what a host-layout step moves over PCIe.  With accumulation enabled the
 * *_source_sink host calls still return tendencies and PH_PREV_* every step, but add each
 * diagnostic into a device-resident accumulator instead of downloading it (the caller's
 * diagnostic arrays are left untouched; a non-NULL member still selects the array).
 * bgc_diag_flush downloads scale * accumulated sum into the given host arrays (Fortran layout,
 * same extents as the calls that accumulated) and optionally resets the accumulators; any of
 * the three blocks may be NULL.  DMS / MACROS diagnostics accumulate over active cells only
 * (the reference leaves them undefined elsewhere) and are zero there after a flush. */
int bgc_diag_accumulate_enable(bgc_ctx *ctx, int enable);
int bgc_diag_flush(bgc_ctx *ctx, BgcDiagnostics *bgc, DmsDiagnostics *dms, MacrosDiagnostics *macros,
                   int nLevelsMax, int nColumnsMax, double scale, int reset);

/* Page-locked host memory for the HOST_FORTRAN path: arrays allocated here (or
 * registered with bgc_host_register) move over PCIe/NVLink-C2C by DMA at full
 * speed; ordinary pageable arrays work too, through the driver's bounce buffer. */
int bgc_host_alloc(void **ptr, size_t bytes);
int bgc_host_free(void *ptr);
int bgc_host_register(void *ptr, size_t bytes);
int bgc_host_unregister(void *ptr);

/* Layout helpers (device kernels, stream-ordered): Fortran (k,col,n) <-> SoA. */
int bgc_layout_to_soa(bgc_ctx *ctx, const double *dev_fortran, double *dev_soa,
                      int nLevelsMax, int nColumnsMax, int nSlabs);
int bgc_layout_to_fortran(bgc_ctx *ctx, const double *dev_soa, double *dev_fortran,
                          int nLevelsMax, int nColumnsMax, int nSlabs);

/* CUDA graphs.  Everything the BGC_MEM_DEVICE_SOA entry points enqueue between _begin and _end
 * (on the ctx stream and the library's side stream) is captured into one graph instead of
 * being executed; bgc_graph_launch replays it on the ctx stream.  The same calls must have run
 * once before (the device arena may not grow and per-kernel timing must be off while
 * capturing); the captured pointers and extents are baked in.  Synchronising entry points
 * (bgc_ctx_synchronize, bgc_inventory_get, bgc_inventory_allreduce[_end], bgc_get_status, any
 * BGC_MEM_HOST_FORTRAN call) must not be called while capturing. */
typedef struct bgc_graph bgc_graph;
int bgc_graph_capture_begin(bgc_ctx *ctx);
int bgc_graph_capture_end(bgc_ctx *ctx, bgc_graph **out);
int bgc_graph_launch(bgc_ctx *ctx, bgc_graph *graph);
int bgc_graph_destroy(bgc_graph *graph);

/* MPAS tracer layout (extension).  MPAS-Ocean stores a tracer group
 * as T(iTracer, k, iCell), tracer index fastest.  These stream-ordered device kernels move such
 * an array to / from the library's SoA layout without the (k,col,tracer) intermediate of the
 * reference API.  slot_of_tracer[n] (n = 0 .. nTracers-1) is the 1-based SoA slot that MPAS
 * tracer n feeds (0 = tracer not used).  bgc_layout_soa_to_mpas computes
 *     T(n,k,cell) = beta * T(n,k,cell) + alpha * soa(cell,k,slot[n]);
 * with soa = the tendency array, alpha = dt and beta = 1 it is the explicit tracer update fused
 * with the layout change (beta = 0 just converts).  All pointers are device pointers. */
int bgc_layout_mpas_to_soa(bgc_ctx *ctx, const double *dev_mpas, double *dev_soa, int nTracers,
                           const int *slot_of_tracer, int nLevelsMax, int nColumnsMax);
int bgc_layout_soa_to_mpas(bgc_ctx *ctx, const double *dev_soa, double *dev_mpas, int nTracers,
                           const int *slot_of_tracer, int nLevelsMax, int nColumnsMax,
                           double alpha, double beta);
/* The same with a per-cell weight in the MPAS layout, w(k, iCell), level fastest:
 *     T(n,k,cell) = beta * T(n,k,cell) + alpha * w(k,cell) * soa(cell,k,slot[n]).
 * With w = layerThickness, alpha = 1, beta = 1 and T = the tracer group's tendency array this is
 * how MPAS-Ocean folds the BGC tendencies (BGC_mod.F90:1583-1790) into its thickness-weighted
 * tracer tendencies; dev_weight == NULL means w = 1. */
int bgc_layout_soa_to_mpas_weighted(bgc_ctx *ctx, const double *dev_soa, double *dev_mpas, int nTracers,
                                    const int *slot_of_tracer, int nLevelsMax, int nColumnsMax,
                                    double alpha, double beta, const double *dev_weight);

/* Device-resident model state (extension).  The fields the reference
 * carries from one time step to the next and writes to restart files - PH_PREV_3D,
 * PH_PREV_ALT_CO2_3D (BGC_output_type, BGC_parms.F90:170-171) and surface_pH, surface_pH_alt_co2
 * (BGC_forcing_type, BGC_parms.F90:151-152) - can live in the ctx between steps: a
 * device-resident host model passes bgc_state_device_ptr's pointers in its BGC_MEM_DEVICE_SOA
 * argument blocks every step and touches the host copies only at restart time.
 * bgc_state_set uploads a host array in the reference layout ((k,col) level fastest for the 3-D
 * fields, (col) for the surface ones), bgc_state_get downloads it (synchronous).  The resident
 * arrays are created zero-filled on first use - zero is the reference's "no previous pH" marker
 * (BGC_mod.F90:944, :2873). */
enum { BGC_STATE_PH_PREV_3D = 0, BGC_STATE_PH_PREV_ALT_CO2_3D = 1, BGC_STATE_SURFACE_PH = 2,
       BGC_STATE_SURFACE_PH_ALT_CO2 = 3, BGC_STATE_COUNT = 4 };
int bgc_state_device_ptr(bgc_ctx *ctx, int which, int nLevelsMax, int nColumnsMax, double **dev_ptr);
int bgc_state_set(bgc_ctx *ctx, int which, const double *host, int nLevelsMax, int nColumnsMax);
int bgc_state_get(bgc_ctx *ctx, int which, double *host, int nLevelsMax, int nColumnsMax);

#ifdef __cplusplus
}
#endif
#endif /* BGC_B200_H */
