/* bgc_oracle.h — CPU oracle for the Ocean-BGC column hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * Fortran (E3SM-Project/Ocean-BGC), routine by routine and in the same
 * evaluation order, used as the checker in tests/, in __graft_entry__.smoke()
 * and as the `cpu_baseline` / `--impl reference` leg of bench.py.  Nothing in
 * the product path (ocean-bgc_b200/, include/) links, imports or calls it.
 *
 * PINNED AGAINST THE REFERENCE RUN THROUGH A SOURCE TRANSLATION: the reference ships no
 * tests, fixtures or golden vectors and no Fortran compiler exists in this image, so the
 * reference's own sources are machine-translated to C (oracle/f90c.py) and compiled with gcc
 * into oracle/_ref/libbgc_ref.so; every routine of this oracle is compared with it BIT FOR BIT
 * on the same inputs (tests/test_reference_translated.py) and tests/golden/ holds outputs of
 * that library.  Caveat: a translation by the same author, not a gfortran build
 * (DESIGN.md section 4 lists what could still slip through; tests/fortran/ closes it wherever
 * gfortran exists).  Further anchors: the comment-level known answers in the reference
 * (O2SAT(10,35)=282.015, dust_to_Fe=626712), published check values of the carbonate
 * constants, the code's own conservation diagnostics Jint_*tot ~ 0, an independent NumPy
 * restatement of co2calc (oracle/co2calc_numpy.py).
 *
 * Build: gcc -O2 -ffp-contract=off (mirrors gfortran -O2 on x86-64: no FMA
 * contraction, no reassociation); see oracle/Makefile.
 */
#ifndef BGC_ORACLE_H
#define BGC_ORACLE_H

#include "bgc_b200.h"   /* interface structs only (the reference's derived types) */

#ifdef __cplusplus
extern "C" {
#endif

/* co2calc module SAVE scratch (co2calc.F90:65-67), made explicit so the oracle
 * is re-entrant (the reference is not). */
typedef struct OracleCo2Save {
  double kw, kb, ks, kf, k1p, k2p, k3p, ksi, bt, st, ft, dic, ta, pt, sit;
} OracleCo2Save;

/* counters filled by the solver for the iteration-count known answers */
typedef struct OracleSolverStats {
  long talk_row_calls;
  long bracket_grow;
  long newton_iters;
  long no_convergence;
} OracleSolverStats;

/* ---- BGC_parms.F90 ---- */
void oracle_BGC_parms_init(BgcParams *p, BgcAutotroph a[4], BgcIndices *ind,
                           int default_real_8);
void oracle_BGC_init(const BgcIndices *ind, BgcAutotroph a[4]);
void oracle_DMS_parms_init(DmsParams *p);
void oracle_MACROS_parms_init(MacrosParams *p);
double oracle_dust_to_Fe(void);

/* ---- co2calc.F90 ---- */
void oracle_comp_co3_coeffs(int k, double depth, double temp, double salt,
                            double *sk0, double *sk1, double *sk2, double *sff,
                            int k1_k2_pH_tot, OracleCo2Save *sv);
void oracle_talk_row(double k1, double k2, double x, double *fn, double *df,
                     const OracleCo2Save *sv);
void oracle_drtsafe_row(int k, double k1, double k2, double *x1, double *x2,
                        double xacc, double *soln, const OracleCo2Save *sv,
                        OracleSolverStats *st);
void oracle_comp_htotal(int k, double temp, double dic_in, double ta_in, double pt_in,
                        double sit_in, double k1, double k2, double *phlo, double *phhi,
                        double *htotal, OracleCo2Save *sv, OracleSolverStats *st);
void oracle_co2calc_1point(double depth, int locmip_k1_k2_bug_fix, int lcomp_co3_coeffs,
                           double temp, double salt, double dic_in, double ta_in,
                           double pt_in, double sit_in, double *phlo, double *phhi,
                           double *ph, double xco2_in, double atmpres, double *co2star,
                           double *dco2star, double *pCO2surf, double *dpco2,
                           OracleSolverStats *st);
void oracle_comp_CO3terms(int k, double depth, int lcomp_co3_coeffs, double temp,
                          double salt, double dic_in, double ta_in, double pt_in,
                          double sit_in, double *phlo, double *phhi, double *pH,
                          double *H2CO3, double *HCO3, double *CO3, OracleSolverStats *st);
void oracle_comp_co3_sat_vals(int k, double depth, double temp, double salt,
                              double *co3_sat_calc, double *co3_sat_arag);

/* batched co2calc_1point over n points (config 2) */
void oracle_co2calc_points(int n, const double *depth, const double *temp,
                           const double *salt, const double *dic, const double *ta,
                           const double *pt, const double *sit, const double *phlo,
                           const double *phhi, const double *xco2, const double *atmpres,
                           double *ph, double *co2star, double *dco2star,
                           double *pco2surf, double *dpco2, OracleSolverStats *st,
                           int nthreads);

/* batched comp_co3_coeffs dump: out[n][14] = k0,k1,k2,ff,kw,kb,ks,kf,k1p,k2p,k3p,ksi,bt,st */
void oracle_co3_coeffs_points(int n, const int *k, const double *depth, const double *temp,
                              const double *salt, double *out);

/* ---- BGC_mod.F90 ---- */
double oracle_O2SAT_singleValue(double SST, double SSS, double T0_Kelvin_BGC);
double oracle_SCHMIDT_O2_singleValue(double SST);
double oracle_SCHMIDT_CO2_singleValue(double SST);

void oracle_BGC_SourceSink(const BgcParams *p, const BgcAutotroph autotrophs[4],
                           const BgcIndices *ind, const BgcInput *in,
                           const BgcForcing *forcing, BgcOutput *out, BgcDiagnostics *diag,
                           int numLevelsMax, int numColumnsMax, int numColumns,
                           int alt_co2_use_eco, int nthreads, OracleSolverStats *st);

void oracle_BGC_SurfaceFluxes(const BgcParams *p, const BgcIndices *ind, const BgcInput *in,
                              BgcForcing *forcing, BgcFluxDiagnostics *diag,
                              int numLevelsMax, int numColumnsMax, int numColumns,
                              int nthreads);

/* ---- DMS_mod.F90 / MACROS_mod.F90 ---- */
double oracle_SCHMIDT_DMS_singleValue(double SST);
void oracle_DMS_SourceSink(const DmsParams *p, const DmsIndices *ind, const DmsInput *in,
                           const DmsForcing *forcing, DmsOutput *out, DmsDiagnostics *diag,
                           int numLevelsMax, int numColumnsMax, int numColumns, int nthreads);
void oracle_DMS_SurfaceFluxes(const DmsParams *p, const DmsIndices *ind, const DmsInput *in,
                              DmsForcing *forcing, DmsFluxDiagnostics *diag,
                              int numLevelsMax, int numColumnsMax, int numColumns);
void oracle_MACROS_SourceSink(const MacrosParams *p, const MacrosIndices *ind,
                              const MacrosInput *in, MacrosOutput *out,
                              MacrosDiagnostics *diag, int numLevelsMax,
                              int numColumnsMax, int numColumns, int nthreads);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
