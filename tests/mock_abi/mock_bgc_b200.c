/* mock_bgc_b200.c — a CPU stand-in for the handful of C-ABI entry points the Fortran shim binds
 * (include/bgc_b200.h), backed by the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The product (ocean-bgc_b200/csrc/libbgc_b200.so) has no CPU path and
 * fails loudly without a CUDA device; this file exists so that the MARSHALLING of the shim
 * (ocean-bgc_b200/fortran/: which component goes to which struct member, in which order, with
 * which extents and flags) can be exercised on a machine without a GPU
 * (tests/test_fortran_shim.py).  On the GPU box the same shim is loaded against the real library.
 * It also records what it was handed, so that the test can check the call sequence.
 */
#include <stdlib.h>
#include <string.h>
#include "bgc_b200.h"
#include "bgc_oracle.h"

struct bgc_ctx {
  int device, nLevelsMax, nColumnsMax;
  int have_bgc, have_dms, have_macros;
  BgcParams p; BgcAutotroph a[4]; BgcIndices ind;
  DmsParams dp; DmsIndices dind;
  MacrosParams mp; MacrosIndices mind;
};

static char g_err[512] = "";
static int g_live = 0, g_created = 0, g_calls = 0, g_last_device = -1;

const char *bgc_last_error(void) { return g_err; }
int mock_live_contexts(void) { return g_live; }
int mock_created_contexts(void) { return g_created; }
int mock_compute_calls(void) { return g_calls; }
int mock_last_device(void) { return g_last_device; }

static int fail(int code, const char *msg) { strncpy(g_err, msg, sizeof g_err - 1); return code; }

int bgc_ctx_create(int device, int nLevelsMax, int nColumnsMax, bgc_ctx **out) {
  if (!out || nLevelsMax < 1 || nColumnsMax < 1) return fail(BGC_ERR_ARG, "mock: bad ctx_create arguments");
  bgc_ctx *c = calloc(1, sizeof *c);
  c->device = device; c->nLevelsMax = nLevelsMax; c->nColumnsMax = nColumnsMax;
  *out = c; g_live++; g_created++; g_last_device = device;
  return BGC_OK;
}
int bgc_ctx_destroy(bgc_ctx *c) { if (!c) return fail(BGC_ERR_ARG, "mock: null ctx"); free(c); g_live--; return BGC_OK; }

int bgc_set_params(bgc_ctx *c, const BgcParams *p, const BgcAutotroph a[4], const BgcIndices *ind) {
  if (!c || !p || !a || !ind) return fail(BGC_ERR_ARG, "mock: null argument to bgc_set_params");
  c->p = *p; memcpy(c->a, a, sizeof c->a); c->ind = *ind; c->have_bgc = 1; return BGC_OK;
}
int dms_set_params(bgc_ctx *c, const DmsParams *p, const DmsIndices *ind) {
  if (!c || !p || !ind) return fail(BGC_ERR_ARG, "mock: null argument to dms_set_params");
  c->dp = *p; c->dind = *ind; c->have_dms = 1; return BGC_OK;
}
int macros_set_params(bgc_ctx *c, const MacrosParams *p, const MacrosIndices *ind) {
  if (!c || !p || !ind) return fail(BGC_ERR_ARG, "mock: null argument to macros_set_params");
  c->mp = *p; c->mind = *ind; c->have_macros = 1; return BGC_OK;
}

static int check(bgc_ctx *c, int have, int nL, int nC, int n, int mem_space) {
  if (!c) return fail(BGC_ERR_ARG, "mock: null ctx");
  if (!have) return fail(BGC_ERR_PARAMS, "mock: *_set_params not called");
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "mock: host layout only");
  if (nL > c->nLevelsMax || nC > c->nColumnsMax || n > nC || n < 0)
    return fail(BGC_ERR_ARG, "mock: block larger than the ctx");
  g_calls++;
  return BGC_OK;
}

int bgc_source_sink(bgc_ctx *c, const BgcInput *in, const BgcForcing *fo, BgcOutput *out, BgcDiagnostics *dg,
                    int nL, int nC, int n, int alt, int mem_space) {
  int rc = check(c, c ? c->have_bgc : 0, nL, nC, n, mem_space);
  if (rc) return rc;
  if (getenv("MOCK_BGC_FAIL")) return fail(BGC_ERR_CUDA, "mock: simulated CUDA failure in bgc_source_sink");
  oracle_BGC_SourceSink(&c->p, c->a, &c->ind, in, fo, out, dg, nL, nC, n, alt, 1, NULL);
  return BGC_OK;
}
int bgc_surface_fluxes(bgc_ctx *c, const BgcInput *in, BgcForcing *fo, BgcFluxDiagnostics *dg,
                       int nL, int nC, int n, int mem_space) {
  int rc = check(c, c ? c->have_bgc : 0, nL, nC, n, mem_space);
  if (rc) return rc;
  oracle_BGC_SurfaceFluxes(&c->p, &c->ind, in, fo, dg, nL, nC, n, 1);
  return BGC_OK;
}
int dms_source_sink(bgc_ctx *c, const DmsInput *in, const DmsForcing *fo, DmsOutput *out, DmsDiagnostics *dg,
                    int nL, int nC, int n, int mem_space) {
  int rc = check(c, c ? c->have_dms : 0, nL, nC, n, mem_space);
  if (rc) return rc;
  oracle_DMS_SourceSink(&c->dp, &c->dind, in, fo, out, dg, nL, nC, n, 1);
  return BGC_OK;
}
int dms_surface_fluxes(bgc_ctx *c, const DmsInput *in, DmsForcing *fo, DmsFluxDiagnostics *dg,
                       int nL, int nC, int n, int mem_space) {
  int rc = check(c, c ? c->have_dms : 0, nL, nC, n, mem_space);
  if (rc) return rc;
  oracle_DMS_SurfaceFluxes(&c->dp, &c->dind, in, fo, dg, nL, nC, n);
  return BGC_OK;
}
int macros_source_sink(bgc_ctx *c, const MacrosInput *in, MacrosOutput *out, MacrosDiagnostics *dg,
                       int nL, int nC, int n, int mem_space) {
  int rc = check(c, c ? c->have_macros : 0, nL, nC, n, mem_space);
  if (rc) return rc;
  oracle_MACROS_SourceSink(&c->mp, &c->mind, in, out, dg, nL, nC, n, 1);
  return BGC_OK;
}

/* the two other public procedures of the reference's co2calc module, batched (host arrays only) */
int bgc_comp_co3terms(bgc_ctx *c, int n, const int *k_level, int k_all, const double *depth, const double *temp,
                      const double *salt, const double *dic, const double *ta, const double *pt, const double *sit,
                      const double *phlo, const double *phhi, double *ph, double *h2co3, double *hco3, double *co3,
                      int mem_space) {
  if (!c) return fail(BGC_ERR_ARG, "mock: null ctx");
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "mock: host arrays only");
  g_calls++;
  for (int i = 0; i < n; ++i) {
    double lo = phlo[i], hi = phhi[i];
    OracleSolverStats st = {0};
    oracle_comp_CO3terms(k_level ? k_level[i] : k_all, depth[i], 1, temp[i], salt[i], dic[i], ta[i], pt[i], sit[i],
                         &lo, &hi, &ph[i], &h2co3[i], &hco3[i], &co3[i], &st);
  }
  return BGC_OK;
}

int bgc_comp_co3_sat_vals(bgc_ctx *c, int n, const int *k_level, int k_all, const double *depth, const double *temp,
                          const double *salt, double *co3_sat_calc, double *co3_sat_arag, int mem_space) {
  if (!c) return fail(BGC_ERR_ARG, "mock: null ctx");
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "mock: host arrays only");
  g_calls++;
  for (int i = 0; i < n; ++i)
    oracle_comp_co3_sat_vals(k_level ? k_level[i] : k_all, depth[i], temp[i], salt[i], &co3_sat_calc[i], &co3_sat_arag[i]);
  return BGC_OK;
}

int bgc_co2calc_points(bgc_ctx *c, int n, const double *depth, const double *temp, const double *salt,
                       const double *dic, const double *ta, const double *pt, const double *sit, const double *phlo,
                       const double *phhi, const double *xco2, const double *atmpres, double *ph, double *co2star,
                       double *dco2star, double *pco2surf, double *dpco2, int mem_space) {
  if (!c) return fail(BGC_ERR_ARG, "mock: null ctx");
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "mock: host arrays only");
  g_calls++;
  OracleSolverStats st = {0};
  oracle_co2calc_points(n, depth, temp, salt, dic, ta, pt, sit, phlo, phhi, xco2, atmpres, ph, co2star, dco2star,
                        pco2surf, dpco2, &st, 1);
  return BGC_OK;
}
