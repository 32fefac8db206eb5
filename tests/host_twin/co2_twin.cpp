// Host twin of the carbonate device functions (bgc_co2.cuh): the header's own code compiled for the
// CPU (one lane = one warp), so that co3_coeffs / solve_htotal / co3_sat_vals can be compared with
// the reference without a GPU.  Built twice: production flavour and -DBGC_STRICT.
#include "cuda_runtime.h"
#define asm(...) r = host_rcp_seed(b)
#include "bgc_co2.cuh"
#undef asm
using namespace bgc;
extern "C" {
// out[i][0..13] = k1, k2, ff, kw, kb, ks, kf, k1p, k2p, k3p, ksi, bt, st, ft
void twin_co3_coeffs(int n, const int *k, const double *depth, const double *temp, const double *salt, double *out) {
  for (int i = 0; i < n; ++i) {
    Co3Consts c;
    co3_coeffs<true>(k[i] > 1, depth[i], temp[i], salt[i], c, ExpPoly());
    const double v[14] = {c.k1, c.k2, c.ff, c.kw, c.kb, c.ks, c.kf, c.k1p, c.k2p, c.k3p, c.ksi, c.bt, c.st, c.ft};
    for (int j = 0; j < 14; ++j) out[14 * i + j] = v[j];
  }
}
// htotal and the solver status for comp_CO3terms-style inputs (mmol/m^3, pH brackets)
void twin_htotal(int n, const int *k, const double *depth, const double *temp, const double *salt,
                 const double *dic, const double *ta, const double *pt, const double *sit,
                 const double *phlo, const double *phhi, double *h, int *status) {
  for (int i = 0; i < n; ++i) {
    Co3Consts c;
    co3_coeffs<false>(k[i] > 1, depth[i], temp[i], salt[i], c, ExpPoly());
    const Co3Totals t = co3_totals(dic[i], ta[i], pt[i], sit[i]);
    unsigned st = 0;
    h[i] = solve_htotal(c, t, phlo[i], phhi[i], st);
    status[i] = (int)st;
  }
}
void twin_sat_vals(int n, const int *k, const double *depth, const double *temp, const double *salt,
                   double *calc, double *arag) {
  for (int i = 0; i < n; ++i) co3_sat_vals(k[i] > 1, depth[i], temp[i], salt[i], calc[i], arag[i], ExpPoly());
}
}
