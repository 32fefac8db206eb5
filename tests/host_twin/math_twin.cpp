// Host twin of the device math (bgc_math.cuh, production flavour) for accuracy tests on CPU.
// The one piece of inline PTX (rcp.approx.ftz.f64) is replaced by host_rcp_seed through the
// preprocessor; everything else is the header's own code, FMAs included (std::fma is exact).
#include "cuda_runtime.h"
#define asm(...) r = host_rcp_seed(b)
#include "bgc_math.cuh"
#undef asm
extern "C" {
void twin_exp(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = bgc::bexp(x[i]); }
void twin_log(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = bgc::blog(x[i]); }
void twin_rcp(int n, const double *x, double *y) { for (int i = 0; i < n; ++i) y[i] = bgc::frcp(x[i]); }
void twin_pow(int n, const double *x, const double *e, double *y) { for (int i = 0; i < n; ++i) y[i] = bgc::fpow(x[i], e[i]); }
void twin_exp_table(int n, const double *x, double *y) {
  bgc::ExpTable f{bgc::kExp2Tab};
  for (int i = 0; i < n; ++i) y[i] = f(x[i]);
}
}
