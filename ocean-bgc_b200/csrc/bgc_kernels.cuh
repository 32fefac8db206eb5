// bgc_kernels.cuh — kernel argument blocks and launch entry points shared by
// bgc_kernels.cu (device code) and bgc_capi.cu (the C ABI).
//
// Device layout ("SoA", column fastest): A(k,col[,n]) at col + nC*(k + nL*n);
// surface / flux arrays F(col[,n]) at col + nC*n.  Consecutive threads own
// consecutive columns, so every level-by-level load or store of a warp is one
// contiguous 256-byte run of FP64.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "bgc_b200.h"

namespace bgc {

// Parameter tables resident in __constant__ memory (one copy per device).
struct BgcTables {
  BgcParams p;
  BgcAutotroph a[BGC_AUTOTROPH_CNT];
  BgcIndices ind;
  int same_grazee[BGC_AUTOTROPH_CNT][BGC_AUTOTROPH_CNT];   // grazee_ind(i) == grazee_ind(j)
};
struct DmsTables { DmsParams p; DmsIndices ind; };
struct MacrosTables { MacrosParams p; MacrosIndices ind; };

cudaError_t upload_bgc_tables(const BgcTables &t, cudaStream_t s);
cudaError_t upload_dms_tables(const DmsTables &t, cudaStream_t s);
cudaError_t upload_macros_tables(const MacrosTables &t, cudaStream_t s);

// ---- carbonate kernel, one thread per CELL (no vertical coupling)
struct Co3Args {
  int nL, nC, nColumns;
  const double *tracers;                 // (k,col,30) SoA
  const double *T, *S, *zmid;            // (k,col)
  const int *kmax;                       // (col)
  double *ph_prev, *ph_prev_alt;         // (k,col) read-modify-write on active cells
  // outputs; any may be NULL.  co3/sat_calc/sat_arag are also consumed by the
  // column sweep (saturation-depth scan), so the caller always provides them
  // (diagnostic arrays or ctx scratch).
  double *co3, *hco3, *h2co3, *ph, *co3_alt, *hco3_alt, *h2co3_alt, *ph_alt, *sat_calc, *sat_arag;
  unsigned long long *status;            // BgcStatus counters
};
cudaError_t launch_co3_cells(const Co3Args &a, cudaStream_t s);

// ---- ecosystem + particulate column sweep, one thread per COLUMN
struct EcoArgs {
  int nL, nC, nColumns, alt_co2_use_eco;
  const double *tracers;                 // (k,col,30)
  const double *T, *S, *zmid, *dz, *zbot;
  const double *lat;
  const int *kmax;
  const double *fesedflux, *rtau, *no3_clim, *po4_clim, *sio3_clim;   // (k,col); *_clim only if lrest_*
  const double *dust_flux_in, *sw_flux;  // (col)
  const double *co3, *sat_calc, *sat_arag;   // from launch_co3_cells
  double *tend;                          // (k,col,30)
  BgcDiagnostics d;                      // carbonate + never-touched members nulled by the caller
  unsigned long long *status;
};
// diag_mode: 0 = no diagnostic array, 1 = any subset (NULL-checked stores), 2 = every array
// the kernel owns is present (unchecked stores).  variant selects the launch shape
// (k_eco.cu: launch_diag); 0 = default.
cudaError_t launch_eco_columns(const EcoArgs &a, int diag_mode, int variant, cudaStream_t s);

// ---- surface fluxes, one thread per column
struct SurfArgs {
  int nL, nC, nColumns;
  const double *tracers;
  BgcForcing f;                          // device pointers, F(col[,n])
  BgcFluxDiagnostics d;
  unsigned long long *status;
};
cudaError_t launch_surface_fluxes(const SurfArgs &a, cudaStream_t s);

// ---- batched co2calc_1point
struct Co2PointsArgs {
  int n;
  const double *depth, *temp, *salt, *dic, *ta, *pt, *sit, *phlo, *phhi, *xco2, *atmpres;
  double *ph, *co2star, *dco2star, *pco2surf, *dpco2;
  unsigned long long *status;
};
cudaError_t launch_co2calc_points(const Co2PointsArgs &a, cudaStream_t s);

// ---- DMS / MACROS
struct DmsArgs {
  int nL, nC, nColumns;
  const double *tracers, *dz;
  const int *kmax;
  const double *sst, *sw_flux;
  double *tend;
  DmsDiagnostics d;
};
cudaError_t launch_dms_columns(const DmsArgs &a, cudaStream_t s);

struct DmsSurfArgs {
  int nL, nC, nColumns;
  const double *tracers;
  DmsForcing f;
  DmsFluxDiagnostics d;
};
cudaError_t launch_dms_surface(const DmsSurfArgs &a, cudaStream_t s);

struct MacrosArgs {
  int nL, nC, nColumns;
  const double *tracers;
  const int *kmax;
  double *tend;
  MacrosDiagnostics d;
};
cudaError_t launch_macros_cells(const MacrosArgs &a, cudaStream_t s);

// ---- layout: Fortran (k fastest) <-> SoA (column fastest), nSlabs 2-D slabs
cudaError_t launch_transpose(const double *src, double *dst, int rows_fast_src, int cols_slow_src,
                             int nSlabs, cudaStream_t s);

// ---- inventory: sum_col sum_k tend(n)*dz over active cells (+ sums of per-column
// diagnostics), deterministic, accumulated into the ctx inventory vector
constexpr int kInvGroup = 8;        // values per group (tracer slots, or 6 slots + the two counters in group 0)
constexpr int kInvMaxGroups = 6;
struct InventoryArgs {
  int nL, nC, nColumns, nGroups;
  const double *tend, *dz;
  const int *kmax;
  int slot[kInvMaxGroups][kInvGroup];        // 0-based tracer slot, -1 = unused
  int out_index[kInvMaxGroups][kInvGroup];   // destination in the inventory vector
  int count_out;                             // >= 0: group 0's last two values are (active cells, active columns)
  const double *colsum[kInvGroup];           // per-column arrays summed as one extra group (NULL = unused)
  int colsum_out;
  double *partials;                          // [inventory_grid][groups][kInvGroup]
  double *inventory;                         // accumulated into (+=)
};
int inventory_grid(int nColumns);
cudaError_t launch_inventory(const InventoryArgs &a, cudaStream_t s);

}  // namespace bgc
