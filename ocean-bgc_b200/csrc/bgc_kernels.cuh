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

constexpr int kInvGroup = 8;        // inventory: values per group (tracer slots; group 0 may end with the two counters)
constexpr int kInvMaxGroups = 6;
constexpr int kEcoInvGroups = 5;   // inventory groups produced by the column sweep (see below)

// ---- carbonate kernel, one thread per CELL (no vertical coupling)

// ---- debugging aid (BGC_BLOCK_TRACE_FILE): where and when every thread block of the two kernels of
// bgc_source_sink ran.  trace[0] counts the records; record r = trace[4 + 4 r ...] =
// {kernel id << 32 | SM id, start ns, end ns, blockIdx.x} (globaltimer).  NULL in production.
constexpr unsigned kBlockTraceCap = 1u << 16;
#ifdef __CUDACC__
__device__ __forceinline__ unsigned block_trace_begin(unsigned long long *trace, unsigned kid) {
  if (!trace) return ~0u;
  const unsigned slot = (unsigned)atomicAdd(trace, 1ull);
  if (slot >= kBlockTraceCap) return ~0u;
  unsigned smid; unsigned long long t;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  trace[4 + 4 * (size_t)slot] = ((unsigned long long)kid << 32) | smid;
  trace[5 + 4 * (size_t)slot] = t;
  trace[7 + 4 * (size_t)slot] = blockIdx.x;
  return slot;
}
__device__ __forceinline__ void block_trace_end(unsigned long long *trace, unsigned slot) {
  if (slot == ~0u) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  trace[6 + 4 * (size_t)slot] = t;
}
#endif

struct Co3Args {
  int nL, nC, nColumns;
  const double *tracers;                 // (k,col,30) SoA
  const double *T, *S, *zmid;            // (k,col)
  const int *kmax;                       // (col)
  double *ph_prev, *ph_prev_alt;         // (k,col) read-modify-write on active cells
  // outputs; any may be NULL.  co3/sat_calc/sat_arag are also consumed by the
  // saturation-depth scan (zsat_columns_kernel), so the caller provides them
  // (diagnostic arrays or ctx scratch) whenever diag_zsatcalc / diag_zsatarag are wanted.
  double *co3, *hco3, *h2co3, *ph, *co3_alt, *hco3_alt, *h2co3_alt, *ph_alt, *sat_calc, *sat_arag;
  unsigned long long *status;            // BgcStatus counters
  // the launch's share of the mesh: cells [cell_begin, cell_end) in (k,col) order; cell_end = 0 means
  // "to the end".  cell_begin must be a multiple of 32 (a warp never straddles two launches).
  unsigned long long cell_begin = 0, cell_end = 0;
  unsigned long long *block_trace = nullptr;   // debugging aid, see block_trace_begin
};
cudaError_t launch_co3_cells(const Co3Args &a, int persistent_blocks, cudaStream_t s);

// ---- saturation-depth scan (BGC_mod.F90:1003-1032), one thread per COLUMN, after the carbonate kernel
struct ZsatArgs {
  int nL, nC, nColumns;
  const int *kmax;
  const double *co3, *sat_calc, *sat_arag;   // from launch_co3_cells
  const double *zmid, *zbot;
  double *zsatcalc, *zsatarag;               // (col); either may be NULL
};
cudaError_t launch_zsat_columns(const ZsatArgs &a, cudaStream_t s);

// ---- ecosystem + particulate column sweep, one thread per COLUMN
struct EcoArgs {
  int nL, nC, nColumns, alt_co2_use_eco;
  int any_restore = 0;                   // lrest_no3 | lrest_po4 | lrest_sio3 (set by the caller from the ctx tables)
  int zero_shortcut;                     // skip the body of a functional group whose biomass is zero in a whole warp
  const double *tracers;                 // (k,col,30)
  const double *T, *S, *zmid, *dz, *zbot;
  const double *lat;
  const int *kmax;
  const double *fesedflux, *rtau, *no3_clim, *po4_clim, *sio3_clim;   // (k,col); *_clim only if lrest_*
  const double *dust_flux_in, *sw_flux;  // (col)
  double *tend;                          // (k,col,30)
  BgcDiagnostics d;                      // carbonate, zsat* and never-touched members nulled by the caller
  unsigned long long *status;
  double *inv_partials;                  // NULL, or the fused stage 1 of the inventory reduction
  unsigned long long *block_trace = nullptr;   // debugging aid, see block_trace_begin
  int bulk;                              // set by launch_eco_columns: stage the inputs with TMA bulk copies (16-byte aligned slabs)
  // Canonical stage row -> 1-based tracer slot (k_eco.cu: rows 0..15 the plain tracers in BgcIndices
  // order, 16+3a+{0,1,2} = C, Chl, Fe of group a, 28 = the Si tracer, 29 = the CaCO3 tracer); filled by
  // eco_rows_from_tables.  tend_off[row] = (slot - 1) * nL * nC is set by launch_eco_columns.
  int slot_of_row[BGC_TRACER_CNT];
  unsigned tend_off[BGC_TRACER_CNT];
};
// fills a.slot_of_row from the ctx's index and functional-group tables; false if they do not cover the 30 slots
bool eco_rows_from_tables(const BgcTables &t, EcoArgs &a);
// diag_mode: 0 = no diagnostic array, 1 = any subset (NULL-checked stores), 2 = every array
// the kernel owns is present (unchecked stores).  variant selects the launch shape
// (k_eco.cu: launch_diag); 0 = default.
cudaError_t launch_eco_columns(const EcoArgs &a, int diag_mode, int variant, cudaStream_t s);
int eco_inventory_parts(const EcoArgs &a, int diag_mode, int variant);   // blocks the launch will use
int eco_sweep_blocks(int nC, int variant);   // thread blocks of the column sweep (one per SM: k_eco.cu)

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

// ---- batched comp_CO3terms / comp_co3_sat_vals (the rest of the co2calc module's public trio)
struct Co3TermsPointsArgs {
  int n, k_all;                 // k_all: the level index of every point when k == NULL
  const int *k;                 // 1-based level index per point, or NULL
  const double *depth, *temp, *salt, *dic, *ta, *pt, *sit, *phlo, *phhi;
  double *ph, *h2co3, *hco3, *co3;
  unsigned long long *status;
};
cudaError_t launch_co3terms_points(const Co3TermsPointsArgs &a, cudaStream_t s);
struct Co3SatPointsArgs {
  int n, k_all;
  const int *k;
  const double *depth, *temp, *salt;
  double *sat_calc, *sat_arag;
};
cudaError_t launch_co3_sat_points(const Co3SatPointsArgs &a, cudaStream_t s);

// ---- DMS / MACROS
struct DmsArgs {
  int nL, nC, nColumns;
  const double *tracers, *dz;
  const int *kmax;
  const double *sst, *sw_flux;
  double *tend;
  DmsDiagnostics d;
  double *inv_partials;   // NULL, or [dms_inventory_parts][kInvGroup]: fused stage 1 of the inventory
  int l2_prefetch = 0;    // set by launch_dms_columns
};
cudaError_t launch_dms_columns(const DmsArgs &a, int variant, cudaStream_t s);

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
  const double *dz;       // read only for the inventory
  double *inv_partials;   // NULL, or [macros_inventory_parts][kInvGroup]
};
cudaError_t launch_macros_cells(const MacrosArgs &a, cudaStream_t s);

// ---- layout: Fortran (k fastest) <-> SoA (column fastest), nSlabs 2-D slabs
cudaError_t launch_transpose(const double *src, double *dst, int rows_fast_src, int cols_slow_src,
                             int nSlabs, cudaStream_t s);

// ---- MPAS tracer layout (tracer index fastest: T(n, k, cell) at n + nT*(k + nL*cell)) <-> SoA
// (the caller side of the boundary).  slot[n] = 1-based SoA slot of MPAS
// tracer n, 0 = not part of this tracer group.
constexpr int kMpasMaxTracers = 64;
struct MpasMap { int nT; int slot[kMpasMaxTracers]; unsigned long long used; };   // used: bit n set <=> slot[n] > 0
cudaError_t launch_mpas_to_soa(const double *mpas, double *soa, const MpasMap &m, int nL, int nC, cudaStream_t s);
// mpas(n,k,cell) = beta * mpas(n,k,cell) + alpha * weight(k,cell) * soa(cell,k,slot[n]); weight = NULL
// means 1.  alpha = dt, beta = 1 is the explicit tracer update fused with the layout change; with
// weight = layerThickness(k,cell) it accumulates the thickness-weighted tendency the way MPAS-Ocean does.
cudaError_t launch_soa_to_mpas(const double *soa, double *mpas, const MpasMap &m, int nL, int nC, double alpha,
                               double beta, const double *weight, cudaStream_t s);

// ---- on-device accumulation of diagnostics (the host time-averages the
// diagnostics for history files; accumulating on the device and downloading at output
// frequency removes their PCIe traffic from every step).  acc(k, c0+col, slab) += w * src(k, col, slab)
// for a chunk of `cc` columns of a mesh of `nC` columns; kmax != NULL restricts the update to
// active cells (k < kmax[col], col < nColumns) for diagnostics the reference leaves undefined elsewhere.
cudaError_t launch_accumulate(const double *src, double *acc, int nL, int cc, int nC, int c0, int nSlabs,
                              const int *kmax, int nColumns, double w, cudaStream_t s);
cudaError_t launch_scale(double *a, size_t n, double w, cudaStream_t s);
// the bottom cell of every column of up to 16 (k,col) arrays -> out[j][col] (k_misc.cu)
struct BottomGatherArgs { const double *src[16]; double *out; const int *kmax; int nL, cc, nColumns, n; };
cudaError_t launch_bottom_gather(const BottomGatherArgs &a, cudaStream_t s);

// ---- inventory: sum_col sum_k tend(n)*dz over active cells (+ sums of per-column
// diagnostics).  Stage 1 is fused into the source-sink kernels (block partials); stage 2:
struct InventoryFoldArgs {
  int nGroups;                               // partial layout [nParts][nGroups][kInvGroup]
  int out_index[kInvMaxGroups][kInvGroup];   // destination in the inventory vector, -1 = ignore
  const double *partials;
  double *inventory;                         // accumulated into (+=)
};
cudaError_t launch_inventory_fold(const InventoryFoldArgs &a, int nParts, cudaStream_t s);
int dms_inventory_parts(int nL, int nC, int variant);
int macros_inventory_parts(int nL, int nC);
// column sweep: [eco_inventory_parts][kEcoInvGroups][kInvGroup] = 40 values per block:
//   [0..29] the 30 tracer slots (0-based slot = value index), [30] active cells, [31] active columns,
//   [32..39] the eight Jint_* column sums

}  // namespace bgc
