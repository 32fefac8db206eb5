// k_misc.cu — DMS / MACROS source-sink kernels, the DMS surface flux, the
// Fortran<->SoA layout transposes and the deterministic inventory reductions.
//
//   dms_cells_kernel      <- DMS_SourceSink     DMS_mod.F90:156-770   (block = 32 columns x all levels, thread = cell)
//   dms_surface_kernel    <- DMS_SurfaceFluxes  DMS_mod.F90:778-908   (thread = column)
//   macros_cells_kernel   <- MACROS_SourceSink  MACROS_mod.F90:137-411 (thread = cell; no vertical coupling)
//
// Both source-sink kernels are HBM-bound streaming kernels: each input element
// is read once and each output element written once, coalesced.
#include <stdlib.h>
#include "bgc_kernels.cuh"
#include "bgc_math.cuh"
#include "bgc_reduce.cuh"

namespace bgc {

__constant__ DmsTables c_dms;
__constant__ MacrosTables c_macros;

cudaError_t upload_dms_tables(const DmsTables &t, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_dms, &t, sizeof(DmsTables), 0, cudaMemcpyHostToDevice, s);
}
cudaError_t upload_macros_tables(const MacrosTables &t, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_macros, &t, sizeof(MacrosTables), 0, cudaMemcpyHostToDevice, s);
}

namespace {

constexpr double dms_epsC = 1.00e-8;   // DMS_parms.F90:194-195 (carries the _r8 suffix: exact)

// ALLDIAG: every diagnostic array is present (unchecked stores)
#define DST(name, val) do { if (ALLDIAG || A.d.name) A.d.name[i2] = (val); } while (0)

// Column-constant factors of DMS_SourceSink: all depend on SST only (DMS_mod.F90:584-592, :637-640).
struct DmsColumnConsts { double cyano_T, yield; };

__device__ __forceinline__ DmsColumnConsts dms_column_consts(double SST_loc) {
  const DmsParams &P = c_dms.p;
  double T_ind = (SST_loc - P.T_lo) / (P.T_hi - P.T_lo);
  if (T_ind <= 0.0) T_ind = 0.0;
  if (T_ind >= 1.0) T_ind = 1.0;
  DmsColumnConsts r;
  r.cyano_T = (T_ind * (P.Max_cyano_frac - P.Min_cyano_frac)) + P.Min_cyano_frac;
  r.yield = (T_ind * (P.Max_yld - P.Min_yld)) + P.Min_yld;
  if (SST_loc < P.T_cryo_hi && SST_loc > P.T_cryo_lo) r.yield = 0.5;
  if (SST_loc < -1.0) r.yield = 0.25;
  return r;
}

// Light attenuation over one cell (DMS_mod.F90:510-527): KPARdz and bexp(-KPARdz).
__device__ __forceinline__ void dms_attenuation(double totalChl, double dz, double &KPARdz, double &eK) {
  const double w = gmax(totalChl, 0.02);
  double kp;
  if (w < 0.13224) kp = 0.000919 * fpow(w, 0.3536);
  else             kp = 0.001131 * fpow(w, 0.4562);
  KPARdz = kp * dz;
  eK = bexp(-KPARdz);
}

// The nine tracers one cell consumes (raw; the clamp is applied in dms_cell).  NO3 and DOC are
// copied by the reference (:471-472) but reach no output (DOC feeds only the unused UV_avg,
// :531-536): not read here.
struct DmsCellIn { double zooC, spC, diatC, diazC, phaeoC, spChl, spCaCO3, dms, dmsp, dz; };   // dz: inventory only

// The cell after next: its nine lines are pulled into L2 (no register, no shared memory), so that the
// register prefetch of the next trip finds them there instead of in HBM.
__device__ __forceinline__ void pf_l2(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void dms_prefetch_cell(const DmsArgs &A, unsigned i2, unsigned nLnC) {
  const DmsIndices &I = c_dms.ind;
  const double *trc = A.tracers;
#define TR(ind_) pf_l2(trc + (i2 + (unsigned)((ind_) - 1) * nLnC))
  TR(I.zooC_ind); TR(I.spC_ind); TR(I.diatC_ind); TR(I.diazC_ind); TR(I.phaeoC_ind); TR(I.spChl_ind);
  TR(I.spCaCO3_ind); TR(I.dms_ind); TR(I.dmsp_ind);
#undef TR
  if (A.inv_partials) pf_l2(A.dz + i2);
}

__device__ __forceinline__ DmsCellIn dms_load_cell(const DmsArgs &A, unsigned i2, unsigned nLnC) {
  const DmsIndices &I = c_dms.ind;
  const double *trc = A.tracers;
#define TR(ind_) trc[i2 + (unsigned)((ind_) - 1) * nLnC]
  DmsCellIn c;
  c.zooC = TR(I.zooC_ind); c.spC = TR(I.spC_ind); c.diatC = TR(I.diatC_ind); c.diazC = TR(I.diazC_ind);
  c.phaeoC = TR(I.phaeoC_ind); c.spChl = TR(I.spChl_ind); c.spCaCO3 = TR(I.spCaCO3_ind);
  c.dms = TR(I.dms_ind); c.dmsp = TR(I.dmsp_ind);
#undef TR
  c.dz = A.inv_partials ? A.dz[i2] : 0.0;
  return c;
}

// Everything of one active cell once PAR_avg is known (DMS_mod.F90:529-765): stores the 27
// diagnostics and returns the two live tendencies.
template <bool ALLDIAG>
__device__ __forceinline__ void dms_cell(const DmsArgs &A, unsigned i2, const DmsCellIn &in, double PAR_avg,
                                         const DmsColumnConsts cc, double &t_dms, double &t_dmsp) {
  const DmsParams &P = c_dms.p;
  const double zooC = gmax(0.0, in.zooC), spC = gmax(0.0, in.spC), diatC = gmax(0.0, in.diatC),
               diazC = gmax(0.0, in.diazC), phaeoC = gmax(0.0, in.phaeoC), spChl = gmax(0.0, in.spChl),
               spCaCO3 = gmax(0.0, in.spCaCO3), DMS_loc = gmax(0.0, in.dms), DMSP_loc = gmax(0.0, in.dmsp);
  const double k_S_p = P.k_S_p_base * (P.mort + cdiv(zooC, 0.3, 1.0 / 0.3));   // literal 0.3, not zooC_avg (:529)
  const double j_dms = P.j_dms_perI * PAR_avg;

  double Fcocco = fdiv(spCaCO3, (spC + dms_epsC));
  if (Fcocco > 0.4) Fcocco = 0.4;
  const double Cocco_frac = Fcocco;
  const double Cyano_frac = (1.0 - Cocco_frac) * cc.cyano_T;
  const double Eukar_frac = 1.0 - Cocco_frac - Cyano_frac;

  const double diatN = P.R * diatC;
  const double phaeoN = P.R * phaeoC;
  const double coccoN = Cocco_frac * P.R * spC;
  const double cyanoN = Cyano_frac * P.R * spC;
  const double eukarN = Eukar_frac * P.R * spC;
  const double diazN = P.R * diazC;
  const double zooN = P.R * zooC;
  const double phytoN = diatN + coccoN + cyanoN + eukarN + diazN + phaeoN;

  double Sp_dec = fdiv((P.Sp_ref - spChl), P.Sp_ref);
  if (Sp_dec <= 0.0) Sp_dec = 0.0;
  if (Sp_dec >= 1.0) Sp_dec = 1.0;
  double Stress_fac = 1.0 + P.Stress_mult * Sp_dec * Sp_dec;
  if (Stress_fac >= 10.0) Stress_fac = 10.0;

  const double diatS = P.Rs2n_diat * diatN;
  const double phaeoS = P.Rs2n_phaeo * phaeoN;
  const double coccoS = P.Rs2n_cocco * coccoN;
  const double cyanoS = P.Rs2n_cyano * cyanoN;
  const double eukarS = P.Rs2n_eukar * eukarN * Stress_fac;
  const double diazS = P.Rs2n_diaz * diazN;
  const double phytoS = diatS + coccoS + cyanoS + eukarS + diazS + P.G_phaeo_S * phaeoS;

  double Rs2n_zoo;
  if (phytoN > 0.0) {
    Rs2n_zoo = (P.Rs2n_diat * diatN +
                P.G_phaeo_S * P.Rs2n_phaeo * phaeoN +
                P.Rs2n_cocco * coccoN +
                P.Rs2n_cyano * cyanoN +
                P.Rs2n_eukar * eukarN * Stress_fac +
                P.Rs2n_diaz * diazN);
    Rs2n_zoo = fdiv(Rs2n_zoo, phytoN);
  } else {
    Rs2n_zoo = (P.Rs2n_diat + P.Rs2n_cocco + P.Rs2n_cyano + P.Rs2n_eukar + P.Rs2n_diaz + P.Rs2n_phaeo) / 6.0;
  }
  const double zooS = Rs2n_zoo * zooN;

  const double B_diagnosed = P.B_preexp * ((phytoN > 0.0) ? fpow(phytoN, P.B_exp) : pow(phytoN, P.B_exp));

  const double dms_s_dmsp = cc.yield * P.k_conv * DMSP_loc;
  const double dms_s = dms_s_dmsp;
  const double dms_r_B = P.k_S_B * B_diagnosed * DMS_loc;
  const double dms_r_phot = j_dms * DMS_loc;
  const double dms_r_bkgnd = P.k_bkgnd * DMS_loc;
  const double dms_r = dms_r_B + dms_r_phot + dms_r_bkgnd;

  const double dmsp_s_phaeo = P.inject_scale * P.k_S_p_base * phaeoS;
  const double dmsp_s_nonphaeo = P.inject_scale * k_S_p * phytoS;
  const double dmsp_s_zoo = P.inject_scale * P.k_S_z * zooS;
  const double dmsp_s = dmsp_s_phaeo + dmsp_s_nonphaeo + dmsp_s_zoo;
  const double dmsp_r_B = P.k_conv * DMSP_loc;
  const double dmsp_r_bkgnd = P.k_bkgnd * DMSP_loc;
  const double dmsp_r = dmsp_r_B + dmsp_r_bkgnd;

  t_dms = dms_s - dms_r;
  t_dmsp = dmsp_s - dmsp_r;

  DST(diag_DMS_S_DMSP, dms_s_dmsp);
  DST(diag_DMS_S_TOTAL, dms_s);
  DST(diag_DMS_R_B, dms_r_B);
  DST(diag_DMS_R_PHOT, dms_r_phot);
  DST(diag_DMS_R_BKGND, dms_r_bkgnd);
  DST(diag_DMS_R_TOTAL, dms_r);
  DST(diag_DMSP_S_PHAEO, dmsp_s_phaeo);
  DST(diag_DMSP_S_NONPHAEO, dmsp_s_nonphaeo);
  DST(diag_DMSP_S_ZOO, dmsp_s_zoo);
  DST(diag_DMSP_S_TOTAL, dmsp_s);
  DST(diag_DMSP_R_B, dmsp_r_B);
  DST(diag_DMSP_R_BKGND, dmsp_r_bkgnd);
  DST(diag_DMSP_R_TOTAL, dmsp_r);
  DST(diag_Cyano_frac, Cyano_frac);
  DST(diag_Cocco_frac, Cocco_frac);
  DST(diag_Eukar_frac, Eukar_frac);
  DST(diag_diatS, diatS);
  DST(diag_diatN, diatN);
  DST(diag_phytoN, phytoN);
  DST(diag_coccoS, coccoS);
  DST(diag_cyanoS, cyanoS);
  DST(diag_eukarS, eukarS);
  DST(diag_diazS, diazS);
  DST(diag_phaeoS, phaeoS);
  DST(diag_zooS, zooS);
  DST(diag_zooCC, zooC);
  DST(diag_RSNzoo, Rs2n_zoo);
}
#undef DST

// DMS_output%DMS_tendencies = 0 (DMS_mod.F90:413) and the two live slots of an active cell.
__device__ __forceinline__ void dms_store_tendencies(const DmsArgs &A, unsigned i2, unsigned nLnC, bool active,
                                                     double t_dms, double t_dmsp) {
  const DmsIndices &I = c_dms.ind;
#pragma unroll
  for (int n = 0; n < DMS_TRACER_CNT; ++n) {
    double v = 0.0;
    if (active && n == I.dms_ind - 1) v = t_dms;
    if (active && n == I.dmsp_ind - 1) v = t_dmsp;
    A.tend[i2 + (unsigned)n * nLnC] = v;
  }
}

// Tile kernel: a block owns kDmsTileCols consecutive columns over ALL levels; warp w takes the
// levels w, w + W, ...  The only vertical coupling of DMS_SourceSink is the PAR attenuation
// product (:510-527), so
//   phase 1  every cell's KPARdz and bexp(-KPARdz) -> shared memory (cell-parallel),
//   phase 2  warp 0 walks its 32 columns top to bottom, PAR_out = PAR_in * bexp(-KPARdz) in the
//            reference's order (bit-identical to the sequential sweep), PAR_in -> shared memory,
//   phase 3  every cell is independent: 9 tracer loads, 14 + 27 stores.
// The mesh offers nL times more parallelism this way than one thread per column, which is
// what an HBM-bound streaming kernel needs to keep enough bytes in flight.
constexpr int kDmsTileCols = 32;
constexpr int kDmsTileWarps = 8;

template <bool ALLDIAG, int MINB>
__global__ void __launch_bounds__(kDmsTileCols * kDmsTileWarps, MINB)
dms_cells_kernel(const __grid_constant__ DmsArgs A) {
  extern __shared__ double dsm[];
  __shared__ double red[kDmsTileWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * kDmsTileCols + lane;
  const int nL = A.nL, nC = A.nC;
  const bool in_range = col < nC;
  const unsigned nLnC = (unsigned)nL * (unsigned)nC;   // 32-bit element indices: see k_eco.cu
  int kmax = (in_range && col < A.nColumns) ? A.kmax[col] : 0;
  if (kmax > nL) kmax = nL;
  if (kmax < 0) kmax = 0;
  double *const s_kp = dsm;                       // [nL][32] KPARdz
  double *const s_ek = dsm + (size_t)nL * 32;     // [nL][32] exp(-KPARdz)
  double *const s_pin = dsm + (size_t)nL * 64;    // [nL][32] PAR_in
  const DmsIndices &I = c_dms.ind;

  double SST_loc = 0.0;
  if (kmax > 0) SST_loc = A.sst[col];
  const DmsColumnConsts cc = dms_column_consts(SST_loc);

  {   // phase 1, software-pipelined like phase 3: the next level's five loads are in flight
    struct ChlIn { double a, b, c, d, dz; };
    auto load_chl = [&](int k) {
      const unsigned i2 = (unsigned)col + (unsigned)nC * (unsigned)k;
#define TR(ind_) A.tracers[i2 + (unsigned)((ind_) - 1) * nLnC]
      ChlIn r = {TR(I.spChl_ind), TR(I.diatChl_ind), TR(I.diazChl_ind), TR(I.phaeoChl_ind), A.dz[i2]};
#undef TR
      return r;
    };
    ChlIn cur = {}, nxt = {};
    if (w < kmax) cur = load_chl(w);
    for (int k = w; k < kmax; k += kDmsTileWarps) {
      if (k + kDmsTileWarps < kmax) nxt = load_chl(k + kDmsTileWarps);
      if (A.l2_prefetch && k + A.l2_prefetch * kDmsTileWarps < kmax) {
        const unsigned j2 = (unsigned)col + (unsigned)nC * (unsigned)(k + A.l2_prefetch * kDmsTileWarps);
        pf_l2(A.tracers + (j2 + (unsigned)(I.spChl_ind - 1) * nLnC)); pf_l2(A.tracers + (j2 + (unsigned)(I.diatChl_ind - 1) * nLnC));
        pf_l2(A.tracers + (j2 + (unsigned)(I.diazChl_ind - 1) * nLnC)); pf_l2(A.tracers + (j2 + (unsigned)(I.phaeoChl_ind - 1) * nLnC));
        pf_l2(A.dz + j2);
      }
      const double totalChl = gmax(0.0, cur.a) + gmax(0.0, cur.b) + gmax(0.0, cur.c) + gmax(0.0, cur.d);
      double kp, ek;
      dms_attenuation(totalChl, cur.dz, kp, ek);
      s_kp[k * 32 + lane] = kp;
      s_ek[k * 32 + lane] = ek;
      cur = nxt;
    }
  }
  __syncthreads();
  if (w == 0 && kmax > 0) {
    double PAR = gmax(0.0, A.sw_flux[col]);
    PAR = PAR * c_dms.p.f_qsw_par_DMS;
    for (int k = 0; k < kmax; ++k) {
      s_pin[k * 32 + lane] = PAR;
      PAR = PAR * s_ek[k * 32 + lane];
    }
  }
  __syncthreads();

  double inv_dms = 0.0, inv_dmsp = 0.0;   // sum over this thread's cells of tendency * dz (inventory)
  if (in_range) {
    // software pipeline: the next cell's nine loads are in flight while this one is computed
    DmsCellIn cur = {}, nxt = {};
    if (w < kmax) cur = dms_load_cell(A, (unsigned)col + (unsigned)nC * (unsigned)w, nLnC);
    for (int k = w; k < nL; k += kDmsTileWarps) {
      const unsigned i2 = (unsigned)col + (unsigned)nC * (unsigned)k;
      const int kn = k + kDmsTileWarps;
      if (kn < kmax) nxt = dms_load_cell(A, i2 + (unsigned)nC * (unsigned)kDmsTileWarps, nLnC);
      if (A.l2_prefetch && k + A.l2_prefetch * kDmsTileWarps < kmax)
        dms_prefetch_cell(A, i2 + (unsigned)nC * (unsigned)(A.l2_prefetch * kDmsTileWarps), nLnC);
      const bool active = k < kmax;
      double t_dms = 0.0, t_dmsp = 0.0;
      if (active) {   // diagnostics keep their previous contents outside active cells
        const double PAR_avg = fdiv(s_pin[k * 32 + lane] * (1.0 - s_ek[k * 32 + lane]), s_kp[k * 32 + lane]);
        dms_cell<ALLDIAG>(A, i2, cur, PAR_avg, cc, t_dms, t_dmsp);
        inv_dms += t_dms * cur.dz;     // cur.dz is 0 without the inventory
        inv_dmsp += t_dmsp * cur.dz;
      }
      dms_store_tendencies(A, i2, nLnC, active, t_dms, t_dmsp);
      cur = nxt;
    }
  }
  if (A.inv_partials) {   // stage 1 of the inventory reduction, fused: one partial per block
    const double a = block_sum(inv_dms, red), b = block_sum(inv_dmsp, red);
    if (threadIdx.x == 0) {
      double *out = A.inv_partials + (size_t)blockIdx.x * kInvGroup;
      out[0] = a; out[1] = b;
#pragma unroll
      for (int j = 2; j < kInvGroup; ++j) out[j] = 0.0;
    }
  }
}

// Column kernel: one thread per column, levels in order, PAR carried down the column in a register
// like the reference's own loop.  All blocks of a wave walk the levels in step, so the chip works on a
// few levels of each array at a time, in 2-KB runs, where the tile kernel's unsynchronised blocks touch
// every level at once in 256-byte pieces.  The tile kernel pays for that with the size of the arrays
// (measured, round 2, ns per cell: 0.083 at 235 160 x 60, 0.089 at 235 160 x 80, 0.102 at 461 654 x 60,
// 0.105 at 461 654 x 80), this kernel does not (0.090 at 235 160 x 60, 0.093 at 461 654 x 80) but needs
// several waves of columns to fill the chip: launch_dms_columns picks it for large blocks of columns and
// for level counts whose tile does not fit shared memory.  (Register prefetch of the next level and L2
// prefetch of the one after made it slower on the large mesh, 3.76 against 3.43 ms: whatever the tile
// kernel exhausts there does not like more requests in flight either.)
template <bool ALLDIAG>
__global__ void __launch_bounds__(256, 2)
dms_columns_kernel(const __grid_constant__ DmsArgs A) {
  __shared__ double red[256 / 32];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int nL = A.nL, nC = A.nC;
  const bool in_range = col < nC;
  const unsigned nLnC = (unsigned)nL * (unsigned)nC;
  int kmax = (in_range && col < A.nColumns) ? A.kmax[col] : 0;
  if (kmax > nL) kmax = nL;
  if (kmax < 0) kmax = 0;
  double inv_dms = 0.0, inv_dmsp = 0.0;
  const DmsIndices &I = c_dms.ind;

  double SST_loc = 0.0, PAR_out = 0.0;
  if (kmax > 0) {
    SST_loc = A.sst[col];
    PAR_out = gmax(0.0, A.sw_flux[col]);
    PAR_out = PAR_out * c_dms.p.f_qsw_par_DMS;
  }
  const DmsColumnConsts cc = dms_column_consts(SST_loc);

  for (int k = 0; in_range && k < nL; ++k) {
    const unsigned i2 = (unsigned)col + (unsigned)nC * (unsigned)k;
    const bool active = k < kmax;
    double t_dms = 0.0, t_dmsp = 0.0;
    if (active) {
#define TR(ind_) gmax(0.0, A.tracers[i2 + (unsigned)((ind_) - 1) * nLnC])
      const double totalChl = TR(I.spChl_ind) + TR(I.diatChl_ind) + TR(I.diazChl_ind) + TR(I.phaeoChl_ind);
#undef TR
      const double dz = A.dz[i2];
      double KPARdz, eK;
      dms_attenuation(totalChl, dz, KPARdz, eK);
      const double PAR_in = PAR_out;
      PAR_out = PAR_in * eK;
      const double PAR_avg = fdiv(PAR_in * (1.0 - eK), KPARdz);
      dms_cell<ALLDIAG>(A, i2, dms_load_cell(A, i2, nLnC), PAR_avg, cc, t_dms, t_dmsp);
      inv_dms += t_dms * dz;
      inv_dmsp += t_dmsp * dz;
    }
    dms_store_tendencies(A, i2, nLnC, active, t_dms, t_dmsp);
  }
  if (A.inv_partials) {
    const double a = block_sum(inv_dms, red), b = block_sum(inv_dmsp, red);
    if (threadIdx.x == 0) {
      double *out = A.inv_partials + (size_t)blockIdx.x * kInvGroup;
      out[0] = a; out[1] = b;
#pragma unroll
      for (int j = 2; j < kInvGroup; ++j) out[j] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(256)
dms_surface_kernel(const __grid_constant__ DmsSurfArgs A) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= A.nColumns) return;
  const size_t nC = (size_t)A.nC;
  const size_t nLnC = (size_t)A.nL * nC;
  const DmsIndices &I = c_dms.ind;
  constexpr double a = 0.31, e2 = 2.85, e3 = 0.612;   // DMS_mod.F90:831-838

  const double seaSurfaceDMS = gmax(0.0, A.tracers[(size_t)col + (size_t)(I.dms_ind - 1) * nLnC]);
  const double sst = A.f.SST[col];
  double ice = A.f.iceFraction[col];
  if (ice < 0.0) ice = 0.0;
  if (ice > 1.0) ice = 1.0;
  A.f.iceFraction[col] = ice;   // in-place clamp (:858-859)

  const double sc = 2674.0 + sst * (-147.12 + sst * (3.726 + sst * (-0.038)));   // Kettle & Andreae 2000 (:915-959)
  const double ws = sqrt(fabs(A.f.windSpeedSquared10m[col])) * 0.01;            // cm/s -> m/s (:866)

  const double XKW_W92 = a * (pow((660.0 / sc), 0.500)) * ws * ws;
  const double XKW_LM86 = e2 * (pow((600.0 / sc), 0.500)) * (ws - 3.6) + e3 * (pow((600.0 / sc), 0.667));
  double xkw = 0.0;
  if (ws < 3.6) xkw = XKW_W92;
  if ((ws >= 3.6) && (ws < 5.6)) {
    const double FLM86 = 0.5 * (ws - 3.6);
    const double FW92 = 1.0 - FLM86;
    xkw = FW92 * XKW_W92 + FLM86 * XKW_LM86;
  }
  if (ws >= 5.6) xkw = XKW_LM86;
  xkw = xkw / 3600.0;
  const double xkw_ice = (1.0 - ice) * xkw;

  const double DMSSAT_1atm = 0.0;   // DMSSAT_singleValue is identically zero (:1003)
  const double pv = xkw_ice * sqrt(660.0 / sc);
  const double pres = A.f.surfacePressure[col];
  const double sat = pres * DMSSAT_1atm;
  A.f.netFlux[(size_t)col + (size_t)(I.dms_ind - 1) * nC] = pv * (sat - seaSurfaceDMS);
  A.f.netFlux[(size_t)col + (size_t)(I.dmsp_ind - 1) * nC] = 0.0;

#define DG(name, val) do { if (A.d.name) A.d.name[col] = (val); } while (0)
  DG(diag_DMS_IFRAC, ice);
  DG(diag_DMS_XKW, xkw_ice);
  DG(diag_DMS_ATM_PRESS, pres);
  DG(diag_DMS_PV, pv);
  DG(diag_DMS_SCHMIDT, sc);
  DG(diag_DMS_SAT, sat);
  DG(diag_DMS_SURF, seaSurfaceDMS);
  DG(diag_DMS_WS, ws);
#undef DG
}

constexpr int kMacrosBlock = 1024;   // few, large blocks: one inventory partial per block

__global__ void __launch_bounds__(kMacrosBlock)
macros_cells_kernel(const __grid_constant__ MacrosArgs A) {
  const size_t nC = (size_t)A.nC;
  const size_t ncell = (size_t)A.nL * nC;
  __shared__ double red[kMacrosBlock / 32];
  const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = cell < ncell;
  // (32-bit division: nL * nC < 2^32 / 30, bgc_capi.cu check_dims; the 64-bit one costs ~30 instructions)
  const int k = in_range ? (int)((unsigned)cell / (unsigned)nC) : 0;
  const int col = in_range ? (int)(cell - (size_t)k * nC) : 0;
  const bool active = in_range && col < A.nColumns && k < A.kmax[col];
  const MacrosParams &P = c_macros.p;
  const MacrosIndices &I = c_macros.ind;

  double t_prot = 0.0, t_poly = 0.0, t_lip = 0.0;
  if (active) {
#define TR(ind_) gmax(0.0, A.tracers[cell + (size_t)((ind_) - 1) * ncell])
    const double zooC = TR(I.zooC_ind), spC = TR(I.spC_ind), diatC = TR(I.diatC_ind), diazC = TR(I.diazC_ind),
                 phaeoC = TR(I.phaeoC_ind), prot = TR(I.prot_ind), poly = TR(I.poly_ind), lip = TR(I.lip_ind);
#undef TR
    const double k_C_p = P.k_C_p_base * (P.mort + fdiv(zooC, P.zooC_avg));
    const double phytoC = diatC + phaeoC + spC + diazC;
    const double prot_s = P.inject_scale * P.f_prot * k_C_p * phytoC;
    const double poly_s = P.inject_scale * P.f_poly * k_C_p * phytoC;
    const double lip_s = P.inject_scale * P.f_lip * k_C_p * phytoC;
    const double prot_r = P.k_prot_bac * prot;
    const double poly_r = P.k_poly_bac * poly;
    const double lip_r = P.k_lip_bac * lip;
    t_prot = prot_s - prot_r;
    t_poly = poly_s - poly_r;
    t_lip = lip_s - lip_r;
    if (A.d.diag_PROT_S_TOTAL) A.d.diag_PROT_S_TOTAL[cell] = prot_s;
    if (A.d.diag_POLY_S_TOTAL) A.d.diag_POLY_S_TOTAL[cell] = poly_s;
    if (A.d.diag_LIP_S_TOTAL) A.d.diag_LIP_S_TOTAL[cell] = lip_s;
    if (A.d.diag_PROT_R_TOTAL) A.d.diag_PROT_R_TOTAL[cell] = prot_r;
    if (A.d.diag_POLY_R_TOTAL) A.d.diag_POLY_R_TOTAL[cell] = poly_r;
    if (A.d.diag_LIP_R_TOTAL) A.d.diag_LIP_R_TOTAL[cell] = lip_r;
  }
  // MACROS_tendencies = 0 everywhere (:267), three live slots on active cells
#pragma unroll
  for (int n = 0; n < MACROS_TRACER_CNT; ++n) {
    double v = 0.0;
    if (n == I.prot_ind - 1) v = t_prot;
    if (n == I.poly_ind - 1) v = t_poly;
    if (n == I.lip_ind - 1) v = t_lip;
    if (in_range) A.tend[cell + (size_t)n * ncell] = v;
  }
  if (A.inv_partials) {   // stage 1 of the inventory reduction, fused (tendencies are 0 off active cells)
    const double dz = active ? A.dz[cell] : 0.0;
    const double a = block_sum(t_prot * dz, red), b = block_sum(t_poly * dz, red), c = block_sum(t_lip * dz, red);
    if (threadIdx.x == 0) {
      double *out = A.inv_partials + (size_t)blockIdx.x * kInvGroup;
      out[0] = a; out[1] = b; out[2] = c;
#pragma unroll
      for (int j = 3; j < kInvGroup; ++j) out[j] = 0.0;
    }
  }
}

// ---------------------------------------------------------------- layout
// dst(c, r) = src(r, c) for each of nSlabs 2-D slabs; src has `R` fastest.
// 32x32 FP64 tile through padded shared memory: both sides coalesced.
__global__ void __launch_bounds__(256)
transpose_kernel(const double *__restrict__ src, double *__restrict__ dst, int R, int C, int long_axis_is_c) {
  __shared__ double tile[32][33];
  const size_t slab = (size_t)blockIdx.z * (size_t)R * (size_t)C;
  // the longer axis (millions of columns) rides on gridDim.x, which has no 65535 limit
  const int r0 = (long_axis_is_c ? blockIdx.y : blockIdx.x) * 32, c0 = (long_axis_is_c ? blockIdx.x : blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + tx, c = c0 + ty + j;
    if (r < R && c < C) tile[ty + j][tx] = src[slab + (size_t)c * R + r];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + tx, r = r0 + ty + j;
    if (r < R && c < C) dst[slab + (size_t)r * C + c] = tile[tx][ty + j];
  }
}

// ---------------------------------------------------------------- MPAS tracer layout
// One block = KB levels x 32 cells x all tracers through a padded shared-memory tile.  In the MPAS
// array the (tracer, level) pairs of one cell are contiguous, so the block moves 32 runs of
// KB*nT doubles on that side (~2 KB each) and runs of 32 consecutive cells on the SoA side.
template <bool TO_SOA>
__global__ void __launch_bounds__(256)
mpas_layout_kernel(const double *__restrict__ src, double *__restrict__ dst, const __grid_constant__ MpasMap M,
                   int nL, int nC, int KB, double alpha, double beta, const double *__restrict__ weight, int ktiles) {
  extern __shared__ double tile[];   // [32][KB*nT + 1]
  // 1-D grid, level blocks of one cell block first: the blocks that are resident at the same time
  // work on the same cells' contiguous MPAS runs (DRAM page locality on the strided side)
  const int cb = blockIdx.x / ktiles, kb = blockIdx.x - cb * ktiles;
  const int c0 = cb * 32, k0 = kb * KB, nT = M.nT;
  const int ncell = min(32, nC - c0), nk = min(KB, nL - k0);
  const int run = nk * nT, pitch = KB * nT + 1;
  const size_t nLnC = (size_t)nL * (size_t)nC;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // MPAS side: warp w takes cells w, w + nw, ...; its lanes walk the cell's contiguous run.
  // SoA side: warp w takes (level, tracer) pairs; its lanes are the 32 cells.  No divisions.
  // The slot map is consulted per TRACER, not per element (a per-element lookup is an indexed
  // constant load through the MIO pipe, and that pipe - not HBM - then bounds the kernel: measured
  // 3100 thread instructions per cell and mio_throttle as the top stall before this was hoisted);
  // on the MPAS side, where a lane walks a run of all tracers, membership is one bit of M.used.
  if (TO_SOA) {
    for (int c = w; c < ncell; c += nw) {
      const double *p = src + (size_t)nT * ((size_t)k0 + (size_t)nL * (size_t)(c0 + c));
      for (int r = lane; r < run; r += 32) tile[c * pitch + r] = p[r];
    }
    __syncthreads();
    if (lane < ncell) {
      for (int n = w; n < nT; n += nw) {
        const int sl = M.slot[n];
        if (sl <= 0) continue;
        double *q = dst + (size_t)(c0 + lane) + (size_t)nC * (size_t)k0 + (size_t)(sl - 1) * nLnC;
        const double *t = tile + lane * pitch + n;
        for (int kk = 0; kk < nk; ++kk, q += nC, t += nT) *q = *t;
      }
    }
  } else {
    if (lane < ncell) {
      for (int n = w; n < nT; n += nw) {
        const int sl = M.slot[n];
        if (sl <= 0) continue;
        const double *q = src + (size_t)(c0 + lane) + (size_t)nC * (size_t)k0 + (size_t)(sl - 1) * nLnC;
        double *t = tile + lane * pitch + n;
        double v[8];
        for (int kk0 = 0; kk0 < nk; kk0 += 8) {   // loads first, then the shared-memory stores: up to 8 in flight
#pragma unroll
          for (int j = 0; j < 8; ++j) if (kk0 + j < nk) v[j] = q[(size_t)(kk0 + j) * nC];
#pragma unroll
          for (int j = 0; j < 8; ++j) if (kk0 + j < nk) t[(kk0 + j) * nT] = v[j];
        }
      }
    }
    __syncthreads();
    const unsigned long long used = M.used;
    const bool rmw = beta != 0.0;
    for (int c = w; c < ncell; c += nw) {
      double *p = dst + (size_t)nT * ((size_t)k0 + (size_t)nL * (size_t)(c0 + c));
      // weight(k, cell), level fastest (MPAS layerThickness): the thickness-weighted tendency
      const double *wp = weight ? weight + (size_t)k0 + (size_t)nL * (size_t)(c0 + c) : nullptr;
      for (int r = lane, n = lane % nT, kk = lane / nT; r < run; r += 32) {
        if ((used >> n) & 1ull) {
          double v = tile[c * pitch + r];
          if (wp) v = wp[kk] * v;
          p[r] = rmw ? beta * p[r] + alpha * v : alpha * v;
        }
        n += 32;
        while (n >= nT) { n -= nT; ++kk; }
      }
    }
  }
}

// Pipelined form of the same tile scheme: a persistent grid (a few blocks per SM) walks the
// (cell block, level block) tiles, level blocks of one cell block first, and keeps TWO tiles in
// shared memory - while the warps write tile i out, tile i + 1 is already landing through
// cp.async (LDGSTS: no registers, no warp waiting).  The one-shot kernel above pays a full
// HBM round trip between its load phase and its store phase in every block; here the only
// exposed latency is the first tile of each block.  VEC = 2: 16-byte copies (even nT and a 16-byte
// aligned array: every run starts aligned and holds whole pairs), VEC = 1: 8-byte copies.
__device__ __forceinline__ void cp_async_8(double *dst, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(double *dst, const double *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <bool TO_SOA, int VEC>
__global__ void __launch_bounds__(256)
mpas_layout_pipe_kernel(const double *__restrict__ src, double *__restrict__ dst, const __grid_constant__ MpasMap M,
                        int nL, int nC, int KB, double alpha, double beta, const double *__restrict__ weight,
                        int ktiles, int ntiles) {
  extern __shared__ __align__(16) double sm[];
  const int nT = M.nT;
  const int pitch = KB * nT + 2 + ((KB * nT) & 1);          // even (16-byte rows), 2 mod 4 mostly: 2-way conflicts at worst
  const int tile_doubles = 32 * pitch;
  // TO_SOA: [2][32][pitch] = the MPAS runs.  !TO_SOA: per buffer two planes, the SoA values
  // (transposed into run order) and, when beta != 0, the old MPAS runs.
  const bool rmw = !TO_SOA && beta != 0.0;
  const int planes = rmw ? 2 : 1;
  const size_t nLnC = (size_t)nL * (size_t)nC;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;

  auto issue = [&](int tile, int buf) {
    const int cb = tile / ktiles, kb = tile - cb * ktiles;
    const int c0 = cb * 32, k0 = kb * KB;
    const int ncell = min(32, nC - c0), nk = min(KB, nL - k0);
    const int run = nk * nT;
    double *base = sm + (size_t)buf * planes * tile_doubles;
    if (TO_SOA || rmw) {   // the cells' contiguous MPAS runs
      double *plane = base + (TO_SOA ? 0 : tile_doubles);   // (rmw: plane 1; TO_SOA: plane 0)
      const double *g = (TO_SOA ? src : dst);
      for (int c = w; c < ncell; c += nw) {
        const double *p = g + (size_t)nT * ((size_t)k0 + (size_t)nL * (size_t)(c0 + c));
        double *q = plane + c * pitch;
        if (VEC == 2) { for (int r = 2 * lane; r < run; r += 64) cp_async_16(q + r, p + r); }
        else          { for (int r = lane; r < run; r += 32) cp_async_8(q + r, p + r); }
      }
    }
    if (!TO_SOA && lane < ncell) {   // SoA rows: lanes are the 32 cells, 8-byte copies into run order
      for (int n = w; n < nT; n += nw) {
        const int sl = M.slot[n];
        if (sl <= 0) continue;
        const double *q = src + (size_t)(c0 + lane) + (size_t)nC * (size_t)k0 + (size_t)(sl - 1) * nLnC;
        double *t = base + lane * pitch + n;
        for (int kk = 0; kk < nk; ++kk, q += nC, t += nT) cp_async_8(t, q);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = blockIdx.x, buf = 0;
  if (tile < ntiles) issue(tile, 0);
  while (tile < ntiles) {
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      issue(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int cb = tile / ktiles, kb = tile - cb * ktiles;
    const int c0 = cb * 32, k0 = kb * KB;
    const int ncell = min(32, nC - c0), nk = min(KB, nL - k0);
    const int run = nk * nT;
    const double *base = sm + (size_t)buf * planes * tile_doubles;
    if (TO_SOA) {
      if (lane < ncell) {   // (level outer, tracer inner: measured 20 % faster than the hoisted form here)
        for (int kk = 0; kk < nk; ++kk)
          for (int n = w; n < nT; n += nw)
            if (M.slot[n] > 0)
              dst[(size_t)(c0 + lane) + (size_t)nC * (size_t)(k0 + kk) + (size_t)(M.slot[n] - 1) * nLnC] =
                  base[lane * pitch + kk * nT + n];
      }
    } else {
      const unsigned long long used = M.used;
      const double *old = base + tile_doubles;
      for (int c = w; c < ncell; c += nw) {
        double *p = dst + (size_t)nT * ((size_t)k0 + (size_t)nL * (size_t)(c0 + c));
        const double *wp = weight ? weight + (size_t)k0 + (size_t)nL * (size_t)(c0 + c) : nullptr;
        for (int r = lane, n = lane % nT, kk = lane / nT; r < run; r += 32) {
          if ((used >> n) & 1ull) {
            double v = base[c * pitch + r];
            if (wp) v = wp[kk] * v;
            p[r] = rmw ? beta * old[c * pitch + r] + alpha * v : alpha * v;
          }
          n += 32;
          while (n >= nT) { n -= nT; ++kk; }
        }
      }
    }
    __syncthreads();   // the buffer just consumed is refilled by the next iteration's issue
    tile = next;
    buf ^= 1;
  }
}


// SoA -> MPAS for WHOLE columns: a block owns CB cells over all levels, gathers their SoA values
// into shared memory in MPAS order and then streams each cell's contiguous nL*nT run - the
// MPAS side moves in 14-KB runs instead of tile-sized pieces.
template <int CB>
__global__ void __launch_bounds__(256)
mpas_columns_kernel(const double *__restrict__ src, double *__restrict__ dst, const __grid_constant__ MpasMap M,
                    int nL, int nC, double alpha, double beta, const double *__restrict__ weight) {
  extern __shared__ double tile[];   // [CB][nL*nT + 1]
  const int nT = M.nT, run = nL * nT, pitch = run + 1;
  const int c0 = blockIdx.x * CB;
  const int ncell = min(CB, nC - c0);
  const size_t nLnC = (size_t)nL * (size_t)nC;
  constexpr int PER = 32 / CB;                      // (level, tracer) pairs one warp instruction covers
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int cell = lane % CB, sub = lane / CB;
  for (int k = w; k < nL; k += nw) {
    for (int n = sub; n < nT; n += PER) {
      if (cell < ncell && M.slot[n] > 0)
        tile[cell * pitch + k * nT + n] =
            src[(size_t)(c0 + cell) + (size_t)nC * (size_t)k + (size_t)(M.slot[n] - 1) * nLnC];
    }
  }
  __syncthreads();
  const bool rmw = beta != 0.0;
  for (int c = w; c < ncell; c += nw) {
    double *p = dst + (size_t)run * (size_t)(c0 + c);
    const double *wp = weight ? weight + (size_t)nL * (size_t)(c0 + c) : nullptr;
    int n = lane % nT, kk = lane / nT;
#pragma unroll 4
    for (int r = lane; r < run; r += 32) {
      if ((M.used >> n) & 1ull) {
        double v = tile[c * pitch + r];
        if (wp) v = wp[kk] * v;
        p[r] = rmw ? beta * p[r] + alpha * v : alpha * v;
      }
      n += 32;
      while (n >= nT) { n -= nT; ++kk; }
    }
  }
}

// ---------------------------------------------------------------- diagnostics accumulation
__global__ void __launch_bounds__(256)
accumulate_kernel(const double *__restrict__ src, double *__restrict__ acc, int nL, int cc, int nC, int c0,
                  const int *__restrict__ kmax, int nColumns, double w) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y, slab = blockIdx.z;
  if (col >= cc) return;
  if (kmax && !(col < nColumns && k < kmax[col])) return;
  const size_t is = (size_t)col + (size_t)cc * ((size_t)k + (size_t)nL * slab);
  const size_t ia = (size_t)(c0 + col) + (size_t)nC * ((size_t)k + (size_t)nL * slab);
  acc[ia] += w * src[is];
}
__global__ void __launch_bounds__(256)
scale_kernel(double *a, size_t n, double w) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] *= w;
}

// ---------------------------------------------------------------- bottom-cell gather
// out[j][col] = src[j](kmax(col) - 1, col): the only element of a column that the nine sediment
// diagnostics of BGC_SourceSink can set (BGC_mod.F90:2522-2631 run in the bottom cell alone; every
// other element is the zero fill of :625-727).  Host-layout calls download these vectors instead of
// the (k,col) slabs (bgc_capi.cu).
__global__ void __launch_bounds__(256)
bottom_gather_kernel(BottomGatherArgs a) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.cc) return;
  int km = col < a.nColumns ? a.kmax[col] : 0;
  if (km > a.nL) km = a.nL;
  for (int j = 0; j < a.n; ++j)
    a.out[(size_t)j * a.cc + col] = km > 0 ? a.src[j][(size_t)(km - 1) * a.cc + col] : 0.0;
}

// ---------------------------------------------------------------- inventory
// Stage 1 of the inventory reduction is fused into the source-sink kernels: every block
// writes its partial sums, [nParts][nGroups][kInvGroup].  Stage 2 below adds the partials
// of each value in a fixed order into the inventory vector: block = group, 128 rows of
// kInvGroup threads, row r walks partials r, r + 128, ... with kFoldDepth loads in flight.
constexpr int kFoldRows = 128;
constexpr int kFoldDepth = 16;

// NG = number of groups when known at compile time (0: gridDim.x): the kFoldDepth loads of a trip then
// differ by immediates from one base address - with a run-time stride each needs its own 64-bit address
// register, and at 1024 threads (64 registers) ptxas keeps two loads in flight instead of sixteen.
template <int NG>
__global__ void __launch_bounds__(kFoldRows * kInvGroup)
inventory_fold_kernel(const __grid_constant__ InventoryFoldArgs A, int nParts) {
  __shared__ double rows[kFoldRows][kInvGroup];
  const int g = blockIdx.x, j = threadIdx.x % kInvGroup, r = threadIdx.x / kInvGroup;
  // One block per group is latency-bound (a DRAM round trip per dependent load): kFoldDepth loads of a
  // row are in flight at once, the tail is predicated instead of walked one load at a time.
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const unsigned stride = (NG ? NG : gridDim.x) * kInvGroup;
  const double *p = A.partials + (size_t)g * kInvGroup + j;
  for (int b = r; b < nParts; b += kFoldDepth * kFoldRows) {
    const double *pb = p + (size_t)((unsigned)b * stride);
    double v[kFoldDepth];
#pragma unroll
    for (int u = 0; u < kFoldDepth; ++u) {
      v[u] = 0.0;
      if (b + u * kFoldRows < nParts) v[u] = __ldg(pb + (size_t)(u * kFoldRows) * stride);
    }
#pragma unroll
    for (int u = 0; u < kFoldDepth; u += 4) { s0 += v[u]; s1 += v[u + 1]; s2 += v[u + 2]; s3 += v[u + 3]; }
  }
  rows[r][j] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (r == 0) {
    double t = 0.0;
    for (int q = 0; q < kFoldRows; ++q) t += rows[q][j];
    const int dst = A.out_index[g][j];
    if (dst >= 0) A.inventory[dst] += t;
  }
}

}  // namespace

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// shared memory of the tile kernel: three [nL][32] FP64 planes
static size_t dms_tile_smem(int nL) { return (size_t)nL * 32 * 3 * sizeof(double); }
static bool dms_use_tiles(int nL) { return dms_tile_smem(nL) <= 160 * 1024; }

template <bool ALLDIAG, int MINB>
static cudaError_t launch_dms_tiles(const DmsArgs &a, cudaStream_t s) {
  const size_t smem = dms_tile_smem(a.nL);
  auto kern = dms_cells_kernel<ALLDIAG, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<cdiv((size_t)a.nC, kDmsTileCols), kDmsTileCols * kDmsTileWarps, smem, s>>>(a);
  return cudaGetLastError();
}

// Which kernel DMS_SourceSink runs as: the tile kernel (nL-fold parallelism) unless its tile does not
// fit shared memory or the mesh has so many columns that the pipelined column kernel fills the chip
// for five waves or more (its whole-column blocks then quantise by less than the tile kernel loses to
// its scattered accesses).  148 SMs x 2 blocks x 256 columns per wave.
// Which kernel DMS_SourceSink runs as: the tile kernel (nL-fold parallelism) unless its tile does not fit
// shared memory or the block has so many columns that the column kernel fills the chip for five waves
// or more (148 SMs x 2 blocks x 256 columns per wave); see dms_columns_kernel.
static bool dms_use_columns(int nL, int nC, int variant) {
  if ((variant & 3) >= 2) return true;    // tuning: column kernel on request
  if ((variant & 3) == 1) return false;   //         tile kernel on request (if it fits)
  return (size_t)nC >= (size_t)5 * 148 * 2 * 256;
}

cudaError_t launch_dms_columns(const DmsArgs &a0, int variant, cudaStream_t s) {
  DmsArgs a = a0;
  if (a.nC <= 0 || a.nL <= 0) return cudaSuccess;
  // distance of the L2 prefetch in trips (2 = the cell after next); variant bits 2.. override it (tuning): 4 = off
  a.l2_prefetch = (variant & 4) ? 0 : ((variant >> 3) ? (variant >> 3) : 2);
  bool all = true;
  double *const *pp = (double *const *)&a.d;
  for (size_t i = 0; i < sizeof(DmsDiagnostics) / sizeof(double *); ++i) all = all && pp[i] != nullptr;
  if (!dms_use_tiles(a.nL) || dms_use_columns(a.nL, a.nC, variant)) {
    if (all) dms_columns_kernel<true><<<cdiv((size_t)a.nC, 256), 256, 0, s>>>(a);
    else     dms_columns_kernel<false><<<cdiv((size_t)a.nC, 256), 256, 0, s>>>(a);
    return cudaGetLastError();
  }
  return all ? launch_dms_tiles<true, 2>(a, s) : launch_dms_tiles<false, 2>(a, s);
}

cudaError_t launch_dms_surface(const DmsSurfArgs &a, cudaStream_t s) {
  if (a.nColumns <= 0) return cudaSuccess;
  dms_surface_kernel<<<cdiv((size_t)a.nColumns, 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_macros_cells(const MacrosArgs &a, cudaStream_t s) {
  const size_t ncell = (size_t)a.nL * (size_t)a.nC;
  if (ncell == 0) return cudaSuccess;
  macros_cells_kernel<<<cdiv(ncell, kMacrosBlock), kMacrosBlock, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_transpose(const double *src, double *dst, int R, int C, int nSlabs, cudaStream_t s) {
  if (R <= 0 || C <= 0 || nSlabs <= 0) return cudaSuccess;
  // gridDim.y/z are limited to 65535: whichever axis is longer goes on x
  const int c_long = C > R ? 1 : 0;
  dim3 grid(cdiv((size_t)(c_long ? C : R), 32), cdiv((size_t)(c_long ? R : C), 32), (unsigned)nSlabs);
  if (grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidConfiguration;
  transpose_kernel<<<grid, 256, 0, s>>>(src, dst, R, C, c_long);
  return cudaGetLastError();
}

// levels per block (measured on B200, 30 tracers: 4 is best towards SoA, 2 for the read-modify-
// write direction; larger tiles lose more to the load/store phase split than they gain in DRAM locality)
static int mpas_levels_per_block(int nT, bool to_soa) {
  const int cap = to_soa ? 4 : 2;
  int kb = (96 * 1024 / 8 / 32 - 1) / nT;
  return kb < 1 ? 1 : (kb > cap ? cap : kb);
}
template <bool TO_SOA>
static cudaError_t launch_mpas_oneshot(const double *src, double *dst, const MpasMap &m, int nL, int nC, double alpha,
                                       double beta, const double *weight, cudaStream_t s) {
  int KB = mpas_levels_per_block(m.nT, TO_SOA);
  if (const char *v = getenv("BGC_MPAS_KB1")) { const int kb = atoi(v); if (kb > 0 && (size_t)32 * (kb * m.nT + 1) * 8 <= 200 * 1024) KB = kb; }
  if (KB > nL) KB = nL;
  const size_t smem = (size_t)32 * (KB * m.nT + 1) * sizeof(double);
  auto kern = mpas_layout_kernel<TO_SOA>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int ktiles = (nL + KB - 1) / KB;
  const long long ntiles = (long long)((nC + 31) / 32) * ktiles;
  if (ntiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)ntiles, 256, smem, s>>>(src, dst, m, nL, nC, KB, alpha, beta, weight, ktiles);
  return cudaGetLastError();
}

// Measured on B200, 30 tracers x 60 levels x 235 160 cells (scripts/micro/mpas_sweep.py,
// profiles/mpas_sweep_r02.txt; ncu: profiles/ncu_r02_mpas_summary.json): the pipelined kernel moves
// 5.86 TB/s towards SoA (0.90 of the copy peak; the one-shot kernel 4.6), 4.62 TB/s for the
// read-modify-write update towards MPAS (0.71; one-shot 3.0) and 4.62 TB/s for the plain conversion.
// Round 1's 2.7 TB/s in the MPAS direction was not a memory limit at all: the slot map was looked up
// per ELEMENT (an indexed constant load through the MIO pipe, 3100 thread instructions per cell,
// mio_throttle the top stall); it is now read once per tracer / tested as one bit of a mask.
// BGC_MPAS_VARIANT (tuning only): 0 = pipelined kernel, 1 = one-shot kernel, 3 / 4 = whole-column
// tiles; BGC_MPAS_KB / BGC_MPAS_BLOCKS_PER_SM override tile and grid.
static int env_int_or(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <bool TO_SOA>
static cudaError_t launch_mpas_layout(const double *src, double *dst, const MpasMap &m, int nL, int nC, double alpha,
                                      double beta, const double *weight, cudaStream_t s) {
  if (nL <= 0 || nC <= 0 || m.nT <= 0) return cudaSuccess;
  if (m.nT > kMpasMaxTracers) return cudaErrorInvalidValue;
  static const int variant = env_int_or("BGC_MPAS_VARIANT", 0);
  if (!TO_SOA && variant >= 3) {   // whole-column tiles (tuning; loses: 64-byte pieces on the SoA side)
    const int CB = variant == 3 ? 8 : 4;
    const size_t smem = (size_t)CB * ((size_t)nL * m.nT + 1) * sizeof(double);
    if (smem <= 227 * 1024) {
      cudaError_t e;
      if (CB == 8) {
        e = cudaFuncSetAttribute(mpas_columns_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mpas_columns_kernel<8><<<cdiv((size_t)nC, 8), 256, smem, s>>>(src, dst, m, nL, nC, alpha, beta, weight);
      } else {
        e = cudaFuncSetAttribute(mpas_columns_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mpas_columns_kernel<4><<<cdiv((size_t)nC, 4), 128, smem, s>>>(src, dst, m, nL, nC, alpha, beta, weight);
      }
      return cudaGetLastError();
    }
  }
  if (variant == 1) return launch_mpas_oneshot<TO_SOA>(src, dst, m, nL, nC, alpha, beta, weight, s);
  static const int kb_env = env_int_or("BGC_MPAS_KB", 0), bps_env = env_int_or("BGC_MPAS_BLOCKS_PER_SM", 0);
  const bool rmw = !TO_SOA && beta != 0.0;
  // measured defaults (profiles/mpas_sweep_r02.txt): towards SoA 4-level tiles, 2 resident blocks per
  // SM; read-modify-write towards MPAS 2-level tiles (two planes per buffer), 3 blocks; plain
  // conversion towards MPAS 4-level tiles, 3 blocks
  int KB = kb_env > 0 ? kb_env : (rmw ? 2 : 4);
  if (KB > nL) KB = nL;
  const int planes = rmw ? 2 : 1;
  auto smem_of = [&](int kb) { return (size_t)2 * planes * 32 * (kb * m.nT + 2 + ((kb * m.nT) & 1)) * sizeof(double); };
  while (KB > 1 && smem_of(KB) > 200 * 1024) --KB;
  const size_t smem = smem_of(KB);
  if (smem > 227 * 1024) return launch_mpas_oneshot<TO_SOA>(src, dst, m, nL, nC, alpha, beta, weight, s);
  // 16-byte copies need every run to start 16-byte aligned on the MPAS side
  const double *mp = TO_SOA ? src : dst;
  const bool vec2 = (((size_t)mp) & 15u) == 0 && (m.nT & 1) == 0;   // even nT: every run is a whole number of 16-byte pairs
  const int ktiles = (nL + KB - 1) / KB;
  const long long ntiles_ll = (long long)((nC + 31) / 32) * ktiles;
  if (ntiles_ll > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  const int ntiles = (int)ntiles_ll;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  const int cap = TO_SOA ? 2 : 3;   // more resident blocks thrash the DRAM pages of the strided side
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  if (bps_env > 0) per_sm = bps_env;
  int grid = sms * per_sm;
  if (grid > ntiles) grid = ntiles;
  cudaError_t e;
  if (vec2) {
    auto kern = mpas_layout_pipe_kernel<TO_SOA, 2>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 256, smem, s>>>(src, dst, m, nL, nC, KB, alpha, beta, weight, ktiles, ntiles);
  } else {
    auto kern = mpas_layout_pipe_kernel<TO_SOA, 1>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 256, smem, s>>>(src, dst, m, nL, nC, KB, alpha, beta, weight, ktiles, ntiles);
  }
  return cudaGetLastError();
}
cudaError_t launch_mpas_to_soa(const double *mpas, double *soa, const MpasMap &m, int nL, int nC, cudaStream_t s) {
  return launch_mpas_layout<true>(mpas, soa, m, nL, nC, 1.0, 0.0, nullptr, s);
}
cudaError_t launch_soa_to_mpas(const double *soa, double *mpas, const MpasMap &m, int nL, int nC, double alpha,
                               double beta, const double *weight, cudaStream_t s) {
  return launch_mpas_layout<false>(soa, mpas, m, nL, nC, alpha, beta, weight, s);
}

cudaError_t launch_accumulate(const double *src, double *acc, int nL, int cc, int nC, int c0, int nSlabs,
                              const int *kmax, int nColumns, double w, cudaStream_t s) {
  if (nL <= 0 || cc <= 0 || nSlabs <= 0) return cudaSuccess;
  dim3 grid(cdiv((size_t)cc, 256), (unsigned)nL, (unsigned)nSlabs);
  if (grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidConfiguration;
  accumulate_kernel<<<grid, 256, 0, s>>>(src, acc, nL, cc, nC, c0, kmax, nColumns, w);
  return cudaGetLastError();
}
cudaError_t launch_bottom_gather(const BottomGatherArgs &a, cudaStream_t s) {
  if (a.cc <= 0 || a.n <= 0) return cudaSuccess;
  bottom_gather_kernel<<<cdiv((size_t)a.cc, 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_scale(double *a, size_t n, double w, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  scale_kernel<<<cdiv(n, 256), 256, 0, s>>>(a, n, w);
  return cudaGetLastError();
}

cudaError_t launch_inventory_fold(const InventoryFoldArgs &a, int nParts, cudaStream_t s) {
  if (nParts <= 0 || a.nGroups <= 0) return cudaSuccess;
  if (a.nGroups > kInvMaxGroups) return cudaErrorInvalidValue;
  if (a.nGroups == 1) inventory_fold_kernel<1><<<1, kFoldRows * kInvGroup, 0, s>>>(a, nParts);
  else if (a.nGroups == kEcoInvGroups) inventory_fold_kernel<kEcoInvGroups><<<a.nGroups, kFoldRows * kInvGroup, 0, s>>>(a, nParts);
  else inventory_fold_kernel<0><<<a.nGroups, kFoldRows * kInvGroup, 0, s>>>(a, nParts);
  return cudaGetLastError();
}
// blocks (= inventory partials) of the kernel launch_dms_columns picks for this shape and variant
int dms_inventory_parts(int nL, int nC, int variant) {
  if (!dms_use_tiles(nL) || dms_use_columns(nL, nC, variant)) return (nC + 255) / 256;
  return (nC + kDmsTileCols - 1) / kDmsTileCols;
}
int macros_inventory_parts(int nL, int nC) { return (int)(((size_t)nL * (size_t)nC + kMacrosBlock - 1) / kMacrosBlock); }

}  // namespace bgc
