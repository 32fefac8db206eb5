// k_dms.cu — DMS_SourceSink and DMS_SurfaceFluxes kernels for sm_100a.
//
//   dms_cells_kernel      <- DMS_SourceSink     DMS_mod.F90:156-770   (block = 32 columns x all levels, thread = cell)
//   dms_columns_kernel    <- DMS_SourceSink     the same routine, one thread per column (large blocks of columns)
//   dms_surface_kernel    <- DMS_SurfaceFluxes  DMS_mod.F90:778-908   (thread = column)
//
// This file is compiled with -fmad=false in BOTH flavours of the library (Makefile).  DMS_SourceSink
// exists as two kernels, and which one a block of columns gets depends on its width; with FMA
// contraction left to the compiler the two instantiations of the same source contracted different
// operations and differed in the last bit (test_dms_kernel_choice_does_not_change_the_bits).  The
// kernels are HBM-bound (FP64 pipe 13 % active): the handful of separate multiplies and adds costs
// nothing measurable, and the results move towards the reference's, which has no FMA either.
// Each input element is read once and each output element written once, coalesced.
#include <stdlib.h>
#include "bgc_kernels.cuh"
#include "bgc_math.cuh"
#include "bgc_reduce.cuh"

namespace bgc {

__constant__ DmsTables c_dms;

cudaError_t upload_dms_tables(const DmsTables &t, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_dms, &t, sizeof(DmsTables), 0, cudaMemcpyHostToDevice, s);
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
// 0.105 at 461 654 x 80), this kernel does not (0.090 at 235 160 x 60 and at 461 654 x 80) but needs
// several waves of columns to fill the chip: launch_dms_columns picks it for large blocks of columns and
// for level counts whose tile does not fit shared memory.  (Register prefetch of the next level and L2
// prefetch of the one after made it slower on the large mesh, 3.76 against 3.31-3.43 ms: whatever the tile
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

// blocks (= inventory partials) of the kernel launch_dms_columns picks for this shape and variant
int dms_inventory_parts(int nL, int nC, int variant) {
  if (!dms_use_tiles(nL) || dms_use_columns(nL, nC, variant)) return (nC + 255) / 256;
  return (nC + kDmsTileCols - 1) / kDmsTileCols;
}
}  // namespace bgc
