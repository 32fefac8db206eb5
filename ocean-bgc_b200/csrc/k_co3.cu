// k_co3.cu — carbonate-chemistry kernels for sm_100a.
//
//   co3_cells_kernel       the two comp_CO3terms calls + comp_co3_sat_vals that
//                          BGC_SourceSink makes per cell (BGC_mod.F90:940-1001),
//                          one thread per CELL: the solve has no vertical coupling,
//                          so a 235k x 60 mesh offers 14M-way parallelism here
//                          instead of 235k-way.
//   co2calc_points_kernel  batched co2calc_1point (co2calc.F90:75-210)
//   surface_fluxes_kernel  BGC_SurfaceFluxes (BGC_mod.F90:2706-2957)
//
// Every warp runs the Newton/bisection iteration warp-synchronously (see
// bgc_co2.cuh), so all 32 lanes enter the solver even when some have no work.
#include <cstdlib>
#include "bgc_kernels.cuh"
#include "bgc_co2.cuh"

namespace bgc {

__constant__ BgcTables c_co3;

cudaError_t upload_bgc_tables_co3(const BgcTables &t, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_co3, &t, sizeof(BgcTables), 0, cudaMemcpyHostToDevice, s);
}

namespace {

// BGC_mod.F90:144-149
constexpr double phlo_surf_init = 7.0, phhi_surf_init = 9.0;
constexpr double phlo_3d_init = 6.0, phhi_3d_init = 9.0;
constexpr double del_ph = 0.20;
constexpr double xkw_coeff = 8.6e-9;   // BGC_parms.F90:488-489

__device__ __forceinline__ void report(unsigned long long *status, unsigned st) {
  if (st && status) {
    if (st & kSolveNoBracket) atomicAdd(&status[0], 1ull);
    if (st & kSolveNoConvergence) atomicAdd(&status[1], 1ull);
  }
}

// The nine inputs of one cell, as stored (no clamp, no unit change).
struct Co3CellIn { double temp, salt, zmid, dic, alk, po4, sio3, ph_prev, ph_prev_alt; };

// Address of input f of a cell (f in the member order of Co3CellIn).  Every cell of the (k,col)
// range is addressable whether it is active or not, so a cell can be requested before kmax is known.
__device__ __forceinline__ const double *co3_input_ptr(const Co3Args &A, int f, size_t cell, size_t nLnC) {
  const BgcIndices &I = c_co3.ind;
  switch (f) {
    case 0: return A.T + cell;
    case 1: return A.S + cell;
    case 2: return A.zmid + cell;
    case 3: return A.tracers + (cell + (size_t)(I.dic_ind - 1) * nLnC);
    case 4: return A.tracers + (cell + (size_t)(I.alk_ind - 1) * nLnC);
    case 5: return A.tracers + (cell + (size_t)(I.po4_ind - 1) * nLnC);
    case 6: return A.tracers + (cell + (size_t)(I.sio3_ind - 1) * nLnC);
    case 7: return A.ph_prev + cell;
    default: return A.ph_prev_alt + cell;
  }
}
__device__ __forceinline__ Co3CellIn co3_load(const Co3Args &A, size_t cell, size_t nLnC) {
  Co3CellIn in;
  in.temp = *co3_input_ptr(A, 0, cell, nLnC); in.salt = *co3_input_ptr(A, 1, cell, nLnC);
  in.zmid = *co3_input_ptr(A, 2, cell, nLnC); in.dic = *co3_input_ptr(A, 3, cell, nLnC);
  in.alk = *co3_input_ptr(A, 4, cell, nLnC);  in.po4 = *co3_input_ptr(A, 5, cell, nLnC);
  in.sio3 = *co3_input_ptr(A, 6, cell, nLnC); in.ph_prev = *co3_input_ptr(A, 7, cell, nLnC);
  in.ph_prev_alt = *co3_input_ptr(A, 8, cell, nLnC);
  return in;
}

// (level, column) of a cell and whether it is an active ocean cell (reads kmax)
struct Co3Where { int k; bool active; };
__device__ __forceinline__ Co3Where co3_where(const Co3Args &A, size_t cell, bool in_range) {
  // (32-bit division: nL * nC < 2^32 / 30, bgc_capi.cu check_dims; the 64-bit one costs ~30 instructions)
  const int k = in_range ? (int)((unsigned)cell / (unsigned)A.nC) : 0;
  const int col = in_range ? (int)(cell - (size_t)k * (size_t)A.nC) : 0;
  Co3Where w;
  w.k = k;
  w.active = in_range && col < A.nColumns && k < A.kmax[col];
  return w;
}

// One cell of the carbonate kernel from its (raw) inputs.  MUST be called by whole warps: lanes
// without work (`in_range` / `active` false) run the solver on benign values.
__device__ __forceinline__ void co3_cell_compute(const Co3Args &A, size_t cell, bool in_range, const Co3Where wh,
                                                 const Co3CellIn &in, const ExpTable &ex) {
  const bool active = wh.active;
  // benign mid-ocean inputs for lanes without work keep the solver well-posed
  double temp = 10.0, salt = 35.0, depth = 100.0, dic = 2000.0, alk = 2300.0, po4 = 1.0, sio3 = 10.0;
  double ph_prev = 8.0, ph_prev_alt = 8.0;
  if (active) {
    temp = in.temp;
    salt = in.salt;
    depth = in.zmid * 0.01;   // cm -> m (BGC_mod.F90:950)
    dic = gmax(0.0, in.dic);
    alk = gmax(0.0, in.alk);
    po4 = gmax(0.0, in.po4);
    sio3 = gmax(0.0, in.sio3);
    ph_prev = in.ph_prev;
    ph_prev_alt = in.ph_prev_alt;
  }
  const bool deep = wh.k > 0;   // the reference's (k > 1), 1-based

  // Both comp_CO3terms calls of a cell receive the SAME DIC/ALK/PO4/SiO3/T/S
  // (BGC_mod.F90:953 vs :975 — the alternative-CO2 call passes DIC_loc, not
  // DIC_ALT_CO2_loc), so the equilibrium constants are computed once.
  Co3Consts K;
  co3_coeffs<false>(deep, depth, temp, salt, K, ex);
  const Co3Totals tot = co3_totals(dic, alk, po4, sio3);

  double lo, hi;
  if (ph_prev != 0.0) { lo = ph_prev - del_ph; hi = ph_prev + del_ph; }
  else                { lo = phlo_3d_init;     hi = phhi_3d_init; }
  unsigned st = 0;
  const double h = solve_htotal(K, tot, lo, hi, st);

  double lo2, hi2;
  if (ph_prev_alt != 0.0) { lo2 = ph_prev_alt - del_ph; hi2 = ph_prev_alt + del_ph; }
  else                    { lo2 = phlo_3d_init;         hi2 = phhi_3d_init; }
  // Identical bracket => identical arithmetic => identical root: reuse it.  The
  // second solve runs only in warps where some lane's brackets really differ.
  const bool same_bracket = (lo2 == lo) && (hi2 == hi);
  double h_alt = h;
  if (!__all_sync(FULL_MASK, same_bracket)) {
    unsigned st2 = 0;
    const double h2 = solve_htotal(K, tot, lo2, hi2, st2);
    if (!same_bracket) { h_alt = h2; st |= st2; }
  }

  double sat_c, sat_a;
  co3_sat_vals(deep, depth, temp, salt, sat_c, sat_a, ex);

  if (!in_range) return;
  if (active) {
    // speciation, co2calc.F90:301-314
    const double k1 = K.k1, k2 = K.k2;
    {
      const double h2 = h * h;
      const double denom = frcp(h2 + k1 * h + k1 * k2);
      const double ph = ph_of_h(h);
      if (A.h2co3) A.h2co3[cell] = (tot.dic * h2 * denom) * kMassToVol;
      if (A.hco3) A.hco3[cell] = (tot.dic * k1 * h * denom) * kMassToVol;
      if (A.co3) A.co3[cell] = (tot.dic * k1 * k2 * denom) * kMassToVol;
      if (A.ph) A.ph[cell] = ph;
      A.ph_prev[cell] = ph;
    }
    {
      const double h2 = h_alt * h_alt;
      const double denom = frcp(h2 + k1 * h_alt + k1 * k2);
      const double ph = ph_of_h(h_alt);
      if (A.h2co3_alt) A.h2co3_alt[cell] = (tot.dic * h2 * denom) * kMassToVol;
      if (A.hco3_alt) A.hco3_alt[cell] = (tot.dic * k1 * h_alt * denom) * kMassToVol;
      if (A.co3_alt) A.co3_alt[cell] = (tot.dic * k1 * k2 * denom) * kMassToVol;
      if (A.ph_alt) A.ph_alt[cell] = ph;
      A.ph_prev_alt[cell] = ph;
    }
    if (A.sat_calc) A.sat_calc[cell] = sat_c;
    if (A.sat_arag) A.sat_arag[cell] = sat_a;
    report(A.status, st);
  } else {
    // the reference zero-fills these diagnostics for land / below-bottom cells
    // (BGC_mod.F90:649-658); PH_PREV_* are left untouched there.
    if (A.h2co3) A.h2co3[cell] = 0.0;
    if (A.hco3) A.hco3[cell] = 0.0;
    if (A.co3) A.co3[cell] = 0.0;
    if (A.ph) A.ph[cell] = 0.0;
    if (A.h2co3_alt) A.h2co3_alt[cell] = 0.0;
    if (A.hco3_alt) A.hco3_alt[cell] = 0.0;
    if (A.co3_alt) A.co3_alt[cell] = 0.0;
    if (A.ph_alt) A.ph_alt[cell] = 0.0;
    if (A.sat_calc) A.sat_calc[cell] = 0.0;
    if (A.sat_arag) A.sat_arag[cell] = 0.0;
  }
}

// One cell, inputs requested here.  `cell_end` bounds the launch's share of the mesh: a launch
// covers the cells [A.cell_begin, cell_end), both multiples of 32 unless they are the ends of the mesh,
// so that a warp never straddles two launches.  The nine inputs do not wait for kmax.
__device__ __forceinline__ void co3_cell(const Co3Args &A, size_t cell, size_t cell_end, const ExpTable &ex) {
  const bool in_range = cell < cell_end;
  const size_t nLnC = (size_t)A.nL * (size_t)A.nC;
  Co3CellIn in = {};
  if (in_range) in = co3_load(A, cell, nLnC);
  co3_cell_compute(A, cell, in_range, co3_where(A, cell, in_range), in, ex);
}

// PERSISTENT = false: one thread per cell of [cell_begin, cell_end), 256-thread blocks, whole warps only
// (the grid is sized so that every launched warp is complete).
// PERSISTENT = true: a FEW 512-thread blocks walk the range with a grid stride.  512 threads x 128
// registers are a whole SM's register file, so each block owns one SM: bgc_capi.cu launches as many
// as the column sweep leaves idle when it is less than one wave, and the carbonate work runs on
// exactly those SMs while the sweep occupies the others (see source_sink_device).
template <int BLOCK, bool PERSISTENT>
__global__ void __launch_bounds__(BLOCK, PERSISTENT ? 1 : 2)
co3_cells_kernel(const __grid_constant__ Co3Args A) {
  __shared__ double s_exp2[64];
  exp_table_load(s_exp2);
  const ExpTable ex{s_exp2};
  const size_t ncell = (size_t)A.nL * (size_t)A.nC;
  const size_t end = A.cell_end ? (size_t)A.cell_end : ncell;
  const unsigned tslot = (PERSISTENT && A.block_trace && threadIdx.x == 0) ? block_trace_begin(A.block_trace, 2u) : ~0u;
  if (PERSISTENT) {
    // a warp takes a trip or skips it as a whole (the solver's votes are warp-wide); no block
    // barrier follows exp_table_load, so the warps of a block need not make the same number of trips
    for (size_t base = (size_t)A.cell_begin + (size_t)blockIdx.x * BLOCK; base < end; base += (size_t)gridDim.x * BLOCK)
      if (base + (threadIdx.x & ~31u) < end) co3_cell(A, base + threadIdx.x, end, ex);
  } else {
    co3_cell(A, (size_t)A.cell_begin + (size_t)blockIdx.x * BLOCK + threadIdx.x, end, ex);
  }
  if (PERSISTENT && A.block_trace) { __syncthreads(); if (threadIdx.x == 0) block_trace_end(A.block_trace, tslot); }
}

// Experiment only (BGC_CO3_DUMMY=<iterations>): SM-filling blocks that run a tiny FP64 loop instead of the
// carbonate code, to tell what a co-resident kernel costs the column sweep apart from its code size.
__global__ void __launch_bounds__(512, 1) fp64_spin_kernel(double *sink, int iters, unsigned long long *trace) {
  const unsigned tslot = (trace && threadIdx.x == 0) ? block_trace_begin(trace, 2u) : ~0u;
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.5, c = 0.25, d = 0.125;
  for (int i = 0; i < iters; ++i) {
    a = fma(a, 1.0000001, 1e-9); b = fma(b, 0.9999999, 1e-9); c = fma(c, 1.0000002, 1e-9); d = fma(d, 0.9999998, 1e-9);
  }
  if (a + b + c + d == 1.2345e300) *sink = a;
  if (trace) { __syncthreads(); if (threadIdx.x == 0) block_trace_end(trace, tslot); }
}

// Saturation-depth scan (BGC_mod.F90:1003-1032): the depth at which CO3 first falls to the
// calcite / aragonite saturation concentration, linearly interpolated between cell centres;
// -1 while still supersaturated, 0 for a column undersaturated at the surface, the bottom
// depth if the whole column is supersaturated.  Per-column diagnostics (zero-filled for
// inactive columns, :625-727).
__global__ void __launch_bounds__(128)
zsat_columns_kernel(const __grid_constant__ ZsatArgs A) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= A.nC) return;
  int kmax = (col < A.nColumns) ? A.kmax[col] : 0;
  if (kmax > A.nL) kmax = A.nL;
  double ZSATCALC = 0.0, ZSATARAG = 0.0, CALC_ANOM_km1 = 0.0, ARAG_ANOM_km1 = 0.0, zmid_km1 = 0.0;
#pragma unroll 4
  for (int k = 0; k < kmax; ++k) {
    const unsigned i2 = (unsigned)col + (unsigned)A.nC * (unsigned)k;
    const double CO3 = A.co3[i2], sat_c = A.sat_calc[i2], sat_a = A.sat_arag[i2], zmid = A.zmid[i2];
    if (k == 0) {
      ZSATCALC = (CO3 > sat_c) ? -1.0 : 0.0;
      ZSATARAG = (CO3 > sat_a) ? -1.0 : 0.0;
    } else {
      const double w4 = zmid_km1 + (zmid - zmid_km1);   // as written in the reference (:1009)
      if (ZSATCALC == -1.0 && CO3 <= sat_c)
        ZSATCALC = fdiv(w4 * CALC_ANOM_km1, (CALC_ANOM_km1 - (CO3 - sat_c)));
      if (ZSATARAG == -1.0 && CO3 <= sat_a)
        ZSATARAG = fdiv(w4 * ARAG_ANOM_km1, (ARAG_ANOM_km1 - (CO3 - sat_a)));
      if (k == kmax - 1 && (ZSATCALC == -1.0 || ZSATARAG == -1.0)) {
        const double zbot = A.zbot[i2];
        if (ZSATCALC == -1.0) ZSATCALC = zbot;
        if (ZSATARAG == -1.0) ZSATARAG = zbot;
      }
    }
    CALC_ANOM_km1 = CO3 - sat_c;
    ARAG_ANOM_km1 = CO3 - sat_a;
    zmid_km1 = zmid;
  }
  if (A.zsatcalc) A.zsatcalc[col] = ZSATCALC;
  if (A.zsatarag) A.zsatarag[col] = ZSATARAG;
}

// co2calc_1point (co2calc.F90:75-210): always level 1 => no pressure correction.
// The reference converts depth -> press_bar and then passes press_bar as the
// `depth` of comp_co3_coeffs (:156-160); with k = 1 neither value is used.
struct SurfaceCo2 { double ph, co2star, dco2star, pco2surf, dpco2; };

__device__ __forceinline__ SurfaceCo2 co2calc_1point(double temp, double salt, double dic_in, double ta_in,
                                                     double pt_in, double sit_in, double phlo, double phhi,
                                                     double xco2_in, double atmpres, unsigned &st) {
  Co3Consts K;
  co3_coeffs<true>(false, 0.0, temp, salt, K, ExpPoly());
  const Co3Totals tot = co3_totals(dic_in, ta_in, pt_in, sit_in);
  const double htotal = solve_htotal(K, tot, phlo, phhi, st);

  const double xco2 = xco2_in * 1e-6;
  const double htotal2 = htotal * htotal;
  SurfaceCo2 r;
  double co2star = fdiv(tot.dic * htotal2, (htotal2 + K.k1 * htotal + K.k1 * K.k2));
  const double co2starair = xco2 * K.ff * atmpres;
  double dco2star = co2starair - co2star;
  r.ph = ph_of_h(htotal);
  double pco2surf = fdiv(co2star, K.ff);
  double dpco2 = pco2surf - xco2 * atmpres;
  r.co2star = co2star * kMassToVol;
  r.dco2star = dco2star * kMassToVol;
  r.pco2surf = pco2surf * 1e6;
  r.dpco2 = dpco2 * 1e6;
  return r;
}

__global__ void __launch_bounds__(256)
co2calc_points_kernel(const __grid_constant__ Co2PointsArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < A.n;
  double temp = 10.0, salt = 35.0, dic = 2000.0, ta = 2300.0, pt = 1.0, sit = 10.0, lo = 7.0, hi = 9.0,
         xco2 = 400.0, atmpres = 1.0;
  if (active) {
    temp = A.temp[i]; salt = A.salt[i]; dic = A.dic[i]; ta = A.ta[i]; pt = A.pt[i]; sit = A.sit[i];
    lo = A.phlo[i]; hi = A.phhi[i]; xco2 = A.xco2[i]; atmpres = A.atmpres[i];
  }
  unsigned st = 0;
  const SurfaceCo2 r = co2calc_1point(temp, salt, dic, ta, pt, sit, lo, hi, xco2, atmpres, st);
  if (active) {
    A.ph[i] = r.ph;
    A.co2star[i] = r.co2star;
    A.dco2star[i] = r.dco2star;
    A.pco2surf[i] = r.pco2surf;
    A.dpco2[i] = r.dpco2;
    report(A.status, st);
  }
}

// Batched comp_CO3terms (co2calc.F90:214-316) and comp_co3_sat_vals (:1096-1238): the two other
// public procedures of the reference's co2calc module, one point per thread.  k is the
// reference's 1-based level index: the pressure correction is keyed on (k > 1), not on depth.
__global__ void __launch_bounds__(256)
co3terms_points_kernel(const __grid_constant__ Co3TermsPointsArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < A.n;
  double temp = 10.0, salt = 35.0, depth = 100.0, dic = 2000.0, ta = 2300.0, pt = 1.0, sit = 10.0, lo = 7.0, hi = 9.0;
  int k = 1;
  if (active) {
    k = A.k ? A.k[i] : A.k_all;
    depth = A.depth[i]; temp = A.temp[i]; salt = A.salt[i]; dic = A.dic[i]; ta = A.ta[i]; pt = A.pt[i];
    sit = A.sit[i]; lo = A.phlo[i]; hi = A.phhi[i];
  }
  __shared__ double s_exp2[64];
  exp_table_load(s_exp2);
  const ExpTable ex{s_exp2};
  Co3Consts K;
  co3_coeffs<false>(k > 1, depth, temp, salt, K, ex);
  const Co3Totals tot = co3_totals(dic, ta, pt, sit);
  unsigned st = 0;
  const double h = solve_htotal(K, tot, lo, hi, st);
  if (active) {
    const double h2 = h * h;
    const double denom = frcp(h2 + K.k1 * h + K.k1 * K.k2);
    A.ph[i] = ph_of_h(h);
    A.h2co3[i] = (tot.dic * h2 * denom) * kMassToVol;
    A.hco3[i] = (tot.dic * K.k1 * h * denom) * kMassToVol;
    A.co3[i] = (tot.dic * K.k1 * K.k2 * denom) * kMassToVol;
    report(A.status, st);
  }
}

__global__ void __launch_bounds__(256)
co3_sat_points_kernel(const __grid_constant__ Co3SatPointsArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double s_exp2[64];
  exp_table_load(s_exp2);
  const ExpTable ex{s_exp2};
  if (i >= A.n) return;
  const int k = A.k ? A.k[i] : A.k_all;
  double sat_c, sat_a;
  co3_sat_vals(k > 1, A.depth[i], A.temp[i], A.salt[i], sat_c, sat_a, ex);
  A.sat_calc[i] = sat_c;
  A.sat_arag[i] = sat_a;
}

__device__ __forceinline__ double schmidt_o2(double SST) {   // Keeling et al. 1998 (BGC_mod.F90:2965-3005)
  return 1638.0 + SST * (-81.83 + SST * (1.483 + SST * (-0.008004)));
}
__device__ __forceinline__ double schmidt_co2(double SST) {  // Wanninkhof 1992 (:3091-3128)
  return 2073.1 + SST * (-125.62 + SST * (3.6276 + SST * (-0.043219)));
}
__device__ __forceinline__ double o2sat(double SST, double SSS, double T0K) {   // Garcia & Gordon 1992 (:3012-3083)
  const double TS = log(((T0K + 25.0) - SST) / (T0K + SST));
  const double r = bexp(2.00907 + TS * (3.22014 + TS * (4.05010 + TS * (4.94457 + TS * (-2.56847E-1 + TS * 3.88767)))) +
                       SSS * ((-6.24523E-3 + TS * (-7.37614E-3 + TS * (-1.03410E-2 + TS * -8.17083E-3))) +
                              SSS * -4.88682E-7));
  return r / 0.0223916;
}

__global__ void __launch_bounds__(128)
surface_fluxes_kernel(const __grid_constant__ SurfArgs A) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nC = (size_t)A.nC;
  const size_t nLnC = (size_t)A.nL * nC;
  const bool in_range = col < A.nC;
  const bool active = col < A.nColumns;
  const BgcIndices &I = c_co3.ind;
  const BgcParams &P = c_co3.p;
#define SURF(ind_) gmax(0.0, A.tracers[(size_t)col + (size_t)((ind_) - 1) * nLnC])
#define FLX(arr, ind_) A.f.arr[(size_t)col + (size_t)((ind_) - 1) * nC]
#define DG(name, val) do { if (A.d.name) A.d.name[col] = (val); } while (0)

  double DIC_loc = 2000.0, DIC_ALT_loc = 2000.0, ALK_loc = 2300.0, PO4_loc = 1.0, SiO3_loc = 10.0, O2_loc = 0.0;
  double sst = 10.0, sss = 35.0, pres = 1.0, xkw_ice = 0.0, ph_s = 0.0, ph_s_alt = 0.0, atm = 400.0, atm_alt = 400.0;
  if (active) {
    DIC_loc = SURF(I.dic_ind);
    DIC_ALT_loc = SURF(I.dic_alt_co2_ind);
    ALK_loc = SURF(I.alk_ind);
    PO4_loc = SURF(I.po4_ind);
    SiO3_loc = SURF(I.sio3_ind);
    O2_loc = SURF(I.o2_ind);

    // in-place side effects on the forcing (BGC_mod.F90:2828-2838)
    FLX(depositionFlux, I.fe_ind) = FLX(depositionFlux, I.fe_ind) * P.parm_Fe_bioavail;
    FLX(riverFlux, I.fe_ind) = FLX(riverFlux, I.fe_ind) * P.parm_Fe_bioavail;
    FLX(gasFlux, I.fe_ind) = FLX(gasFlux, I.fe_ind) * P.parm_Fe_bioavail;
    FLX(seaIceFlux, I.fe_ind) = FLX(seaIceFlux, I.fe_ind) * P.parm_Fe_bioavail;
    double ice = A.f.iceFraction[col];
    if (ice < 0.0) ice = 0.0;
    if (ice > 1.0) ice = 1.0;
    A.f.iceFraction[col] = ice;

    const double xkw = xkw_coeff * A.f.windSpeedSquared10m[col];
    xkw_ice = (1.0 - ice) * xkw;
    sst = A.f.SST[col];
    sss = A.f.SSS[col];
    pres = A.f.surfacePressure[col];
    if (A.f.lcalc_CO2_gas_flux) {
      ph_s = A.f.surface_pH[col];
      ph_s_alt = A.f.surface_pH_alt_co2[col];
      atm = A.f.atmCO2[col];
      atm_alt = A.f.atmCO2_ALT_CO2[col];
    }
  }

  if (in_range && !active) {   // whole-array zero fill of the flux diagnostics (:2789-2802)
    DG(pistonVel_O2, 0.0); DG(SCHMIDT_O2, 0.0); DG(O2SAT, 0.0); DG(xkw, 0.0);
  } else if (active) {
    if (A.f.lcalc_O2_gas_flux) {
      const double sc = schmidt_o2(sst);
      const double sat1 = o2sat(sst, sss, P.T0_Kelvin_BGC);
      const double pv = xkw_ice * sqrt(660.0 / sc);
      const double sat = pres * sat1;
      FLX(gasFlux, I.o2_ind) = pv * (sat - O2_loc);
      DG(pistonVel_O2, pv); DG(SCHMIDT_O2, sc); DG(O2SAT, sat); DG(xkw, xkw_ice);
    } else {
      DG(pistonVel_O2, 0.0); DG(SCHMIDT_O2, 0.0); DG(O2SAT, 0.0); DG(xkw, 0.0);
    }
  }

  if (A.f.lcalc_CO2_gas_flux) {   // uniform across the grid: every lane takes the solver path
    const double sc = schmidt_co2(sst);
    const double pv = xkw_ice * sqrt(660.0 / sc);
    double lo, hi;
    if (ph_s != 0.0) { lo = ph_s - del_ph; hi = ph_s + del_ph; }
    else             { lo = phlo_surf_init; hi = phhi_surf_init; }
    unsigned st = 0;
    const SurfaceCo2 r = co2calc_1point(sst, sss, DIC_loc, ALK_loc, PO4_loc, SiO3_loc, lo, hi, atm, pres, st);
    if (ph_s_alt != 0.0) { lo = ph_s_alt - del_ph; hi = ph_s_alt + del_ph; }
    else                 { lo = phlo_surf_init;    hi = phhi_surf_init; }
    const SurfaceCo2 q = co2calc_1point(sst, sss, DIC_ALT_loc, ALK_loc, PO4_loc, SiO3_loc, lo, hi, atm_alt, pres, st);
    if (active) {
      A.f.surface_pH[col] = r.ph;
      FLX(gasFlux, I.dic_ind) = pv * r.dco2star;
      DG(co2star, r.co2star); DG(dco2star, r.dco2star); DG(pco2surf, r.pco2surf); DG(dpco2, r.dpco2);
      DG(pistonVel_CO2, pv); DG(SCHMIDT_CO2, sc);
      A.f.surface_pH_alt_co2[col] = q.ph;
      FLX(gasFlux, I.dic_alt_co2_ind) = pv * q.dco2star;
      DG(co2star_alt_co2, q.co2star); DG(dco2star_alt_co2, q.dco2star);
      DG(pco2surf_alt_co2, q.pco2surf); DG(dpco2_alt_co2, q.dpco2);
      report(A.status, st);
    }
  }
  if (in_range && (!active || !A.f.lcalc_CO2_gas_flux)) {
    DG(co2star, 0.0); DG(dco2star, 0.0); DG(pco2surf, 0.0); DG(dpco2, 0.0);
    DG(pistonVel_CO2, 0.0); DG(SCHMIDT_CO2, 0.0);
    DG(co2star_alt_co2, 0.0); DG(dco2star_alt_co2, 0.0); DG(pco2surf_alt_co2, 0.0); DG(dpco2_alt_co2, 0.0);
  }

  if (active) {   // net flux and the ALK correction (:2929-2942)
#pragma unroll 6
    for (int n = 1; n <= BGC_TRACER_CNT; ++n)
      FLX(netFlux, n) = FLX(depositionFlux, n) + FLX(gasFlux, n) + FLX(riverFlux, n) + FLX(seaIceFlux, n);
    FLX(netFlux, I.alk_ind) = FLX(netFlux, I.alk_ind) + FLX(netFlux, I.nh4_ind) - FLX(netFlux, I.no3_ind);
  }
#undef SURF
#undef FLX
#undef DG
}

}  // namespace

static inline int ceil_div_sz(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// persistent_blocks = 0: one thread per cell of the range; > 0: that many SM-filling blocks (see the kernel)
cudaError_t launch_co3_cells(const Co3Args &a, int persistent_blocks, cudaStream_t s) {
  const size_t ncell = (size_t)a.nL * (size_t)a.nC;
  const size_t end = a.cell_end ? (size_t)a.cell_end : ncell;
  if (ncell == 0 || end <= (size_t)a.cell_begin) return cudaSuccess;
  if (end > ncell || ((a.cell_begin & 31) != 0)) return cudaErrorInvalidValue;
  static int dummy = -1;
  if (dummy < 0) { const char *v = getenv("BGC_CO3_DUMMY"); dummy = v ? atoi(v) : 0; }
  if (persistent_blocks > 0 && dummy > 0) {
    fp64_spin_kernel<<<persistent_blocks, 512, 0, s>>>(a.ph_prev + 0 * ncell, dummy, a.block_trace);
    return cudaGetLastError();
  }
  if (persistent_blocks > 0) co3_cells_kernel<512, true><<<persistent_blocks, 512, 0, s>>>(a);
  else co3_cells_kernel<256, false><<<ceil_div_sz(end - (size_t)a.cell_begin, 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_zsat_columns(const ZsatArgs &a, cudaStream_t s) {
  if (a.nC <= 0) return cudaSuccess;
  zsat_columns_kernel<<<ceil_div_sz((size_t)a.nC, 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_co2calc_points(const Co2PointsArgs &a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  co2calc_points_kernel<<<ceil_div_sz((size_t)a.n, 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_co3terms_points(const Co3TermsPointsArgs &a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  co3terms_points_kernel<<<(a.n + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_co3_sat_points(const Co3SatPointsArgs &a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  co3_sat_points_kernel<<<(a.n + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_surface_fluxes(const SurfArgs &a, cudaStream_t s) {
  if (a.nC <= 0) return cudaSuccess;
  surface_fluxes_kernel<<<ceil_div_sz((size_t)a.nC, 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace bgc
