// k_eco.cu — the ecosystem + sinking-particle column sweep for sm_100a.
//
// Replaces the column_loop of BGC_SourceSink (BGC_mod.F90:799-1970) together
// with init_particulate_terms (:2006-2109) and compute_particulate_terms
// (:2116-2699).  One thread owns one ocean column and walks it top to bottom.
// The setup_loop clamp pass (:733-789) and the whole-array zero fills (:570,
// :625-727) are folded into the same sweep: a cell is read once and every output
// element is written exactly once.
//
// What bounds it.  ~3.6-4.5 k instructions per cell, of which a third are FP64 with
// an 8-cycle dependent latency (scripts/micro/fp64_lat.cu), at 255 registers per
// thread = 8 warps per SM: the kernel is bound by exposed dependency latency, not by
// HBM or by the FP64 pipe (profiles/README.md).  Registers are the scarce resource
// (a thread's state is ~1.5 KB across registers and shared memory), so
//   * only the particle fluxes stay in registers across levels; the column state that
//     is touched once per level (PAR, dust fluxes, the QA deficit, the previous zbot)
//     and the fourteen column integrals live in per-thread shared-memory rows;
//   * the loop over the four functional groups is unrolled, so the group's table
//     fields are constant-bank operands; every SUM(x(:)) of the reference is a
//     running accumulator in the same left-to-right order as Fortran's SUM;
//   * exp / log / reciprocal come from bgc_math.cuh (constant-bank coefficients,
//     Estrin evaluation, one Newton step on the MUFU seed, no slow-path branches);
//   * a functional group whose biomass is exactly zero in a whole warp skips its body
//     (zero-biomass shortcut, see the group loop);
//   * a block-wide barrier per level (needed by the stage hand-over below) also keeps
//     the warps of a block on the same stretch of code, which the instruction cache likes.
//
// Input staging.  With ~8 resident warps per SM (255 registers per thread) nothing
// hides an HBM round trip, and a level has ~8 dependent batches of loads.  So one
// elected thread per block fetches the NEXT level's slab of every input array -
// 27 tracers, T, S, zmid, dz, zbot, FESEDFLUX,
// BLOCK consecutive columns = one contiguous BLOCK*8-byte run each - with
// cp.async.bulk (the TMA unit's 1-D bulk copy) into a double-buffered shared-memory
// stage, completion signalled on an mbarrier; the compute threads only ever read
// shared memory.  With an odd nColumnsMax every other level of a block starts 8 bytes
// off a 16-byte boundary: the bulk copy then fetches the ALIGNED SUPERSET of the run
// (one element earlier, rounded up to 16 bytes; the stage rows are two doubles wider
// for it) and the level's stage pointer is shifted by that one element, so the odd
// width costs nothing.  Only when a caller's base pointer itself is not 16-byte aligned
// (or nColumnsMax and the level count are both odd: the tracer slabs then alternate
// between the two alignments) is the stage filled by 8-byte cp.async copies (LDGSTS),
// one column per thread, two levels ahead as well.  The staging differs, the arithmetic is the SAME instantiation,
// so results do not depend on the width of the block a host happens to pick.
//
// The carbonate solve of each cell has no vertical coupling and runs in the
// cell-parallel kernel of k_co3.cu, and the saturation-depth scan (:1003-1032) that
// consumes its results is a separate small column kernel there: this kernel does not
// depend on the carbonate kernel at all, so the two run concurrently (bgc_capi.cu) and
// the FP64-bound carbonate work fills the SMs that the last, partial wave of this kernel
// leaves idle (235 160 columns = 6.2 waves of 148 blocks x 256 columns).
#include <cstdlib>
#include "bgc_kernels.cuh"
#include "bgc_math.cuh"
#include "bgc_reduce.cuh"

namespace bgc {

// Quantities derived once per bgc_set_params from the parameter tables.
struct EcoDerived {
  double r_kNO3[4], r_kNH4[4], r_kPO4[4], r_kDOP[4];
  double cks_kFe[4], r_cks_kFe[4], cksi_kFe[4], cksi_kSiO3[4], r_cksi_kSiO3[4];
  double agg_max_dps[4], agg_min_dps[4];
  double r_dTN[4], r_dTS[4];
  double r_o2_min_delta;
  double r_scalelen_dz[4];
};

__constant__ BgcTables c_eco;
__constant__ EcoDerived c_der;

cudaError_t upload_bgc_tables_eco(const BgcTables &t, cudaStream_t s) {
  constexpr double dps = 1.0 / 86400.0;
  EcoDerived d;
  for (int a = 0; a < BGC_AUTOTROPH_CNT; ++a) {
    const BgcAutotroph &at = t.a[a];
    d.r_kNO3[a] = 1.0 / at.kNO3;
    d.r_kNH4[a] = 1.0 / at.kNH4;
    d.r_kPO4[a] = 1.0 / at.kPO4;
    d.r_kDOP[a] = 1.0 / at.kDOP;
    d.cks_kFe[a] = t.p.cks * at.kFe;
    d.r_cks_kFe[a] = 1.0 / d.cks_kFe[a];
    d.cksi_kFe[a] = t.p.cksi * at.kFe;
    d.cksi_kSiO3[a] = t.p.cksi * at.kSiO3;
    d.r_cksi_kSiO3[a] = 1.0 / d.cksi_kSiO3[a];
    d.agg_max_dps[a] = at.agg_rate_max * dps;
    d.agg_min_dps[a] = at.agg_rate_min * dps;
    d.r_dTN[a] = 1.0 / (at.temp_thresN - at.temp_optN);
    d.r_dTS[a] = 1.0 / (at.temp_thresS - at.temp_optS);
  }
  d.r_o2_min_delta = 1.0 / t.p.parm_o2_min_delta;
  d.r_scalelen_dz[0] = 0.0;
  for (int n = 1; n < 4; ++n) d.r_scalelen_dz[n] = 1.0 / (t.p.parm_scalelen_z[n] - t.p.parm_scalelen_z[n - 1]);
  cudaError_t e = cudaMemcpyToSymbolAsync(c_eco, &t, sizeof(BgcTables), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  // `d` is pageable stack memory: the runtime stages it before returning.
  return cudaMemcpyToSymbolAsync(c_der, &d, sizeof(EcoDerived), 0, cudaMemcpyHostToDevice, s);
}

namespace {

// BGC_parms.F90:37-40
constexpr double spd = 86400.0;
constexpr double dps = 1.0 / spd;
constexpr double yps = 1.0 / (365.0 * spd);
// BGC_parms.F90:327-339
constexpr double parm_Red_D_C_P = 117.0;
constexpr double parm_Red_D_C_O2 = parm_Red_D_C_P / 170.0;
constexpr double parm_Remin_D_C_O2 = parm_Red_D_C_P / 138.0;
constexpr double parm_Red_Fe_C = 3.0e-6;
constexpr double parm_Red_D_C_O2_diaz = parm_Red_D_C_P / 150.0;
// :371-386
constexpr double fe_scavenge_thres1 = 0.8e-3;
constexpr double fe_max_scale2 = 1200.0;
constexpr double dust_to_Fe = 0.035 / 55.847 * 1.0e9;
// :394-429
constexpr double caco3_poc_min = 0.4;
constexpr double spc_poc_fac = 0.11;
constexpr double f_graze_sp_poc_lim = 0.3;
constexpr double f_photosp_CaCO3 = 0.4;
constexpr double f_graze_CaCO3_remin = 0.33;
constexpr double f_graze_si_remin = 0.35;
constexpr double r_Nfix_photo = 1.25;
constexpr double Qn = 0.137;            // "Q", N/C
constexpr double Qp_zoo_pom = 0.00855;
constexpr double Qfe_zoo = 3.0e-6;
constexpr double gQsi_0 = 0.137;
constexpr double gQsi_max = 0.685;
constexpr double gQsi_min = 0.0457;
constexpr double QCaCO3_max = 0.4;
constexpr double denitrif_C_N = parm_Red_D_C_P / 136.0;
// :435-477
constexpr double thres_z1 = 100.0e2;
constexpr double thres_z2 = 150.0e2;
constexpr double loss_thres_zoo = 0.005;
constexpr double CaCO3_temp_thres1 = 6.0;
constexpr double CaCO3_temp_thres2 = -2.0;
constexpr double CaCO3_sp_thres = 4.0;
constexpr double f_qsw_par = 0.45;
constexpr double Tref = 30.0;
constexpr double Q_10 = 1.5;
constexpr double kLnQ10 = 0.4054651081081644;      // log(1.5), correctly rounded
constexpr double kLn099 = -0.01005033585350145;    // log(0.99)
constexpr double DOC_reminR = (1.0 / 250.0) * dps;
constexpr double DON_reminR = (1.0 / 160.0) * dps;
constexpr double DOFe_reminR = (1.0 / 160.0) * dps;
constexpr double DOP_reminR = (1.0 / 160.0) * dps;
constexpr double DONr_reminR = (1.0 / (365.0 * 2.5)) * dps;
constexpr double DOPr_reminR = (1.0 / (365.0 * 2.5)) * dps;
constexpr double DONrefract = 0.08;
constexpr double DOPrefract = 0.03;
constexpr double mpercm = 0.01;

// sinking_particle class constants, init_particulate_terms (BGC_mod.F90:2046-2069)
constexpr double POC_mass = 12.01;
constexpr double CaCO3_gamma = 0.30, CaCO3_mass = 100.09, CaCO3_rho = 0.05 * CaCO3_mass / POC_mass;
constexpr double SiO2_gamma = 0.030, SiO2_mass = 60.08, SiO2_rho = 0.05 * SiO2_mass / POC_mass;
constexpr double dust_diss0 = 20000.0, dust_gamma = 0.97, dust_mass = 1.0e9,
                 dust_rho = 0.05 * dust_mass / POC_mass;
constexpr double P_iron_gamma = 0.0;

constexpr int NA = BGC_AUTOTROPH_CNT;

// DIAG: 0 = no diagnostic array, 1 = any subset (NULL checks), 2 = every array present
#define HAS(name) (DIAG == 2 || A.d.name != nullptr)
#define ST2(name, val) do { if (HAS(name)) A.d.name[i2] = (val); } while (0)
#define STA(name, val) do { if (HAS(name)) A.d.name[ia] = (val); } while (0)
#define STA_AT(name, a_, val) do { if (HAS(name)) A.d.name[i2 + (unsigned)(a_) * nLnC] = (val); } while (0)
#define STC(name, val) do { if (HAS(name)) A.d.name[col] = (val); } while (0)
#define STCA(name, a_, val) do { if (HAS(name)) A.d.name[(unsigned)col + (unsigned)(a_) * (unsigned)nC] = (val); } while (0)

// The (k,col) diagnostics this kernel owns: BGC_DIAG_K2_LIST minus the ten carbonate
// arrays (written by co3_cells_kernel) and the three arrays the reference declares but
// never zeroes nor writes (diag_POC_ACCUM, diag_DONr_remin, diag_DOPr_remin).
#define ECO_DIAG_K2_LIST(X) \
  X(diag_tot_Nfix) X(diag_O2_PRODUCTION) X(diag_O2_CONSUMPTION) X(diag_AOU) \
  X(diag_PO4_RESTORE) X(diag_NO3_RESTORE) X(diag_SiO3_RESTORE) X(diag_PAR_avg) \
  X(diag_POC_FLUX_IN) X(diag_POC_PROD) X(diag_POC_REMIN) \
  X(diag_CaCO3_FLUX_IN) X(diag_CaCO3_PROD) X(diag_CaCO3_REMIN) X(diag_SiO2_FLUX_IN) \
  X(diag_SiO2_PROD) X(diag_SiO2_REMIN) X(diag_dust_FLUX_IN) X(diag_dust_REMIN) \
  X(diag_P_iron_FLUX_IN) X(diag_P_iron_PROD) X(diag_P_iron_REMIN) X(diag_auto_graze_TOT) \
  X(diag_zoo_loss) X(diag_photoC_TOT) X(diag_photoC_NO3_TOT) X(diag_DOC_prod) \
  X(diag_DOC_remin) X(diag_DON_prod) X(diag_DON_remin) X(diag_DOFe_prod) \
  X(diag_DOFe_remin) X(diag_DOP_prod) X(diag_DOP_remin) X(diag_Fe_scavenge) \
  X(diag_Fe_scavenge_rate) X(diag_NITRIF) X(diag_DENITRIF) \
  X(diag_calcToSed) X(diag_pocToSed) X(diag_ponToSed) \
  X(diag_popToSed) X(diag_bsiToSed) X(diag_dustToSed) X(diag_pfeToSed) \
  X(diag_SedDenitrif) X(diag_OtherRemin) X(diag_tot_CaCO3_form)
// The per-column diagnostics this kernel owns: BGC_DIAG_C1_LIST minus the two saturation
// depths (zsat_columns_kernel, k_co3.cu).
#define ECO_DIAG_C1_LIST(X) \
  X(diag_photoC_TOT_zint) X(diag_photoC_NO3_TOT_zint) X(diag_Jint_Ctot) \
  X(diag_Jint_100m_Ctot) X(diag_Jint_Ntot) X(diag_Jint_100m_Ntot) X(diag_Jint_Ptot) \
  X(diag_Jint_100m_Ptot) X(diag_Jint_Sitot) X(diag_Jint_100m_Sitot) \
  X(diag_Chl_TOT_zint_100m) X(diag_tot_CaCO3_form_zint) X(diag_tot_bSi_form) \
  X(diag_O2_ZMIN) X(diag_O2_ZMIN_DEPTH)
#define COUNT_ONE(name) +1
static_assert((0 ECO_DIAG_C1_LIST(COUNT_ONE)) + 2 == (0 BGC_DIAG_C1_LIST(COUNT_ONE)),
              "ECO_DIAG_C1_LIST is out of step with BGC_DIAG_C1_LIST");
static_assert((0 ECO_DIAG_K2_LIST(COUNT_ONE)) + 13 == (0 BGC_DIAG_K2_LIST(COUNT_ONE)),
              "ECO_DIAG_K2_LIST is out of step with BGC_DIAG_K2_LIST");

#define ZERO_K2(name) ST2(name, 0.0);
#define ZERO_KA(name) { STA_AT(name, 0, 0.0); STA_AT(name, 1, 0.0); STA_AT(name, 2, 0.0); STA_AT(name, 3, 0.0); }
#define ZERO_CA(name) { STCA(name, 0, 0.0); STCA(name, 1, 0.0); STCA(name, 2, 0.0); STCA(name, 3, 0.0); }
#define ZERO_C1(name) STC(name, 0.0);

// Shared memory of a block, in rows of BLOCK doubles (one slot per thread):
//   2 stages x R_ROWS   the level's input slab: rows 0..29 = tracer slots, then the rows below
//   X_ROWS              per-thread scratch: the column's depth (as an int; rows 1-3 spare); DIAG: the three per-group column
//                       integrals (:1838-1846, :1268), 4 each, and the fourteen column integrals
//                       that are touched once per level (Jint_*, the z-integrals, the O2 minimum):
//                       one shared-memory read-modify-write per level each instead of 28 registers
//                       held across the whole level body (they used to spill to local memory,
//                       which misses the few KB of L1 left beside a 221 KB carve-out)
//   2 mbarriers
enum { R_T = BGC_TRACER_CNT, R_ZMID, R_DZ, R_ZBOT, R_FESED, R_S, R_ROWS };
// Canonical stage rows of the 30 tracers.  The host chooses the tracer slots (BGC_indices_type), so a
// slot number is a run-time value; the stage is filled through a row -> slot table instead
// (EcoArgs::slot_of_row, built by bgc_capi.cu from the ctx's index tables), and from there on every
// tracer is a compile-time row: one LDS / STS with an immediate offset, no index arithmetic.
//   rows 0..15  the plain tracers in the declaration order of BgcIndices
//   rows 16+3a+{0,1,2}  C, Chl, Fe of functional group a
//   row 28 / 29 the Si tracer of the silicifier / the CaCO3 tracer of the calcifier
enum { po4_row = 0, no3_row, sio3_row, nh4_row, fe_row, o2_row, dic_row, dic_alt_co2_row, alk_row, doc_row,
       don_row, dofe_row, dop_row, dopr_row, donr_row, zooC_row, GROUP_ROW0 = 16, SI_ROW = 28, CA_ROW = 29 };
#define G_C(a_) (GROUP_ROW0 + 3 * (a_))
#define G_CHL(a_) (GROUP_ROW0 + 3 * (a_) + 1)
#define G_FE(a_) (GROUP_ROW0 + 3 * (a_) + 2)
static_assert(G_FE(BGC_AUTOTROPH_CNT - 1) + 1 == SI_ROW && CA_ROW + 1 == BGC_TRACER_CNT, "canonical tracer rows");
enum { X_KMAX = 0, X_ZPHOTO = 4, X_ZNO3 = 8, X_ZCACO3 = 12,
       X_JC = 16, X_JC100, X_JN, X_JN100, X_JP, X_JP100, X_JSI, X_JSI100,
       X_CHL100, X_BSI, X_CACO3ZINT, X_PHOTOCZINT, X_PHOTOCNO3ZINT, X_O2MIN,
       // carried column state that is touched once per level (registers are the scarce resource)
       X_PAROUT, X_QADUST, X_ZBOTKM1, X_O2MINDEPTH, X_DUS, X_DUH, X_ROWS };
static_assert(X_ROWS == 36, "shared-memory budget of the column sweep: 2*R_ROWS + X_ROWS rows of BLOCK doubles");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP); bytes % 16 == 0,
// both addresses 16-byte aligned; completion is counted on the mbarrier.
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 8-byte asynchronous copy global -> shared (SASS: LDGSTS), per thread; groups are committed
// once per level and waited for with cp.async.wait_group
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int DIAG, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
eco_columns_kernel(const __grid_constant__ EcoArgs A) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * BLOCK;
  const int col = col0 + tid;
  const int nL = A.nL, nC = A.nC;
  const bool in_range = col < nC;
  __shared__ unsigned s_tslot;   // (debugging aid; in shared memory: no register is free to hold it)
  if (A.block_trace && tid == 0) s_tslot = block_trace_begin(A.block_trace, 1u);
  // Element indices are 32-bit: one IMAD.WIDE.U32 forms an address from (index, base pointer).
  // The largest index, 30 * nL * nC, stays below 2^32 for any block that fits one GPU's memory
  // (180 GB / 2200 B per cell = 82 M cells); bgc_capi.cu rejects larger blocks.
  const unsigned nLnC = (unsigned)nL * (unsigned)nC;
  // stage rows are PITCH doubles apart: BLOCK columns + the element in front of a run that starts 8 bytes off
  // a 16-byte boundary (odd nC) + the one that rounds the copy up to 16 bytes
  constexpr int PITCH = BLOCK + 2;
  double *const xs = smem + 2 * R_ROWS * PITCH;                       // per-thread scratch rows
  unsigned long long *const bars = (unsigned long long *)(xs + X_ROWS * BLOCK);
#define XS(row) xs[(row) * BLOCK + tid]
#define IN(row) st[(row) * PITCH + tid]

  // The column's depth is compared once or twice per level and would otherwise be spilled to
  // local memory (every register is taken): it lives in a shared-memory row of its own.
  int *const kmax_row = (int *)(xs + X_KMAX * BLOCK);
  {
    int km = (in_range && col < A.nColumns) ? A.kmax[col] : 0;
    if (km > nL) km = nL;
    if (km < 0) km = 0;
    kmax_row[tid] = km;
  }
#define kmax (kmax_row[tid])

  const BgcParams &P = c_eco.p;
  const BgcIndices &I = c_eco.ind;
  const EcoDerived &D = c_der;
  const double epsC = P.epsC, epsTinv = P.epsTinv;
  const double T0K = P.T0_Kelvin_BGC;

  // ---- various k==1 initialisations (BGC_mod.F90:808-814, :2046-2104)
  // P_iron's hard flux is identically zero: it starts at zero and P_iron%gamma = 0 (:2066, :2483)
  constexpr double Fe_h = 0.0;
  double lat = 0.0;
  double POC_s = 0.0, POC_h = 0.0, Ca_s = 0.0, Ca_h = 0.0, Si_s = 0.0, Si_h = 0.0, Fe_s = 0.0;
#pragma unroll
  for (int r = X_ZPHOTO; r < X_ROWS; ++r) XS(r) = 0.0;
  if (kmax > 0) {
    lat = A.lat[col];
    const double dust_in = gmax(0.0, A.dust_flux_in[col]);
    double du_s = 0.0, du_h = 0.0;
    if (dust_in != 0.0) {
      du_s = (1.0 - dust_gamma) * dust_in;
      du_h = dust_gamma * dust_in;
    }
    XS(X_DUS) = du_s;
    XS(X_DUH) = du_h;
    XS(X_QADUST) = dust_rho * (du_s + du_h);
    double PAR_out = gmax(0.0, A.sw_flux[col]);
    PAR_out = PAR_out * f_qsw_par;
    XS(X_PAROUT) = PAR_out;
  }
  const bool north = lat >= 0.0;

  // ---- column integrals / scan state (diagnostics only)


  // ---- inventory (fused stage 1): sum_k tendency*dz.  Every tendency*dz of a level is written
  //      back into the stage row of its own tracer slot (the input has been consumed by then); a
  //      warp-level transpose at the end of the level leaves lane l with the warp's sum of slot l,
  //      so the running inventory of all 30 tracers is ONE register per lane.
  const bool inv = A.inv_partials != nullptr;
  double inv_acc = 0.0;

  // ---- deepest active level of the block: nothing below it is fetched
  __shared__ int s_kmax_blk;
  const bool bulk = A.bulk != 0;   // 16-byte aligned slabs: TMA bulk copies; otherwise 8-byte cp.async per thread
  if (tid == 0) {
    s_kmax_blk = 0;
    if (bulk) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); }
  }
  __syncthreads();
  if (kmax > 0) atomicMax(&s_kmax_blk, kmax);
  if (bulk && tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int kmax_blk = s_kmax_blk;
  const int cols_blk = min(BLOCK, nC - col0);                  // columns of this block
  const bool last_blk = col0 + BLOCK >= nC;

  // Source of every stage row (level 0, this block's first column): the tracer rows go through the
  // row -> slot table, the rest are the caller's separate arrays.
  __shared__ const double *s_rowsrc[R_ROWS];
  if (tid < R_ROWS) {
    const double *p;
    if (tid < BGC_TRACER_CNT) p = A.tracers + (size_t)(A.slot_of_row[tid] - 1) * (size_t)nLnC;
    else if (tid == R_T) p = A.T;
    else if (tid == R_ZMID) p = A.zmid;
    else if (tid == R_DZ) p = A.dz;
    else if (tid == R_ZBOT) p = A.zbot;
    else if (tid == R_FESED) p = A.fesedflux;
    else p = A.S;
    s_rowsrc[tid] = p + col0;
  }
  __syncthreads();

  // Fetch level `kk` of every input array into stage `kk & 1`.  The first lanes of every warp issue a
  // share of the bulk copies (rows w, w + NW, ...: one per lane) so that no single warp carries the whole
  // issue cost; thread 0 also arms the mbarrier with the byte count of the whole level.  (A
  // copy that lands before the arm only drives the transaction count negative for a moment:
  // the phase cannot complete before thread 0's arrival.)  Rows the sweep never reads are not
  // fetched: DIC, ALK (carbonate kernel only), DIC_ALT_CO2 (dead, :748), S without diagnostics.
  constexpr int NW = BLOCK / 32;
  constexpr int N_FETCH_ROWS = DIAG ? R_ROWS : R_S;
  auto row_is_fetched = [](int r) { return r != dic_row && r != dic_alt_co2_row && r != alk_row; };
  auto fetch_level = [&](int kk) {
    double *dst = smem + (size_t)(kk & 1) * R_ROWS * PITCH;
    const size_t off = (size_t)nC * (size_t)kk;
    if (bulk) {
      // The run [off + col0, + cols_blk) of every row starts on a 16-byte boundary or 8 bytes behind one
      // (the row bases are 16-byte aligned, col0 is even): fetch from the boundary, a multiple of 16 bytes.
      // Rounding UP reads one element past the run - the next block's first column, or the next level's
      // first element - which exists except behind the last level of the last block: there the copy is
      // rounded DOWN and the thread of the last column fetches its own element.
      const unsigned mis = (unsigned)((off + (size_t)col0) & 1);
      const unsigned need = (unsigned)cols_blk + mis;
      unsigned n = (need + 1u) & ~1u;
      const bool tail = (n != need) && last_blk && kk == nL - 1;
      if (tail) {
        n = need - 1u;
        if (tid == cols_blk - 1) {
#pragma unroll 1
          for (int r = 0; r < N_FETCH_ROWS; ++r)
            if (row_is_fetched(r)) dst[r * PITCH + mis + tid] = s_rowsrc[r][off + tid];
        }
      }
      // lane j of warp w issues the copy of row w + j * NW: the addresses of a warp's rows are formed side by
      // side in its first lanes and the copies leave in one pass (a loop of lane 0 over its rows cost 130 warp
      // instructions per level, 4 % of the kernel's, for five copies)
      const int j = tid & 31;
      if (j >= (N_FETCH_ROWS + NW - 1) / NW) return;
      const unsigned bar = smem_u32(&bars[kk & 1]);
      const unsigned bytes = n * (unsigned)sizeof(double);
      if (tid == 0) mbar_expect_tx(bar, bytes * (unsigned)(N_FETCH_ROWS - 3));
      if (bytes == 0) return;
      const int r = (tid >> 5) + j * NW;
      if (r < N_FETCH_ROWS && row_is_fetched(r))
        bulk_g2s(smem_u32(dst + r * PITCH), s_rowsrc[r] + off - mis, bytes, bar);
    } else {
      if (in_range) {
#pragma unroll 1
        for (int r = 0; r < N_FETCH_ROWS; ++r)
          if (row_is_fetched(r)) cp_async8(smem_u32(dst + r * PITCH + tid), s_rowsrc[r] + off + tid);
      }
      cp_async_commit();
    }
  };
  // Level 0 is fetched here; level k + 1 is requested from inside level k (FETCH_NEXT below), after the
  // first functional group: right after a block barrier every warp would stall on the issue latency
  // at the same time, while a group later the warps have drifted apart and the other warp of the
  // scheduler has work.  (Requested after the LAST group, half a level ahead, the data was late often
  // enough for the mbarrier wait to collect 2.5 % of the kernel's stall samples.)
  if (kmax_blk > 0) fetch_level(0);

  for (int k = 0; k < nL; ++k) {
    const unsigned i2 = (unsigned)col + (unsigned)nC * (unsigned)k;
    // (bulk staging: the level's run sits one element into its stage rows when it starts 8 bytes off a boundary)
    double *const st = smem + (size_t)(k & 1) * R_ROWS * PITCH + (bulk ? (((unsigned)nC * (unsigned)k + (unsigned)col0) & 1u) : 0u);

    if (bulk) {
      if (k < kmax_blk) mbar_wait(smem_u32(&bars[k & 1]), (unsigned)((k >> 1) & 1));
    } else {
      cp_async_wait_all();       // this thread's own column of level k has landed (it reads no other)
    }
    bool fetched_next = false;
#define FETCH_NEXT() do { if (!fetched_next) { fetched_next = true; if (k + 1 < kmax_blk) fetch_level(k + 1); } } while (0)

    if (k >= kmax) {
      FETCH_NEXT();
      // ---- inactive cell: the reference's whole-array zero fills
      if (inv && k < kmax_blk) {
#pragma unroll
        for (int q = 0; q < BGC_TRACER_CNT; ++q) IN(q) = 0.0;   // this column's share of the level's inventory sums
      }
      if (in_range) {
#pragma unroll 6
        for (int n = 0; n < BGC_TRACER_CNT; ++n) A.tend[i2 + (unsigned)n * nLnC] = 0.0;
        if (DIAG) {
          ECO_DIAG_K2_LIST(ZERO_K2)
          BGC_DIAG_KA_LIST(ZERO_KA)
        }
      }
    } else {

#define TR(row_) gmax(0.0, IN(row_))
#define TEND(row_) A.tend[i2 + A.tend_off[row_]]

    // ---- this level's inputs (setup_loop clamp folded in, :747-783)
    const double TEMP = IN(R_T);
    const double zmid = IN(R_ZMID);
    const double dz = IN(R_DZ);
    const double zbot = IN(R_ZBOT);
    // (the tracers that only the code after the functional-group loop needs are read from the
    //  stage there: they would otherwise sit in registers across the whole loop)
    const double PO4_loc = TR(po4_row), NO3_loc = TR(no3_row), SiO3_loc = TR(sio3_row),
                 NH4_loc = TR(nh4_row), Fe_loc = TR(fe_row), DOP_loc = TR(dop_row),
                 zooC_loc = TR(zooC_row);

    // ---- temperature function, loss thresholds (:1041-1094)
    const double Tfunc = fpow_base(Q_10, kLnQ10, cdiv(((TEMP + T0K) - (Tref + T0K)), 10.0, 0.1));
    double f_loss_thres;
    if (zmid > thres_z1) {
      if (zmid < thres_z2) f_loss_thres = cdiv((thres_z2 - zmid), (thres_z2 - thres_z1), 1.0 / (thres_z2 - thres_z1));
      else                 f_loss_thres = 0.0;
    } else {
      f_loss_thres = 1.0;
    }

    const double ztop = (k > 0) ? XS(X_ZBOTKM1) : 0.0;
    const double pt100 = gmax(gmin(100.0e2 - ztop, dz), 0.0);   // upper-100 m part of this layer (:1880-1885)

    // ---- functional-group tracers: clamp, zero mask (:826-844), Pprime (:1083-1094);
    //      staged in shared memory for the rolled group loop below
    double Chl_sum = 0.0, Chl_100 = 0.0;
    double Pp[NA];   // Pprime of the four groups (the group loops are fully unrolled: registers)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const BgcAutotroph &at = c_eco.a[a];
      double vChl = TR(G_CHL(a)), vC = TR(G_C(a)), vFe = TR(G_FE(a));
      double vSi = (at.Si_ind > 0) ? TR(SI_ROW) : 0.0;
      double vCa = (at.CaCO3_ind > 0) ? TR(CA_ROW) : 0.0;
      bool zero_mask = vChl == 0.0 || vC == 0.0 || vFe == 0.0;
      if (at.Si_ind > 0) zero_mask = zero_mask || vSi == 0.0;
      if (zero_mask) { vChl = 0.0; vC = 0.0; vFe = 0.0; vSi = 0.0; vCa = 0.0; }
      // masked values go back into the stage rows, where the rolled loop picks them up
      IN(G_CHL(a)) = vChl; IN(G_C(a)) = vC; IN(G_FE(a)) = vFe;
      if (at.Si_ind > 0) IN(SI_ROW) = vSi;
      if (at.CaCO3_ind > 0) IN(CA_ROW) = vCa;
      Chl_sum = Chl_sum + vChl;
      if (DIAG) Chl_100 = Chl_100 + vChl * pt100;

      double C_loss_thres = f_loss_thres * at.loss_thres;
      if (at.temp_function == BGC_TFNC_Q10) {
        if (TEMP < at.temp_thres) C_loss_thres = f_loss_thres * at.loss_thres2;
      } else if (at.temp_function == BGC_TFNC_QUASI_MMRT) {
        const double tmpTmax = north ? at.temp_thresN : at.temp_thresS;
        if (TEMP > tmpTmax) C_loss_thres = f_loss_thres * at.loss_thres2;
      }
      Pp[a] = gmax(vC - C_loss_thres, 0.0);
    }
    if (DIAG) XS(X_CHL100) = XS(X_CHL100) + Chl_100;

    // ---- PAR (Morel & Maritorena 2001), :907-924
    const double PAR_in = XS(X_PAROUT);
    double KPARdz;
    {
      const double w = gmax(Chl_sum, 0.02);
      if (w < 0.13224) KPARdz = 0.000919 * fpow(w, 0.3536);
      else             KPARdz = 0.001131 * fpow(w, 0.4562);
    }
    KPARdz = KPARdz * dz;
    const double eKPAR = bexp(-KPARdz);
    const double PAR_out = PAR_in * eKPAR;
    XS(X_PAROUT) = PAR_out;
    const double PAR_avg = fdiv(PAR_in * (1.0 - eKPAR), KPARdz);
    if (DIAG) ST2(diag_PAR_avg, PAR_avg);
    // light factor of the nitrification term (:1545-1556), formed here so that PAR_in, PAR_out and
    // KPARdz need not stay in registers across the functional-group loop
    double nitrif_light = 0.0;
    if (PAR_out < P.parm_nitrif_par_lim) {
      nitrif_light = 1.0;
      if (PAR_in > P.parm_nitrif_par_lim)
        nitrif_light = fdiv(log(fdiv(PAR_out, P.parm_nitrif_par_lim)), (-KPARdz));
    }

    // ---- running sums over the functional groups.  Fortran's SUM(x(:)) adds the
    //      elements left to right starting from zero: so do these accumulators.
    double s_auto_loss_doc = 0.0, s_auto_graze_doc = 0.0, s_auto_graze_poc = 0.0, s_auto_agg = 0.0,
           s_auto_loss_poc = 0.0, s_NO3_V = 0.0, s_NH4_V = 0.0, s_auto_loss_dic = 0.0, s_auto_graze_dic = 0.0,
           s_photoFe = 0.0, s_PO4_V = 0.0, s_auto_graze_zoo = 0.0, s_DOP_V = 0.0, s_photoC = 0.0,
           s_auto_graze = 0.0;
    double zd_num = 0.0, zd_den = 0.0;          // f_zoo_detr (:1395-1401)
    double acc_DOP_prod = 0.0, acc_DOFe_prod = 0.0, acc_Fe_prod = 0.0;
    double acc_t_fe = 0.0, acc_t_nh4 = 0.0, acc_t_sio3 = 0.0, acc_t_po4 = 0.0, acc_t_dic = 0.0;
    double O2_PRODUCTION = 0.0;
    double Ca_prod = 0.0, Si_prod = 0.0;        // last writer wins among qualifying groups (:1480-1498)
    double tot_CaCO3_form = 0.0, tot_Nfix = 0.0;
    double s_tC = 0.0, s_tCaCO3 = 0.0, s_tSi = 0.0, s_QpC = 0.0, s_Nfix_J = 0.0, photoC_NO3_TOT = 0.0;
    double bSi_form_k = 0.0, CaCO3_zint_k = 0.0, NO3_zint_k = 0.0;   // this level's additions to column integrals
    // The zero-biomass shortcut below writes the zeros a group body would compute - which it would
    // only from FINITE factors: with a NaN or Inf among them the reference's 0 * x is NaN, and it
    // must stay NaN here.  One sum tells (non-finite as soon as one term is).
    const bool lvl_finite = fabs(Tfunc + PAR_avg + NO3_loc + NH4_loc + PO4_loc + DOP_loc + Fe_loc + SiO3_loc + zooC_loc + dz)
                            <= 1.7976931348623157e308;

    // ---- per functional group: quotas (:850-898), uptake, photosynthesis, losses,
    //      grazing, routing (:1107-1388), tendencies (:1700-1745)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (a == 1) FETCH_NEXT();   // request level k + 1 (see fetch_level) once the first group is done
      const BgcAutotroph &at = c_eco.a[a];
      const unsigned ia = i2 + (unsigned)a * nLnC;
      const double aChl = IN(G_CHL(a)), aC = IN(G_C(a)), aFe = IN(G_FE(a)), Pprime = Pp[a];
      const bool has_Si = at.Si_ind > 0, has_Ca = at.CaCO3_ind > 0;

      // ---- zero-biomass shortcut.  A group whose Chl, C or Fe is exactly zero has been zeroed
      //      as a whole (:826-844: clamped undershoots and the lightless deep ocean), and then every
      //      product of the group body below is exactly zero: the only things that survive are
      //      the nutrient-limitation diagnostics (they do not involve the biomass), the
      //      epsC*epsTinv terms of f_zoo_detr (:1395-1401) and the running NO3 integral (:1844-1846).
      //      When that holds for every active lane of the warp - and every factor of the level is
      //      finite in every lane (lvl_finite) - the body is skipped; the bits produced are the same
      //      (x*0 = 0 and s+0 = s for finite values).
      if (A.zero_shortcut && !__any_sync(__activemask(), aC != 0.0 || !lvl_finite)) {
        if (DIAG) {
          const double rNO3 = cdiv(NO3_loc, at.kNO3, D.r_kNO3[a]), rNH4 = cdiv(NH4_loc, at.kNH4, D.r_kNH4[a]);
#ifdef BGC_STRICT
          double VNtot = rNO3 / (1.0 + rNO3 + rNH4) + rNH4 / (1.0 + rNO3 + rNH4);
#else
          const double rN = frcp(1.0 + rNO3 + rNH4);
          double VNtot = rNO3 * rN + rNH4 * rN;
#endif
          if (at.Nfixer) VNtot = 1.0;
          const double rPO4 = cdiv(PO4_loc, at.kPO4, D.r_kPO4[a]), rDOP = cdiv(DOP_loc, at.kDOP, D.r_kDOP[a]);
#ifdef BGC_STRICT
          const double VPtot = rPO4 / (1.0 + rPO4 + rDOP) + rDOP / (1.0 + rPO4 + rDOP);
#else
          const double rPd = frcp(1.0 + rPO4 + rDOP);
          const double VPtot = rPO4 * rPd + rDOP * rPd;
#endif
          STA(diag_N_lim, VNtot);
          STA(diag_Fe_lim, fdiv(Fe_loc, Fe_loc + at.kFe));
          STA(diag_P_lim, VPtot);
          STA(diag_SiO3_lim, (at.kSiO3 > 0.0) ? fdiv(SiO3_loc, SiO3_loc + at.kSiO3) : 0.0);
          STA(diag_light_lim, 0.0);
          STA(diag_photoNO3, 0.0); STA(diag_photoNH4, 0.0); STA(diag_PO4_uptake, 0.0); STA(diag_DOP_uptake, 0.0);
          STA(diag_photoFe, 0.0); STA(diag_bSi_form, 0.0); STA(diag_CaCO3_form, 0.0); STA(diag_Nfix, 0.0);
          STA(diag_auto_graze, 0.0); STA(diag_auto_loss, 0.0); STA(diag_auto_agg, 0.0);
          STA(diag_photoC, 0.0); STA(diag_photoC_NO3, 0.0);
          NO3_zint_k = NO3_zint_k + XS(X_ZNO3 + a);
        }
        zd_num = zd_num + at.f_zoo_detr * (0.0 + epsC * epsTinv);
        zd_den = zd_den + (0.0 + epsC * epsTinv);
        if (has_Ca) Ca_prod = 0.0;
        if (has_Si) Si_prod = 0.0;
        TEND(G_C(a)) = 0.0; TEND(G_CHL(a)) = 0.0; TEND(G_FE(a)) = 0.0;
        if (has_Si) TEND(SI_ROW) = 0.0;
        if (has_Ca) TEND(CA_ROW) = 0.0;
        // (inventory: the group's stage rows already hold the zeros the zero mask wrote - aC == 0 implies the mask -
        //  and tendency * dz = 0 is what they have to carry)
        continue;
      }

      const double rCden = frcp(aC + epsC);
#ifdef BGC_STRICT
      const double thetaC = aChl / (aC + epsC), Qfe = aFe / (aC + epsC);
#else
      const double thetaC = aChl * rCden, Qfe = aFe * rCden;
#endif
      double Qsi = 0.0, gQsi = 0.0, QCaCO3 = 0.0;
      if (has_Si) {
#ifdef BGC_STRICT
        Qsi = gmin(IN(SI_ROW) / (aC + epsC), gQsi_max);
#else
        Qsi = gmin(IN(SI_ROW) * rCden, gQsi_max);
#endif
      }
      double gQfe = at.gQfe_0;
      if (Fe_loc < D.cks_kFe[a])
        gQfe = gmax(cdiv(gQfe * Fe_loc, D.cks_kFe[a], D.r_cks_kFe[a]), at.gQfe_min);
      if (has_Si) {
        double g = gQsi_0;
        if ((Fe_loc < D.cksi_kFe[a]) && (Fe_loc > 0.0) && (SiO3_loc > D.cksi_kSiO3[a]))
          g = gmin(fdiv(g * P.cksi * at.kFe, Fe_loc), gQsi_max);
        if (Fe_loc == 0.0) g = gQsi_max;
        if (SiO3_loc < D.cksi_kSiO3[a])
          g = gmax(cdiv(g * SiO3_loc, D.cksi_kSiO3[a], D.r_cksi_kSiO3[a]), gQsi_min);
        gQsi = g;
      }
      if (has_Ca) {
#ifdef BGC_STRICT
        QCaCO3 = IN(CA_ROW) / (aC + epsC);
#else
        QCaCO3 = IN(CA_ROW) * rCden;
#endif
        if (QCaCO3 > QCaCO3_max) QCaCO3 = QCaCO3_max;
      }

      // nutrient limitation (:1107-1150)
      const double rNO3 = cdiv(NO3_loc, at.kNO3, D.r_kNO3[a]), rNH4 = cdiv(NH4_loc, at.kNH4, D.r_kNH4[a]);
#ifdef BGC_STRICT
      const double VNO3 = rNO3 / (1.0 + rNO3 + rNH4);
      const double VNH4 = rNH4 / (1.0 + rNO3 + rNH4);
#else
      const double rN = frcp(1.0 + rNO3 + rNH4);
      const double VNO3 = rNO3 * rN, VNH4 = rNH4 * rN;
#endif
      double VNtot = VNO3 + VNH4;
      if (at.Nfixer) VNtot = 1.0;

      const double VFe = fdiv(Fe_loc, Fe_loc + at.kFe);
      double f_nut = gmin(VNtot, VFe);

      const double rPO4 = cdiv(PO4_loc, at.kPO4, D.r_kPO4[a]), rDOP = cdiv(DOP_loc, at.kDOP, D.r_kDOP[a]);
#ifdef BGC_STRICT
      const double VPO4 = rPO4 / (1.0 + rPO4 + rDOP);
      const double VDOP = rDOP / (1.0 + rPO4 + rDOP);
#else
      const double rPd = frcp(1.0 + rPO4 + rDOP);
      const double VPO4 = rPO4 * rPd, VDOP = rDOP * rPd;
#endif
      const double VPtot = VPO4 + VDOP;
      f_nut = gmin(f_nut, VPtot);

      double VSiO3 = 0.0;
      if (at.kSiO3 > 0.0) {
        VSiO3 = fdiv(SiO3_loc, SiO3_loc + at.kSiO3);
        f_nut = gmin(f_nut, VSiO3);
      }
      if (DIAG) {
        STA(diag_N_lim, VNtot);
        STA(diag_Fe_lim, VFe);
        STA(diag_P_lim, VPtot);
        STA(diag_SiO3_lim, VSiO3);
      }

      // photosynthesis (:1157-1182)
      double PCmax = at.PCref * f_nut * Tfunc;
      if (TEMP < at.temp_thres) PCmax = 0.0;
      if (at.temp_function == BGC_TFNC_QUASI_MMRT) {
        const double tmpTopt = north ? at.temp_optN : at.temp_optS;
        const double tmpTmax = north ? at.temp_thresN : at.temp_thresS;
        PCmax = PCmax * gmin(1.0, cdiv(tmpTmax - TEMP, tmpTmax - tmpTopt, north ? D.r_dTN[a] : D.r_dTS[a]));
        if (TEMP > tmpTmax) PCmax = 0.0;
      }
      const double aPI = at.alphaPI * thetaC * PAR_avg;
      const double light_lim = (1.0 - bexp(fdiv(-1.0 * aPI, PCmax + epsTinv)));
      const double PCphoto = PCmax * light_lim;
      if (DIAG) STA(diag_light_lim, light_lim);

      const double photoC = PCphoto * aC;

      // uptake ratios (:1190-1222)
      double NO3_V, NH4_V, VNC, photoC_NO3, PO4_V, DOP_V;
      if (VNtot > 0.0) {
#ifdef BGC_STRICT
        NO3_V = (VNO3 / VNtot) * photoC * Qn;
        NH4_V = (VNH4 / VNtot) * photoC * Qn;
        photoC_NO3 = (VNO3 / VNtot) * photoC;
#else
        const double rV = frcp(VNtot);
        NO3_V = (VNO3 * rV) * photoC * Qn;
        NH4_V = (VNH4 * rV) * photoC * Qn;
        photoC_NO3 = (VNO3 * rV) * photoC;
#endif
        VNC = PCphoto * Qn;
      } else {
        NO3_V = 0.0; NH4_V = 0.0; VNC = 0.0; photoC_NO3 = 0.0;
      }
      if (VPtot > 0.0) {
#ifdef BGC_STRICT
        PO4_V = (VPO4 / VPtot) * photoC * at.Qp;
        DOP_V = (VDOP / VPtot) * photoC * at.Qp;
#else
        const double rV = frcp(VPtot);
        PO4_V = (VPO4 * rV) * photoC * at.Qp;
        DOP_V = (VDOP * rV) * photoC * at.Qp;
#endif
      } else {
        PO4_V = 0.0; DOP_V = 0.0;
      }
      const double photoFe = photoC * gQfe;

      double photoSi = 0.0;
      if (has_Si) {
        photoSi = photoC * gQsi;
        bSi_form_k = bSi_form_k + photoSi;   // (:1230-1231, no dz)
      }
      if (DIAG) {
        STA(diag_photoNO3, NO3_V);
        STA(diag_photoNH4, NH4_V);
        STA(diag_PO4_uptake, PO4_V);
        STA(diag_DOP_uptake, DOP_V);
        STA(diag_photoFe, photoFe);
        STA(diag_bSi_form, photoSi);
      }

      // Chl synthesis, GD98 (:1240-1246)
      double photoacc = 0.0;
      if (aPI > 0.0) {
        const double pChl = fdiv(at.thetaN_max * PCphoto, aPI);
        photoacc = fdiv(pChl * VNC, thetaC) * aChl;
      }

      // implicit calcification (:1255-1278)
      double CaCO3_PROD = 0.0;
      if (at.imp_calcifier) {
        double cp = P.parm_f_prod_sp_CaCO3 * photoC;
        cp = cp * f_nut;
        if (TEMP < CaCO3_temp_thres1)
          cp = cp * gmax((TEMP - CaCO3_temp_thres2), 0.0) / (CaCO3_temp_thres1 - CaCO3_temp_thres2);
        if (aC > CaCO3_sp_thres)
          cp = gmin((cp * aC / CaCO3_sp_thres), (f_photosp_CaCO3 * photoC));
        CaCO3_PROD = cp;
        tot_CaCO3_form = tot_CaCO3_form + cp;
        if (DIAG) {
          const double w = dz * cp;
          XS(X_ZCACO3 + a) = XS(X_ZCACO3 + a) + w;
          CaCO3_zint_k = CaCO3_zint_k + w;
        }
      }
      if (DIAG) STA(diag_CaCO3_form, CaCO3_PROD);

      // losses and aggregation (:1285-1290)
      const double auto_loss = at.mort * Pprime * Tfunc;
      double auto_agg = gmin(D.agg_max_dps[a] * Pprime, at.mort2 * Pprime * Pprime);
      auto_agg = gmax(D.agg_min_dps[a] * Pprime, auto_agg);

      // grazing (:1297-1324)
      double grazee_C = 0.0;
#pragma unroll
      for (int b = 0; b < NA; ++b)
        if (c_eco.same_grazee[a][b]) grazee_C = grazee_C + Pp[b];

      double z_umax = at.z_umax_0 * Tfunc;
      if (a + 1 == I.diat_ind) {
        if (north && (TEMP > at.temp_optN)) {
          z_umax = z_umax * gmax(cdiv(at.temp_thresN - TEMP, at.temp_thresN - at.temp_optN, D.r_dTN[a]), 0.95);
        } else if ((lat <= 0.0) && (TEMP > at.temp_optS)) {
          z_umax = z_umax * gmax(cdiv(at.temp_thresS - TEMP, at.temp_thresS - at.temp_optS, D.r_dTS[a]), 0.95);
        }
      }
      double auto_graze = 0.0;
      if (grazee_C > 0.0) {
#ifdef BGC_STRICT
        auto_graze = (Pprime / grazee_C) * z_umax * zooC_loc * (grazee_C / (grazee_C + at.z_grz));
#else
        // (P/g) * u * z * (g/(g+zg)) with both quotients formed from reciprocals
        auto_graze = (Pprime * frcp(grazee_C)) * z_umax * zooC_loc * (grazee_C * frcp(grazee_C + at.z_grz));
#endif
      }

      // N fixation (:1331-1338)
      double Nfix = 0.0, Nexcrete = 0.0;
      if (at.Nfixer) {
        const double w = photoC * Qn;
        Nfix = (w * r_Nfix_photo) - NO3_V - NH4_V;
        Nexcrete = Nfix + NO3_V + NH4_V - w;
        tot_Nfix = tot_Nfix + Nfix;
        acc_t_nh4 = acc_t_nh4 + Nexcrete;
        s_Nfix_J = s_Nfix_J + Nfix;
      }
      if (DIAG) STA(diag_Nfix, Nfix);

      // routing (:1354-1372)
      const double auto_graze_zoo = at.graze_zoo * auto_graze;
      double auto_graze_poc, auto_loss_poc;
      if (at.imp_calcifier) {
        auto_graze_poc = auto_graze * gmax((caco3_poc_min * QCaCO3),
                                           gmin(spc_poc_fac * gmax(1.0, Pprime), f_graze_sp_poc_lim));
        auto_loss_poc = QCaCO3 * auto_loss;
      } else {
        auto_graze_poc = at.graze_poc * auto_graze;
        auto_loss_poc = at.loss_poc * auto_loss;
      }
      const double auto_graze_doc = at.graze_doc * auto_graze;
      const double auto_graze_dic = auto_graze - (auto_graze_zoo + auto_graze_poc + auto_graze_doc);
      const double auto_loss_doc = (1.0 - P.parm_labile_ratio) * (auto_loss - auto_loss_poc);
      const double auto_loss_dic = P.parm_labile_ratio * (auto_loss - auto_loss_poc);

      // P routing for groups whose Qp differs from Qp_zoo_pom (:1380-1386, :1434-1440, :1668-1674)
      if (at.Qp != Qp_zoo_pom) {
        const double remaining_P = ((auto_graze + auto_loss + auto_agg) * at.Qp)
                                 - ((auto_graze_zoo) * Qp_zoo_pom)
                                 - ((auto_graze_poc + auto_loss_poc + auto_agg) * Qp_zoo_pom);
        acc_DOP_prod = acc_DOP_prod + (1.0 - P.parm_labile_ratio) * remaining_P;
        acc_t_po4 = acc_t_po4 + P.parm_labile_ratio * remaining_P;
      } else {
        acc_DOP_prod = acc_DOP_prod + at.Qp * (auto_loss_doc + auto_graze_doc);
        acc_t_po4 = acc_t_po4 + at.Qp * (auto_loss_dic + auto_graze_dic);
      }

      // running sums
      s_auto_loss_doc = s_auto_loss_doc + auto_loss_doc;
      s_auto_graze_doc = s_auto_graze_doc + auto_graze_doc;
      s_auto_graze_poc = s_auto_graze_poc + auto_graze_poc;
      s_auto_agg = s_auto_agg + auto_agg;
      s_auto_loss_poc = s_auto_loss_poc + auto_loss_poc;
      s_NO3_V = s_NO3_V + NO3_V;
      s_NH4_V = s_NH4_V + NH4_V;
      s_auto_loss_dic = s_auto_loss_dic + auto_loss_dic;
      s_auto_graze_dic = s_auto_graze_dic + auto_graze_dic;
      s_photoFe = s_photoFe + photoFe;
      s_PO4_V = s_PO4_V + PO4_V;
      s_auto_graze_zoo = s_auto_graze_zoo + auto_graze_zoo;
      s_DOP_V = s_DOP_V + DOP_V;
      s_photoC = s_photoC + photoC;
      s_auto_graze = s_auto_graze + auto_graze;
      zd_num = zd_num + at.f_zoo_detr * (auto_graze + epsC * epsTinv);
      zd_den = zd_den + (auto_graze + epsC * epsTinv);
      acc_DOFe_prod = acc_DOFe_prod + Qfe * (auto_loss_doc + auto_graze_doc);
      acc_Fe_prod = acc_Fe_prod + Qfe * (auto_agg + auto_graze_poc + auto_loss_poc);
      acc_t_fe = acc_t_fe + (Qfe * (auto_loss_dic + auto_graze_dic)) + auto_graze_zoo * (Qfe - Qfe_zoo);
      if (has_Ca) {
        Ca_prod = ((1.0 - f_graze_CaCO3_remin) * auto_graze + auto_loss + auto_agg) * QCaCO3;
        acc_t_dic = acc_t_dic + f_graze_CaCO3_remin * auto_graze * QCaCO3 - CaCO3_PROD;
      }
      if (has_Si) {
        Si_prod = Qsi * ((1.0 - f_graze_si_remin) * auto_graze + auto_agg + at.loss_poc * auto_loss);
        acc_t_sio3 = acc_t_sio3 - photoSi +
                     Qsi * (f_graze_si_remin * auto_graze + (1.0 - at.loss_poc) * auto_loss);
      }

      // O2 production (:1752-1775)
      if (photoC > 0.0) {
        if (!at.Nfixer) {
#ifdef BGC_STRICT
          const double den = NO3_V + NH4_V;
          O2_PRODUCTION = O2_PRODUCTION + photoC *
              ((NO3_V / den) / parm_Red_D_C_O2 + (NH4_V / den) / parm_Remin_D_C_O2);
#else
          const double rden = frcp(NO3_V + NH4_V);
          O2_PRODUCTION = O2_PRODUCTION + photoC *
              ((NO3_V * rden) * (1.0 / parm_Red_D_C_O2) + (NH4_V * rden) * (1.0 / parm_Remin_D_C_O2));
#endif
        } else {
#ifdef BGC_STRICT
          const double den = NO3_V + NH4_V + Nfix;
          O2_PRODUCTION = O2_PRODUCTION + photoC *
              ((NO3_V / den) / parm_Red_D_C_O2 + (NH4_V / den) / parm_Remin_D_C_O2 +
               (Nfix / den) / parm_Red_D_C_O2_diaz);
#else
          const double rden = frcp(NO3_V + NH4_V + Nfix);
          O2_PRODUCTION = O2_PRODUCTION + photoC *
              ((NO3_V * rden) * (1.0 / parm_Red_D_C_O2) + (NH4_V * rden) * (1.0 / parm_Remin_D_C_O2) +
               (Nfix * rden) * (1.0 / parm_Red_D_C_O2_diaz));
#endif
        }
      }

      // the group's own tendencies (:1700-1745)
      {
        const double w = auto_graze + auto_loss + auto_agg;
        const double t_autoC = photoC - w;
        const double t_autoChl = photoacc - thetaC * w, t_autoFe = photoFe - Qfe * w;
        TEND(G_C(a)) = t_autoC;
        TEND(G_CHL(a)) = t_autoChl;
        TEND(G_FE(a)) = t_autoFe;
        s_tC = s_tC + t_autoC;
        s_QpC = s_QpC + at.Qp * t_autoC;
        if (inv) {   // the group's inputs have been read: their rows carry tendency*dz from here on
          IN(G_CHL(a)) = t_autoChl * dz;
          IN(G_C(a)) = t_autoC * dz;
          IN(G_FE(a)) = t_autoFe * dz;
        }
        if (has_Si) {
          const double t = photoSi - Qsi * w;
          TEND(SI_ROW) = t;
          s_tSi = s_tSi + t;
          if (inv) IN(SI_ROW) = t * dz;
        }
        if (has_Ca) {
          const double t = CaCO3_PROD - QCaCO3 * w;
          TEND(CA_ROW) = t;
          s_tCaCO3 = s_tCaCO3 + t;
          if (inv) IN(CA_ROW) = t * dz;
        }
      }

      if (DIAG) {   // :1815-1846
        STA(diag_auto_graze, auto_graze);
        STA(diag_auto_loss, auto_loss);
        STA(diag_auto_agg, auto_agg);
        STA(diag_photoC, photoC);
        STA(diag_photoC_NO3, photoC_NO3);
        XS(X_ZPHOTO + a) = XS(X_ZPHOTO + a) + dz * photoC;
        const double zn = XS(X_ZNO3 + a) + photoC_NO3 * dz;
        XS(X_ZNO3 + a) = zn;
        photoC_NO3_TOT = photoC_NO3_TOT + photoC_NO3;
        // adds the RUNNING per-group integral every level (:1844-1846)
        NO3_zint_k = NO3_zint_k + zn;
      }
    }   // functional groups

    FETCH_NEXT();   // (a table with a single functional group gets here first)
    if (DIAG) {   // (every diagnostic is stored as soon as its value is final: short live ranges, no store bursts)
      ST2(diag_tot_Nfix, tot_Nfix);
      ST2(diag_tot_CaCO3_form, tot_CaCO3_form);
      ST2(diag_auto_graze_TOT, s_auto_graze);
      ST2(diag_photoC_TOT, s_photoC);
      ST2(diag_photoC_NO3_TOT, photoC_NO3_TOT);
      ST2(diag_O2_PRODUCTION, O2_PRODUCTION);
    }

    const double O2_loc = TR(o2_row), DOC_loc = TR(doc_row), DON_loc = TR(don_row),
                 DOFe_loc = TR(dofe_row), DOPr_loc = TR(dopr_row), DONr_loc = TR(donr_row);
    const double fesed = IN(R_FESED);

    // ---- zooplankton routing (:1395-1415)
    const double f_zoo_detr = fdiv(zd_num, zd_den);
    const double Zprime = gmax(zooC_loc - f_loss_thres * loss_thres_zoo, 0.0);
    const double zoo_loss = (P.parm_z_mort2_0 * fpow15(Zprime) + P.parm_z_mort_0 * Zprime) * Tfunc;
    const double zoo_loss_doc = (1.0 - P.parm_labile_ratio) * (1.0 - f_zoo_detr) * zoo_loss;
    const double zoo_loss_dic = P.parm_labile_ratio * (1.0 - f_zoo_detr) * zoo_loss;
    if (DIAG) ST2(diag_zoo_loss, zoo_loss);

    // ---- DOM (:1421-1461)
    const double DOC_prod = zoo_loss_doc + s_auto_loss_doc + s_auto_graze_doc;
    const double DON_prod = Qn * DOC_prod;
    const double DOP_prod = Qp_zoo_pom * zoo_loss_doc + acc_DOP_prod;
    const double DOFe_prod = Qfe_zoo * zoo_loss_doc + acc_DOFe_prod;
    if (DIAG) { ST2(diag_DOC_prod, DOC_prod); ST2(diag_DON_prod, DON_prod); ST2(diag_DOP_prod, DOP_prod); ST2(diag_DOFe_prod, DOFe_prod); }

    double DOC_remin = DOC_loc * DOC_reminR;
    double DON_remin = DON_loc * DON_reminR;
    double DOFe_remin = DOFe_loc * DOFe_reminR;
    double DOP_remin = DOP_loc * DOP_reminR;
    double DONr_remin, DOPr_remin;
    if (PAR_avg > 1.0) {
      DONr_remin = DONr_loc * DONr_reminR;
      DOPr_remin = DOPr_loc * DOPr_reminR;
    } else {
      DONr_remin = DONr_loc * (1.0 / (365.0 * 670.0)) * dps;
      DOPr_remin = DOPr_loc * (1.0 / (365.0 * 460.0)) * dps;
      DOC_remin = DOC_remin * 0.0685;
      DON_remin = DON_remin * 0.1;
      DOFe_remin = DOFe_remin * 0.05;
      DOP_remin = DOP_remin * 0.05;
    }

    if (DIAG) { ST2(diag_DOC_remin, DOC_remin); ST2(diag_DON_remin, DON_remin); ST2(diag_DOP_remin, DOP_remin); ST2(diag_DOFe_remin, DOFe_remin); }

    // ---- particle production (:1467-1529)
    const double POC_prod = f_zoo_detr * zoo_loss + s_auto_graze_poc + s_auto_agg + s_auto_loss_poc;

    double Fe_scavenge_rate = P.parm_fe_scavenge_rate0;
    Fe_scavenge_rate = Fe_scavenge_rate *
        ((POC_s + POC_h) * 120.1 +
         (Ca_s + Ca_h) * CaCO3_mass +
         (Si_s + Si_h) * SiO2_mass +
         (XS(X_DUS) + XS(X_DUH)) * P.dust_fescav_scale);
    if (Fe_loc > fe_scavenge_thres1)
      Fe_scavenge_rate = Fe_scavenge_rate + (Fe_loc - fe_scavenge_thres1) * fe_max_scale2;
    const double Fe_scavenge = yps * Fe_loc * Fe_scavenge_rate;
    const double Fe_prod = ((zoo_loss * f_zoo_detr * Qfe_zoo) + Fe_scavenge) + acc_Fe_prod;
    if (DIAG) { ST2(diag_Fe_scavenge, Fe_scavenge); ST2(diag_Fe_scavenge_rate, Fe_scavenge_rate); }

    // =====================================================================
    // compute_particulate_terms (BGC_mod.F90:2116-2699) for this level
    // =====================================================================
    const double Ca_s_in = Ca_s, Ca_h_in = Ca_h, Si_s_in = Si_s, Si_h_in = Si_h,
                 du_s_in = XS(X_DUS), du_h_in = XS(X_DUH), POC_s_in = POC_s, POC_h_in = POC_h,
                 Fe_s_in = Fe_s, Fe_h_in = Fe_h;
    double du_s, du_h;
    double POC_sed = 0.0, Ca_sed = 0.0, Si_sed = 0.0, du_sed = 0.0, Fe_sed = 0.0;
    double SED_DENITRIF = 0.0, OTHER_REMIN = 0.0;
    double POC_remin, Ca_remin, Si_remin, du_remin, Fe_remin;
    if (DIAG) {   // :2637-2694, the part that is known on entry
      ST2(diag_POC_FLUX_IN, POC_s_in + POC_h_in);
      ST2(diag_POC_PROD, POC_prod);
      ST2(diag_CaCO3_FLUX_IN, Ca_s_in + Ca_h_in);
      ST2(diag_CaCO3_PROD, Ca_prod);
      ST2(diag_SiO2_FLUX_IN, Si_s_in + Si_h_in);
      ST2(diag_SiO2_PROD, Si_prod);
      ST2(diag_dust_FLUX_IN, du_s_in + du_h_in);
      ST2(diag_P_iron_FLUX_IN, Fe_s_in + Fe_h_in);
      ST2(diag_P_iron_PROD, Fe_prod);
    }
    {
      double scalelength;   // piecewise-linear in zbot, :2273-2286
      if (zbot < P.parm_scalelen_z[0]) {
        scalelength = P.parm_scalelen_vals[0];
      } else if (zbot >= P.parm_scalelen_z[3]) {
        scalelength = P.parm_scalelen_vals[3];
      } else {
        scalelength = 0.0;
#pragma unroll
        for (int n = 3; n >= 1; --n) {   // first n (ascending) with zbot < z[n]  <=>  last assignment descending
          if (zbot < P.parm_scalelen_z[n])
            scalelength = P.parm_scalelen_vals[n - 1] +
                          cdiv((P.parm_scalelen_vals[n] - P.parm_scalelen_vals[n - 1]) *
                                   (zbot - P.parm_scalelen_z[n - 1]),
                               (P.parm_scalelen_z[n] - P.parm_scalelen_z[n - 1]), D.r_scalelen_dz[n]);
        }
      }

      const double DECAY_Hard = bexp(cdiv(-dz, 4.0e6, 1.0 / 4.0e6));
      const double DECAY_HardDust = bexp(cdiv(-dz, 1.2e7, 1.0 / 1.2e7));
      const double TfuncS = Tfunc;   // 1.5**(same exponent) (:2295) is bit-identical to Tfunc (:1041)

      const double dzr = frcp(dz);

      double poc_diss = P.parm_POC_diss;
      if ((O2_loc >= 5.0) && (O2_loc < 40.0)) {
        poc_diss = P.parm_POC_diss * (1.0 + cdiv((3.3 - 1.0) * (40.0 - O2_loc), 35.0, 1.0 / 35.0));
      } else if (O2_loc < 5.0) {
        poc_diss = P.parm_POC_diss * 3.3;
      }
      poc_diss = scalelength * poc_diss;
      double sio2_diss = scalelength * P.parm_SiO2_diss;
      const double caco3_diss = scalelength * P.parm_CaCO3_diss;
      const double dust_diss = scalelength * dust_diss0;
      sio2_diss = fdiv(sio2_diss, TfuncS);

      const double decay_POC_E = bexp(fdiv(-dz, poc_diss));
      const double decay_SiO2 = bexp(fdiv(-dz, sio2_diss));
      const double decay_CaCO3 = bexp(fdiv(-dz, caco3_diss));
      const double decay_dust = bexp(fdiv(-dz, dust_diss));

      Ca_s = Ca_s_in * decay_CaCO3 + Ca_prod * ((1.0 - CaCO3_gamma) * (1.0 - decay_CaCO3) * caco3_diss);
      Ca_h = Ca_h_in * DECAY_Hard + Ca_prod * (CaCO3_gamma * dz);
      Si_s = Si_s_in * decay_SiO2 + Si_prod * ((1.0 - SiO2_gamma) * (1.0 - decay_SiO2) * sio2_diss);
      Si_h = Si_h_in * DECAY_Hard + Si_prod * (SiO2_gamma * dz);
      du_s = du_s_in * decay_dust;
      du_h = du_h_in * DECAY_HardDust;

      double POC_PROD_avail = POC_prod - CaCO3_rho * Ca_prod - SiO2_rho * Si_prod;
      if (POC_PROD_avail < 0.0 && A.status) atomicAdd(&A.status[2], 1ull);   // computed and never reported by the reference (:2381-2383)

      double new_QA_dust_def;
      const double QA_dust_def = XS(X_QADUST);
      if (QA_dust_def > 0.0) {
        new_QA_dust_def = fdiv(QA_dust_def * (du_s + du_h), (du_s_in + du_h_in));
      } else {
        new_QA_dust_def = 0.0;
      }
      if (new_QA_dust_def > 0.0) {
        new_QA_dust_def = new_QA_dust_def - POC_PROD_avail * dz;
        if (new_QA_dust_def < 0.0) {
          POC_PROD_avail = -new_QA_dust_def * dzr;
          new_QA_dust_def = 0.0;
        } else {
          POC_PROD_avail = 0.0;
        }
      }
      XS(X_QADUST) = new_QA_dust_def;

      if (POC_h_in == 0.0 && POC_prod == 0.0) {
        POC_h = 0.0;
      } else {
        POC_h = CaCO3_rho * (Ca_s + Ca_h) + SiO2_rho * (Si_s + Si_h) + dust_rho * (du_s + du_h) -
                new_QA_dust_def;
        POC_h = gmax(POC_h, 0.0);
      }
      POC_s = POC_s_in * decay_POC_E + POC_PROD_avail * ((1.0 - decay_POC_E) * poc_diss);

      Ca_remin = Ca_prod + ((Ca_s_in - Ca_s) + (Ca_h_in - Ca_h)) * dzr;
      Si_remin = Si_prod + ((Si_s_in - Si_s) + (Si_h_in - Si_h)) * dzr;
      POC_remin = POC_prod + ((POC_s_in - POC_s) + (POC_h_in - POC_h)) * dzr;
      du_remin = ((du_s_in - du_s) + (du_h_in - du_h)) * dzr;

      if (POC_s_in + POC_h_in == 0.0) {
        Fe_remin = (POC_remin * parm_Red_Fe_C);
      } else {
        Fe_remin = fdiv(POC_remin * (Fe_s_in + Fe_h_in), (POC_s_in + POC_h_in));
      }
      Fe_remin = Fe_remin + (Fe_s_in * 1.5e-5);
      Fe_s = Fe_s_in + dz * ((1.0 - P_iron_gamma) * Fe_prod - Fe_remin);
      if (Fe_s < 0.0) {
        Fe_s = 0.0;
        Fe_remin = Fe_s_in * dzr + (1.0 - P_iron_gamma) * Fe_prod;
      }
      Fe_remin = Fe_remin + du_remin * dust_to_Fe + (fesed * dzr);

      if (k == kmax - 1) {   // bottom cell: burial, sediment denitrification (:2522-2631)
        double flux = POC_s + POC_h;
        if (flux > 0.0) {
          double flux_alt = flux * mpercm * spd;
          POC_sed = flux * gmin(0.8, P.parm_POMbury *
                                         (0.013 + fdiv(0.53 * flux_alt * flux_alt,
                                                       ((7.0 + flux_alt) * (7.0 + flux_alt)))));
          SED_DENITRIF = dzr * flux * (0.06 + 0.19 * fpow_base(0.99, kLn099, (O2_loc - NO3_loc)));
          if (NO3_loc < 5.0) SED_DENITRIF = 0.0;
          flux_alt = flux * 1.0e-6 * spd * 365.0;
          OTHER_REMIN = dzr * gmin(gmin(0.1 + flux_alt, 0.5) * (flux - POC_sed),
                                   (flux - POC_sed - (SED_DENITRIF * dz * denitrif_C_N)));
          if (O2_loc < 1.0) OTHER_REMIN = dzr * (flux - POC_sed - (SED_DENITRIF * dz * denitrif_C_N));
        }

        flux = Si_s + Si_h;
        {
          const double flux_alt = flux * mpercm * spd;
          Si_sed = (flux_alt > 2.0) ? 0.2 : 0.04;
          Si_sed = flux * P.parm_BSIbury * Si_sed;
        }
        if (zbot < 3300.0e2) Ca_sed = Ca_s + Ca_h;

        flux = Ca_s + Ca_h;
        if (flux > 0.0) Ca_remin = Ca_remin + ((flux - Ca_sed) * dzr);
        flux = Si_s + Si_h;
        if (flux > 0.0) Si_remin = Si_remin + ((flux - Si_sed) * dzr);
        flux = POC_s + POC_h;
        if (flux > 0.0) POC_remin = POC_remin + ((flux - POC_sed) * dzr);

        flux = (Fe_s + Fe_h);
        if (flux > 0.0) Fe_sed = flux;
        du_sed = du_s + du_h;

        Ca_s = 0.0; Ca_h = 0.0; Si_s = 0.0; Si_h = 0.0; du_s = 0.0; du_h = 0.0;
        POC_s = 0.0; POC_h = 0.0; Fe_s = 0.0;
      }
      XS(X_DUS) = du_s;
      XS(X_DUH) = du_h;

      if (DIAG) {   // :2637-2694
        ST2(diag_POC_REMIN, POC_remin);
        ST2(diag_CaCO3_REMIN, Ca_remin);
        ST2(diag_SiO2_REMIN, Si_remin);
        ST2(diag_dust_REMIN, du_remin);
        ST2(diag_P_iron_REMIN, Fe_remin);
        ST2(diag_calcToSed, Ca_sed);
        ST2(diag_bsiToSed, Si_sed);
        ST2(diag_pocToSed, POC_sed);
        ST2(diag_SedDenitrif, SED_DENITRIF * dz);
        ST2(diag_OtherRemin, OTHER_REMIN * dz);
        ST2(diag_ponToSed, (POC_sed * Qn));
        ST2(diag_popToSed, (POC_sed * Qp_zoo_pom));
        ST2(diag_dustToSed, du_sed);
        ST2(diag_pfeToSed, Fe_sed);
      }
    }

    // ---- nitrification / denitrification (:1545-1577)
    double RESTORE_NO3 = 0.0, RESTORE_SiO3 = 0.0, RESTORE_PO4 = 0.0;
    if (A.any_restore) {   // (one flag test per cell while restoring is off: the reference never switches it on, Q8)
      if (P.lrest_no3) RESTORE_NO3 = A.rtau[i2] * (A.no3_clim[i2] - NO3_loc);
      if (P.lrest_sio3) RESTORE_SiO3 = A.rtau[i2] * (A.sio3_clim[i2] - SiO3_loc);
      if (P.lrest_po4) RESTORE_PO4 = A.rtau[i2] * (A.po4_clim[i2] - PO4_loc);
    }

    const double NITRIF = (P.parm_kappa_nitrif * NH4_loc) * nitrif_light;

    double DENITRIF;
    {
      double w = cdiv(((P.parm_o2_min + P.parm_o2_min_delta) - O2_loc), P.parm_o2_min_delta, D.r_o2_min_delta);
      w = gmin(gmax(w, 0.0), 1.0);
      if (NO3_loc == 0.0) w = 0.0;
      DENITRIF = w * (cdiv((DOC_remin + POC_remin - OTHER_REMIN), denitrif_C_N, 1.0 / denitrif_C_N) - SED_DENITRIF);
    }

    // ---- tendencies (:1583-1790)
    const double t_no3 = RESTORE_NO3 + NITRIF - DENITRIF - SED_DENITRIF - s_NO3_V;
    const double t_nh4 = (-s_NH4_V - NITRIF + DON_remin + DONr_remin +
                          Qn * (zoo_loss_dic + s_auto_loss_dic + s_auto_graze_dic + POC_remin * (1.0 - DONrefract))) +
                         acc_t_nh4;
    const double t_fe = (Fe_remin + (Qfe_zoo * zoo_loss_dic) + DOFe_remin - s_photoFe - Fe_scavenge) + acc_t_fe;
    const double t_sio3 = (RESTORE_SiO3 + Si_remin) + acc_t_sio3;
    const double t_po4 = (RESTORE_PO4 + DOP_remin + DOPr_remin - s_PO4_V +
                          Qp_zoo_pom * ((1.0 - DOPrefract) * POC_remin + zoo_loss_dic)) + acc_t_po4;
    const double t_zooC = s_auto_graze_zoo - zoo_loss;
    const double t_doc = DOC_prod - DOC_remin;
    const double t_don = (DON_prod * (1.0 - DONrefract)) - DON_remin;
    const double t_donr = (DON_prod * DONrefract) - DONr_remin + (POC_remin * DONrefract * Qn);
    const double t_dop = (DOP_prod * (1.0 - DOPrefract)) - DOP_remin - s_DOP_V;
    const double t_dopr = (DOP_prod * DOPrefract) - DOPr_remin + (POC_remin * DOPrefract * Qp_zoo_pom);
    const double t_dofe = DOFe_prod - DOFe_remin;
    const double t_dic = (s_auto_loss_dic + s_auto_graze_dic - s_photoC + DOC_remin + POC_remin + zoo_loss_dic +
                          Ca_remin) + acc_t_dic;
    const double t_dic_alt = A.alt_co2_use_eco ? t_dic : 0.0;
    const double t_alk = (-t_no3 + t_nh4 + 2.0 * Ca_remin) + 2.0 * acc_t_dic;

    double O2_CONSUMPTION;
    {
      double w = cdiv((O2_loc - P.parm_o2_min), P.parm_o2_min_delta, D.r_o2_min_delta);
      w = gmin(gmax(w, 0.0), 1.0);
      O2_CONSUMPTION = w * (cdiv((POC_remin + DOC_remin - (SED_DENITRIF * denitrif_C_N) - OTHER_REMIN +
                                  zoo_loss_dic + s_auto_loss_dic + s_auto_graze_dic),
                                 parm_Remin_D_C_O2, 1.0 / parm_Remin_D_C_O2) +
                            (2.0 * NITRIF));
    }
    const double t_o2 = O2_PRODUCTION - O2_CONSUMPTION;

    {   // BgcStatus.nonfinite: NaN / Inf in any tendency of the cell shows up in their sum
      const double chk = (((t_no3 + t_nh4) + (t_fe + t_sio3)) + ((t_po4 + t_zooC) + (t_doc + t_don))) +
                         (((t_donr + t_dop) + (t_dopr + t_dofe)) + ((t_dic + t_alk) + (t_o2 + s_tC))) +
                         (s_tCaCO3 + s_tSi);
      if (!(fabs(chk) <= 1.7976931348623157e308) && A.status) atomicAdd(&A.status[3], 1ull);
    }
    TEND(no3_row) = t_no3;
    TEND(nh4_row) = t_nh4;
    TEND(fe_row) = t_fe;
    TEND(sio3_row) = t_sio3;
    TEND(po4_row) = t_po4;
    TEND(zooC_row) = t_zooC;
    TEND(doc_row) = t_doc;
    TEND(don_row) = t_don;
    TEND(donr_row) = t_donr;
    TEND(dop_row) = t_dop;
    TEND(dopr_row) = t_dopr;
    TEND(dofe_row) = t_dofe;
    TEND(dic_row) = t_dic;
    TEND(dic_alt_co2_row) = t_dic_alt;
    TEND(alk_row) = t_alk;
    TEND(o2_row) = t_o2;
    if (inv) {   // every tracer input of this level has been consumed (see the group loop)
      IN(no3_row) = t_no3 * dz; IN(nh4_row) = t_nh4 * dz; IN(fe_row) = t_fe * dz;
      IN(sio3_row) = t_sio3 * dz; IN(po4_row) = t_po4 * dz; IN(zooC_row) = t_zooC * dz;
      IN(doc_row) = t_doc * dz; IN(don_row) = t_don * dz; IN(donr_row) = t_donr * dz;
      IN(dop_row) = t_dop * dz; IN(dopr_row) = t_dopr * dz; IN(dofe_row) = t_dofe * dz;
      IN(dic_row) = t_dic * dz; IN(dic_alt_co2_row) = t_dic_alt * dz;
      IN(alk_row) = t_alk * dz; IN(o2_row) = t_o2 * dz;
    }

    // ---- diagnostics and column integrals (:1796-1945)
    if (DIAG) {
      ST2(diag_NO3_RESTORE, RESTORE_NO3);
      ST2(diag_SiO3_RESTORE, RESTORE_SiO3);
      ST2(diag_PO4_RESTORE, RESTORE_PO4);
      ST2(diag_NITRIF, NITRIF);
      ST2(diag_DENITRIF, DENITRIF);
      ST2(diag_O2_CONSUMPTION, O2_CONSUMPTION);
      if (HAS(diag_AOU)) {   // O2SAT_singleValue, Garcia & Gordon 1992 (:3012-3083)
        const double SALT = IN(R_S);
        const double TS = blog(fdiv(((T0K + 25.0) - TEMP), (T0K + TEMP)));
        double o2sat = bexp(2.00907 + TS * (3.22014 + TS * (4.05010 + TS * (4.94457 + TS * (-2.56847E-1 + TS * 3.88767)))) +
                           SALT * ((-6.24523E-3 + TS * (-7.37614E-3 + TS * (-1.03410E-2 + TS * -8.17083E-3))) +
                                   SALT * -4.88682E-7));
        o2sat = cdiv(o2sat, 0.0223916, 1.0 / 0.0223916);
        A.d.diag_AOU[i2] = o2sat - O2_loc;
      }
      XS(X_PHOTOCZINT) = XS(X_PHOTOCZINT) + s_photoC * dz;
      XS(X_PHOTOCNO3ZINT) = XS(X_PHOTOCNO3ZINT) + NO3_zint_k;
      XS(X_BSI) = XS(X_BSI) + bSi_form_k;
      XS(X_CACO3ZINT) = XS(X_CACO3ZINT) + CaCO3_zint_k;


      const bool shallow = zbot <= 100.0e2;

      double w1 = (t_dic + t_doc + t_zooC + s_tC) + s_tCaCO3;
      XS(X_JC) = XS(X_JC) + w1 * dz + POC_sed + Ca_sed;
      XS(X_JC100) = XS(X_JC100) + w1 * pt100 + (shallow ? (POC_sed + Ca_sed) : 0.0);

      w1 = t_no3 + t_nh4 + t_don + t_donr + Qn * t_zooC + Qn * s_tC;
      w1 = (w1 + DENITRIF + SED_DENITRIF) - s_Nfix_J;
      XS(X_JN) = XS(X_JN) + w1 * dz + POC_sed * Qn;
      XS(X_JN100) = XS(X_JN100) + w1 * pt100 + (shallow ? (POC_sed * Qn) : 0.0);

      w1 = (t_po4 + t_dop + t_dopr + Qp_zoo_pom * t_zooC) + s_QpC;
      XS(X_JP) = XS(X_JP) + w1 * dz + POC_sed * Qp_zoo_pom;
      XS(X_JP100) = XS(X_JP100) + w1 * pt100 + (shallow ? (POC_sed * Qp_zoo_pom) : 0.0);

      w1 = t_sio3 + s_tSi;
      XS(X_JSI) = XS(X_JSI) + w1 * dz + Si_sed;
      XS(X_JSI100) = XS(X_JSI100) + w1 * pt100 + (shallow ? Si_sed : 0.0);

      // O2 minimum scan (:1954-1968)
      if (k == 0 || O2_loc < XS(X_O2MIN)) { XS(X_O2MIN) = O2_loc; XS(X_O2MINDEPTH) = IN(R_ZMID); }
    }
    XS(X_ZBOTKM1) = zbot;
#undef TR
#undef TEND
    }   // active cell
    if (inv && k < kmax_blk) {   // block-uniform condition
      // Transpose through shared memory: lane l adds row l (= tracer slot l) over the warp's 32
      // columns in the order c ^ l (skewed: conflict-free, and one XOR per address since the warp's
      // columns start on a 256-byte boundary of the row); fixed order, no atomics.
      __syncwarp();
      const int lane = tid & 31;
      if (lane < BGC_TRACER_CNT) {
        const double *src = st + lane * PITCH + (tid & ~31);
        double d = 0.0;
#pragma unroll
        for (int c = 0; c < 32; ++c) d += src[c ^ lane];
        inv_acc += d;
      }
    }

    // ---- end of level: every thread is done with stage k&1 (generic-proxy reads and the
    //      in-place mask writes) before the TMA unit refills it with level k+2.  The same
    //      barrier keeps the block's warps on one stretch of code (I-cache).
    FETCH_NEXT();   // (not reached earlier only if a branch above was changed without its FETCH_NEXT)
#undef FETCH_NEXT
    if (bulk) fence_proxy_async();
    __syncthreads();
  }   // level loop

  if (A.block_trace && tid == 0) block_trace_end(A.block_trace, s_tslot);   // (after the last level's barrier)
  // ---- per-column diagnostics
  if (DIAG && in_range) {
    if (kmax > 0) {
      STC(diag_photoC_TOT_zint, XS(X_PHOTOCZINT));
      STC(diag_photoC_NO3_TOT_zint, XS(X_PHOTOCNO3ZINT));
      STC(diag_Jint_Ctot, XS(X_JC));       STC(diag_Jint_100m_Ctot, XS(X_JC100));
      STC(diag_Jint_Ntot, XS(X_JN));       STC(diag_Jint_100m_Ntot, XS(X_JN100));
      STC(diag_Jint_Ptot, XS(X_JP));       STC(diag_Jint_100m_Ptot, XS(X_JP100));
      STC(diag_Jint_Sitot, XS(X_JSI));     STC(diag_Jint_100m_Sitot, XS(X_JSI100));
      STC(diag_Chl_TOT_zint_100m, XS(X_CHL100));
      STC(diag_tot_CaCO3_form_zint, XS(X_CACO3ZINT));
      STC(diag_tot_bSi_form, XS(X_BSI));
      STC(diag_O2_ZMIN, XS(X_O2MIN));
      STC(diag_O2_ZMIN_DEPTH, XS(X_O2MINDEPTH));
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        STCA(diag_photoC_zint, a, XS(X_ZPHOTO + a));
        STCA(diag_photoC_NO3_zint, a, XS(X_ZNO3 + a));
        STCA(diag_CaCO3_form_zint, a, XS(X_ZCACO3 + a));
      }
    } else {
      ECO_DIAG_C1_LIST(ZERO_C1)
      BGC_DIAG_CA_LIST(ZERO_CA)
    }
  }
  // ---- inventory partials of this block: [kEcoInvGroups][kInvGroup] =
  //      30 tracer slots, active cells, active columns, the eight Jint_* column sums
  if (inv) {
    __shared__ double s_red[BLOCK / 32];
    __shared__ double s_slot[BLOCK / 32][32];
    double *out = A.inv_partials + (size_t)blockIdx.x * (kEcoInvGroups * kInvGroup);
    s_slot[tid >> 5][tid & 31] = inv_acc;
    __syncthreads();
    if (tid < BGC_TRACER_CNT) {   // lane = canonical row; the partials are indexed by tracer slot
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < BLOCK / 32; ++w) t += s_slot[w][tid];
      out[A.slot_of_row[tid] - 1] = t;
    }
    {
      const double cells = block_sum((double)kmax, s_red), cols = block_sum(kmax > 0 ? 1.0 : 0.0, s_red);
      if (tid == 0) { out[30] = cells; out[31] = cols; }
    }
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {   // zero without diagnostics (the rows are never added to)
      const double t = block_sum(XS(X_JC + q), s_red);
      if (tid == 0) out[32 + q] = t;
    }
  }
#undef XS
#undef IN
#undef kmax
}

template <int DIAG, int BLOCK, int MINB>
cudaError_t launch_variant(const EcoArgs &a, cudaStream_t s) {
  const size_t smem = ((size_t)2 * R_ROWS * (BLOCK + 2) + (size_t)X_ROWS * BLOCK) * sizeof(double) + 2 * sizeof(unsigned long long);
  auto kern = eco_columns_kernel<DIAG, BLOCK, MINB>;
  // > 48 KB of dynamic shared memory is opt-in, per device: cheap enough to set on every launch
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = (a.nC + BLOCK - 1) / BLOCK;
  kern<<<grid, BLOCK, smem, s>>>(a);
  return cudaGetLastError();
}

// cp.async.bulk needs 16-byte aligned global addresses and sizes: every run starts at
// base + 8*(k*nC + BLOCK*j).  With every base pointer 16-byte aligned a run starts on a boundary
// or 8 bytes behind one, and the kernel fetches the aligned superset (fetch_level).
// Otherwise the stage is filled with 8-byte cp.async copies (same kernel, same arithmetic).
// (With nC AND nL odd the 30 tracer slabs, nL*nC elements apart, alternate between the two alignments, so the
// rows of one level would need different shifts: that case takes the cp.async staging as well.)
bool slabs_are_bulk_copyable(const EcoArgs &a, bool diag) {
  if ((a.nC & 1) && (a.nL & 1)) return false;
  const void *p[] = {a.tracers, a.T, a.zmid, a.dz, a.zbot, a.fesedflux, diag ? a.S : a.T};
  for (const void *q : p) if (((size_t)q) & 15u) return false;
  return true;
}

// 0 (default): one 8-warp block per SM; 1: two 4-warp blocks per SM; 9: force the cp.async
// staging (also taken whenever the slabs are not bulk-copyable)
int block_of(int variant) { return variant == 1 ? 128 : 256; }

template <int DIAG>
cudaError_t launch_diag(const EcoArgs &a0, int variant, cudaStream_t s) {
  EcoArgs a = a0;
  a.bulk = (slabs_are_bulk_copyable(a, DIAG != 0) && variant != 9) ? 1 : 0;
  for (int r = 0; r < BGC_TRACER_CNT; ++r) {
    if (a.slot_of_row[r] < 1 || a.slot_of_row[r] > BGC_TRACER_CNT) return cudaErrorInvalidValue;
    a.tend_off[r] = (unsigned)(a.slot_of_row[r] - 1) * (unsigned)a.nL * (unsigned)a.nC;
  }
  if (block_of(variant) == 256) return launch_variant<DIAG, 256, 1>(a, s);
  return launch_variant<DIAG, 128, 2>(a, s);
}

}  // namespace

bool eco_rows_from_tables(const BgcTables &t, EcoArgs &a) {
  const BgcIndices &I = t.ind;
  const int plain[16] = {I.po4_ind, I.no3_ind, I.sio3_ind, I.nh4_ind, I.fe_ind, I.o2_ind, I.dic_ind, I.dic_alt_co2_ind,
                         I.alk_ind, I.doc_ind, I.don_ind, I.dofe_ind, I.dop_ind, I.dopr_ind, I.donr_ind, I.zooC_ind};
  bool seen[BGC_TRACER_CNT + 1] = {false};
  auto put = [&](int row, int slot) {
    a.slot_of_row[row] = slot;
    if (slot >= 1 && slot <= BGC_TRACER_CNT) seen[slot] = true;
  };
  for (int r = 0; r < 16; ++r) put(r, plain[r]);
  int si = 0, ca = 0;
  for (int g = 0; g < BGC_AUTOTROPH_CNT; ++g) {
    put(G_C(g), t.a[g].C_ind); put(G_CHL(g), t.a[g].Chl_ind); put(G_FE(g), t.a[g].Fe_ind);
    if (t.a[g].Si_ind > 0) si = t.a[g].Si_ind;
    if (t.a[g].CaCO3_ind > 0) ca = t.a[g].CaCO3_ind;
  }
  // a table without a silicifier / calcifier leaves that tracer unused by every group: its row takes
  // the slot the index table names (it is still a tracer of the array: its tendency is written as zero)
  put(SI_ROW, si > 0 ? si : I.diatSi_ind);
  put(CA_ROW, ca > 0 ? ca : I.spCaCO3_ind);
  for (int s = 1; s <= BGC_TRACER_CNT; ++s) if (!seen[s]) return false;
  return true;
}

int eco_sweep_blocks(int nC, int variant) {
  const int block = block_of(variant);
  return (nC + block - 1) / block;
}

int eco_inventory_parts(const EcoArgs &a, int diag_mode, int variant) {
  (void)diag_mode;
  const int block = block_of(variant);
  return (a.nC + block - 1) / block;
}

// diag_mode: 0 none, 1 some (NULL-checked stores), 2 every array present
cudaError_t launch_eco_columns(const EcoArgs &a, int diag_mode, int variant, cudaStream_t s) {
  if (a.nC <= 0 || a.nL <= 0) return cudaSuccess;
  if (diag_mode == 0) return launch_diag<0>(a, variant, s);
  if (diag_mode == 2) return launch_diag<2>(a, variant, s);
  return launch_diag<1>(a, variant, s);
}

}  // namespace bgc
