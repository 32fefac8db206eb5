// k_misc.cu — the MACROS source-sink kernel, the Fortran<->SoA and MPAS layout kernels, diagnostics
// accumulation and the deterministic inventory reductions.  (DMS: k_dms.cu.)
//
//   macros_cells_kernel   <- MACROS_SourceSink  MACROS_mod.F90:137-411 (thread = cell; no vertical coupling)
//
// An HBM-bound streaming kernel: each input element is read once and each output element written
// once, coalesced.
#include <stdlib.h>
#include "bgc_kernels.cuh"
#include "bgc_math.cuh"
#include "bgc_reduce.cuh"

namespace bgc {

__constant__ MacrosTables c_macros;

cudaError_t upload_macros_tables(const MacrosTables &t, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_macros, &t, sizeof(MacrosTables), 0, cudaMemcpyHostToDevice, s);
}

namespace {

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
int macros_inventory_parts(int nL, int nC) { return (int)(((size_t)nL * (size_t)nC + kMacrosBlock - 1) / kMacrosBlock); }

}  // namespace bgc
