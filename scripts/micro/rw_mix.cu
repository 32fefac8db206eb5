// HBM bandwidth for the read:write mix and the access pattern of the column sweep
// (eco_columns_kernel): thread = column, levels in order, NR input arrays read and NW output
// arrays written per level, each array a (level, column) slab set with the column index fastest.
// The sweep moves 37 reads + 146 writes per cell (1464 B) and the whole BGC_SourceSink 37 + 162;
// MEASURED_PEAKS.json's 6548 GB/s is a 1:1 copy, so this microbenchmark answers: what does the
// memory system sustain for an 80 % write stream spread over ~180 concurrently open arrays, and
// how much of that needs more than the sweep's 8 warps per SM?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rw_mix rw_mix.cu && ./rw_mix
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", \
  cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Ptrs { const double *in; double *out; };

// flat grid-stride streams (the shape MEASURED_PEAKS was taken with)
__global__ void flat_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void flat_write(double2 *__restrict__ b, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    b[i] = make_double2(v, v);
}
__global__ void flat_read(const double2 *__restrict__ a, double *sink, size_t n) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double2 x = a[i]; s += x.x + x.y; }
  if (s == 1.2345e300) *sink = s;
}

// the sweep's pattern: one thread per column, NR + NW arrays touched at every level
template <int NR, int NW>
__global__ void column_streams(Ptrs p, int nL, int nC, double *sink) {
  extern __shared__ double pad[];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nC) return;
  const size_t slab = (size_t)nL * nC;
  double s = 0;
  for (int k = 0; k < nL; ++k) {
    const size_t e = (size_t)k * nC + col;
#pragma unroll
    for (int r = 0; r < NR; ++r) s += p.in[r * slab + e];
#pragma unroll
    for (int w = 0; w < NW; ++w) p.out[w * slab + e] = s + w;
  }
  if (s == 1.2345e300) *sink = s + pad[0];
}

// Same traffic, three layouts, occupancy forced by __launch_bounds__ (registers capped):
//   LAYOUT 0: the C ABI's SoA   [array][level][column]              (2 KB runs, 183 streams)
//   LAYOUT 1: block-tiled SoA   [array][block][level][256 columns]  (consecutive levels adjacent)
//   LAYOUT 2: AoSoA             [block][level][array][256 columns]  (one flat stream per block)
template <int NR, int NW, int MINB, int LAYOUT>
__global__ void __launch_bounds__(256, MINB) column_layouts(Ptrs p, int nL, int nC, double *sink) {
  const int tid = threadIdx.x, blk = blockIdx.x;
  const int col = blk * 256 + tid;
  if (col >= nC) return;
  const size_t slab = (size_t)nL * nC;
  const size_t nB = (nC + 255) / 256;
  double s = 0;
  for (int k = 0; k < nL; ++k) {
    size_t e, stride;
    if (LAYOUT == 0) { e = (size_t)k * nC + col; stride = slab; }
    else if (LAYOUT == 1) { e = ((size_t)blk * nL + k) * 256 + tid; stride = nB * nL * 256; }
    else { e = 0; stride = 256; }
    const size_t rbase = LAYOUT == 2 ? (((size_t)blk * nL + k) * NR) * 256 + tid : e;
    const size_t wbase = LAYOUT == 2 ? (((size_t)blk * nL + k) * NW) * 256 + tid : e;
#pragma unroll 8
    for (int r = 0; r < NR; ++r) s += p.in[rbase + r * stride];
#pragma unroll 8
    for (int w = 0; w < NW; ++w) p.out[wbase + w * stride] = s + w;
  }
  if (s == 1.2345e300) *sink = s;
}

// ---- how the stores are issued (write-only, the sweep's 146 output streams, SoA layout) ----
//   MODE 0: one 8-byte st.global per thread and array (what the sweep does)
//   MODE 1: the same with the streaming hint (st.global.cs)
//   MODE 2: a thread owns two adjacent columns: 16-byte stores, half the store instructions
//   MODE 3: values staged in shared memory, one elected thread hands 2-KB runs to the TMA unit
//           (cp.async.bulk shared -> global), double-buffered in groups of 8 arrays
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int NW, int MODE>
__global__ void __launch_bounds__(256, 1) store_paths(double *out, int nL, int nC, double seed) {
  extern __shared__ __align__(128) double stage[];       // MODE 3: 2 x 8 x 256 doubles
  const int tid = threadIdx.x;
  const size_t slab = (size_t)nL * nC;
  if (MODE == 2) {
    const int col = (blockIdx.x * 256 + tid) * 2;
    if (col >= nC) return;
    for (int k = 0; k < nL; ++k) {
      const size_t e = (size_t)k * nC + col;
#pragma unroll 8
      for (int w = 0; w < NW; ++w) *reinterpret_cast<double2 *>(out + w * slab + e) = make_double2(seed + w, seed - w);
    }
    return;
  }
  const int col = blockIdx.x * 256 + tid;
  if (MODE != 3) {
    if (col >= nC) return;
    for (int k = 0; k < nL; ++k) {
      const size_t e = (size_t)k * nC + col;
#pragma unroll 8
      for (int w = 0; w < NW; ++w) {
        if (MODE == 1) __stcs(out + w * slab + e, seed + w);
        else out[w * slab + e] = seed + w;
      }
    }
    return;
  }
  // MODE 3 (whole blocks only: nC is a multiple of 256)
  constexpr int G = 8;
  int phase = 0;
  for (int k = 0; k < nL; ++k) {
    const size_t e0 = (size_t)k * nC + (size_t)blockIdx.x * 256;
    for (int w0 = 0; w0 < NW; w0 += G, phase ^= 1) {
      double *st = stage + phase * G * 256;
      // the bulk copies that read this half two rounds ago must have finished reading it
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
#pragma unroll
      for (int g = 0; g < G; ++g) if (w0 + g < NW) st[g * 256 + tid] = seed + (w0 + g);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (w0 + g < NW)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::
                         "l"(out + (size_t)(w0 + g) * slab + e0), "r"(smem_addr(st + g * 256)), "n"(2048) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// ---- the sweep's own structure without its arithmetic: inputs staged by the TMA unit ----
// One 256-thread block per SM; lane 0 of every warp fetches its share of the NEXT level's NR input
// rows with cp.async.bulk (2-KB runs) into a double-buffered shared-memory stage signalled on an
// mbarrier, exactly as eco_columns_kernel does; the compute threads read shared memory only and issue
// the NW plain 8-byte stores of the level.  What this kernel sustains is what the memory system gives
// the sweep's traffic at the sweep's occupancy when NO load latency is exposed; the gap between it and
// the sweep is the sweep's arithmetic (dependency latency at 8 warps per SM), not HBM.
template <int NR, int NW, int SPREAD>
__global__ void __launch_bounds__(256, 1) column_tma(Ptrs p, int nL, int nC, double *sink) {
  extern __shared__ __align__(128) double stage[];           // 2 x NR x 256 doubles, then 2 mbarriers
  unsigned long long *bars = (unsigned long long *)(stage + 2 * NR * 256);
  const int tid = threadIdx.x;
  const size_t slab = (size_t)nL * nC;
  const size_t col0 = (size_t)blockIdx.x * 256;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bars[b])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto fetch = [&](int k) {
    if (tid & 31) return;
    const unsigned bar = smem_addr(&bars[k & 1]);
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NR * 2048) : "memory");
    for (int r = tid >> 5; r < NR; r += 8)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(stage + ((k & 1) * NR + r) * 256)), "l"(p.in + r * slab + (size_t)k * nC + col0),
                     "r"(2048), "r"(bar) : "memory");
  };
  fetch(0);
  double s = 0;
  for (int k = 0; k < nL; ++k) {
    const unsigned bar = smem_addr(&bars[k & 1]), parity = (unsigned)((k >> 1) & 1);
    asm volatile("{\n\t.reg .pred P1;\n\tLW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra LW;\n\tDN:\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
    if (k + 1 < nL) fetch(k + 1);
    const double *st = stage + (k & 1) * NR * 256;
#pragma unroll
    for (int r = 0; r < NR; ++r) s += st[r * 256 + tid];
    const size_t e = (size_t)k * nC + col0 + tid;
    // SPREAD = 0: the stores of a level back to back; 1: a dependent FP64 chain between them (8 x 8
    // cycles each), so that they leave at the sweep's pace instead of in one burst
#pragma unroll 8
    for (int w = 0; w < NW; ++w) {
      if (SPREAD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fma(s, 1.0000001, 1e-9);
      }
      p.out[w * slab + e] = s + w;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (s == 1.2345e300) *sink = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

template <int NR, int NW>
static void run_columns(const double *in, double *out, double *sink, int nL, int nC, int block, int smem_kb,
                        const char *what) {
  auto kern = column_streams<NR, NW>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, block, smem_kb * 1024));
  Ptrs p{in, out};
  const int grid = (nC + block - 1) / block;
  float ms = time_ms([&] { kern<<<grid, block, smem_kb * 1024>>>(p, nL, nC, sink); });
  const double bytes = (double)(NR + NW) * 8.0 * nL * nC;
  printf("columns %3dR+%3dW block %4d, %2d blocks/SM (%2d warps/SM): %7.3f ms  %7.1f GB/s   %s\n", NR, NW, block, perSM,
         perSM * block / 32, ms, bytes / ms * 1e-6, what);
}

template <int NR, int NW, int MINB, int LAYOUT>
static void run_layout(const double *in, double *out, double *sink, int nL, int nC, const char *what) {
  auto kern = column_layouts<NR, NW, MINB, LAYOUT>;
  // occupancy is pinned by a dynamic shared-memory reservation (the kernels need few registers)
  const int smem = (MINB == 1 ? 200 : MINB == 2 ? 100 : MINB == 4 ? 50 : 0) * 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, 256, smem));
  Ptrs p{in, out};
  const int grid = (nC + 255) / 256;
  float ms = time_ms([&] { kern<<<grid, 256, smem>>>(p, nL, nC, sink); });
  const double bytes = (double)(NR + NW) * 8.0 * nL * nC;
  printf("layout %d %3dR+%3dW, %2d blocks/SM (%2d warps/SM): %7.3f ms  %7.1f GB/s   %s\n", LAYOUT, NR, NW, perSM,
         perSM * 8, ms, bytes / ms * 1e-6, what);
}

template <int NW, int MODE>
static void run_stores(double *out, int nL, int nC, const char *what) {
  auto kern = store_paths<NW, MODE>;
  const int smem = 200 * 1024;               // one block per SM, as the sweep
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int cols_per_block = MODE == 2 ? 512 : 256;
  const int grid = (nC + cols_per_block - 1) / cols_per_block;
  float ms = time_ms([&] { kern<<<grid, 256, smem>>>(out, nL, nC, 1.0); });
  printf("stores mode %d, %3d write streams, 1 block/SM: %7.3f ms  %7.1f GB/s   %s\n", MODE, NW, ms,
         (double)NW * 8.0 * nL * nC / ms * 1e-6, what);
}

template <int NR, int NW, int SPREAD>
static void run_tma(const double *in, double *out, double *sink, int nL, int nC, const char *what) {
  auto kern = column_tma<NR, NW, SPREAD>;
  const int smem = 2 * NR * 2048 + 16;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Ptrs p{in, out};
  float ms = time_ms([&] { kern<<<nC / 256, 256, smem>>>(p, nL, nC, sink); });
  printf("TMA-fed  %3dR+%3dW,  1 block/SM ( 8 warps/SM): %7.3f ms  %7.1f GB/s   %s\n", NR, NW, ms,
         (double)(NR + NW) * 8.0 * nL * nC / ms * 1e-6, what);
}

int main() {
  const int nL = 60, nC = 235160;              // EC60to30
  const size_t cells = (size_t)nL * nC;
  const int NRMAX = 37, NWMAX = 162;
  double *in, *out, *sink;
  CK(cudaMalloc(&in, NRMAX * cells * 8)); CK(cudaMalloc(&out, NWMAX * cells * 8)); CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(in, 0, NRMAX * cells * 8));
  const size_t n2 = (size_t)16 * cells / 2;    // 16 slabs = 1.8 GB per stream (>> 126 MB L2)
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float ms;
  ms = time_ms([&] { flat_copy<<<sms * 16, 512>>>((const double2 *)in, (double2 *)out, n2); });
  printf("flat copy 1R+1W : %7.3f ms  %7.1f GB/s\n", ms, 2.0 * n2 * 16 / ms * 1e-6);
  ms = time_ms([&] { flat_write<<<sms * 16, 512>>>((double2 *)out, n2, 1.0); });
  printf("flat write      : %7.3f ms  %7.1f GB/s\n", ms, 1.0 * n2 * 16 / ms * 1e-6);
  ms = time_ms([&] { flat_read<<<sms * 16, 512>>>((const double2 *)in, sink, n2); });
  printf("flat read       : %7.3f ms  %7.1f GB/s\n", ms, 1.0 * n2 * 16 / ms * 1e-6);
  ms = time_ms([&] { CK(cudaMemsetAsync(out, 0, n2 * 16)); });
  printf("cudaMemset      : %7.3f ms  %7.1f GB/s\n", ms, 1.0 * n2 * 16 / ms * 1e-6);

  const int nCt = 235008;                    // whole 256-column blocks for the tiled / bulk variants
  if (getenv("RW_MIX_TMA")) {   // the sweep's structure (TMA-staged inputs, plain stores) without arithmetic
    run_tma<37, 146, 0>(in, out, sink, nL, nCt, "sweep mix, inputs by cp.async.bulk, stores in a burst");
    run_tma<37, 146, 1>(in, out, sink, nL, nCt, "sweep mix, inputs by cp.async.bulk, stores paced by an FP64 chain");
    run_tma<33, 146, 0>(in, out, sink, nL, nCt, "33 fetched rows (as the sweep), stores in a burst");
    run_layout<37, 146, 1, 0>(in, out, sink, nL, nCt, "per-thread loads, 1 block/SM (the r01 proxy)");
    return 0;
  }
  if (getenv("RW_MIX_STORES") && !getenv("RW_MIX_ALL")) {
    run_stores<146, 0>(out, nL, nCt, "st.global, 8 B per thread");
    run_stores<146, 1>(out, nL, nCt, "st.global.cs");
    run_stores<146, 2>(out, nL, nCt, "16-byte stores, two columns per thread");
    run_stores<146, 3>(out, nL, nCt, "cp.async.bulk shared -> global (TMA), 2-KB runs");
    return 0;
  }
  // the sweep's mix at the sweep's occupancy (one 256-thread block per SM) and above it
  run_columns<37, 146>(in, out, sink, nL, nC, 256, 200, "sweep mix, sweep occupancy");
  run_columns<37, 146>(in, out, sink, nL, nC, 256, 100, "sweep mix, 2 blocks/SM");
  run_columns<37, 146>(in, out, sink, nL, nC, 256, 48, "sweep mix, 4 blocks/SM");
  run_columns<37, 146>(in, out, sink, nL, nC, 256, 0, "sweep mix, full occupancy");
  run_columns<37, 146>(in, out, sink, nL, nC, 128, 0, "sweep mix, 128-thread blocks");
  run_columns<37, 162>(in, out, sink, nL, nC, 256, 0, "BGC_SourceSink mix");
  run_columns<13, 41>(in, out, sink, nL, nC, 256, 0, "DMS mix");
  run_columns<8, 14>(in, out, sink, nL, nC, 256, 0, "MACROS mix");
  run_columns<1, 1>(in, out, sink, nL, nC, 256, 0, "copy, column pattern");
  run_columns<0, 16>(in, out, sink, nL, nC, 256, 0, "16 write streams");
  run_columns<0, 146>(in, out, sink, nL, nC, 256, 0, "146 write streams");
  run_columns<37, 0>(in, out, sink, nL, nC, 256, 0, "37 read streams");

  // occupancy and layout: tiles need whole blocks (nCt is a multiple of 256)
  run_layout<37, 146, 1, 0>(in, out, sink, nL, nCt, "SoA, 1 block/SM");
  run_layout<37, 146, 2, 0>(in, out, sink, nL, nCt, "SoA, 2 blocks/SM");
  run_layout<37, 146, 4, 0>(in, out, sink, nL, nCt, "SoA, 4 blocks/SM");
  run_layout<37, 146, 8, 0>(in, out, sink, nL, nCt, "SoA, 8 blocks/SM");
  run_layout<37, 146, 1, 1>(in, out, sink, nL, nCt, "block-tiled SoA, 1 block/SM");
  run_layout<37, 146, 4, 1>(in, out, sink, nL, nCt, "block-tiled SoA, 4 blocks/SM");
  run_layout<37, 146, 1, 2>(in, out, sink, nL, nCt, "AoSoA, 1 block/SM");
  run_layout<37, 146, 2, 2>(in, out, sink, nL, nCt, "AoSoA, 2 blocks/SM");
  run_layout<37, 146, 4, 2>(in, out, sink, nL, nCt, "AoSoA, 4 blocks/SM");
  run_layout<37, 146, 8, 2>(in, out, sink, nL, nCt, "AoSoA, 8 blocks/SM");

  if (getenv("RW_MIX_STORES")) {
    run_stores<146, 0>(out, nL, nCt, "st.global, 8 B per thread");
    run_stores<146, 1>(out, nL, nCt, "st.global.cs");
    run_stores<146, 2>(out, nL, nCt, "16-byte stores, two columns per thread");
    run_stores<146, 3>(out, nL, nCt, "cp.async.bulk shared -> global (TMA), 2-KB runs");
  }
  return 0;
}
