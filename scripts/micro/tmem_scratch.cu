// Tensor memory (TMEM) as a per-thread scratchpad for a kernel that runs no MMA at all: can the column sweep keep its
// once-per-level column state and running sums there instead of in shared memory / registers?
//   1. correctness: every thread stores its own doubles with tcgen05.st.32x32b.x2 and reads them back with tcgen05.ld
//   2. latency of a dependent ld -> fma -> st -> ld chain through TMEM, against the same chain through shared memory
//   3. throughput of NACC accumulators per thread read-modify-written per "level" beside an FP64 chain, 8 and 12 warps per SM
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_scratch tmem_scratch.cu && ./tmem_scratch
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tm_st(unsigned taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ double tm_ld(unsigned taddr) {
  unsigned lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: correctness; 1: TMEM chain; 2: shared-memory chain; 3: TMEM accumulators; 4: shared-memory accumulators
template <int THREADS, int NACC>
__global__ void __launch_bounds__(THREADS, 1) k(int mode, int iters, double *out, unsigned long long *cyc, int *bad) {
  extern __shared__ double sm[];
  __shared__ unsigned s_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned base = s_base;
  // this warp's lanes: quadrant warp % 4; its columns: slot warp / 4 of 512 / (THREADS / 128) columns
  constexpr int SLOTS = THREADS / 128, COLS = 512 / SLOTS;
  const unsigned my = base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * COLS);
  double acc = 1.0 + tid * 1e-6;
  unsigned long long t0 = clock64();
  if (mode == 0) {
    for (int j = 0; j < COLS / 2; ++j) tm_st(my + 2 * j, tid * 1000.0 + j);
    tm_wait_st();
    for (int j = 0; j < COLS / 2; ++j) if (tm_ld(my + 2 * j) != tid * 1000.0 + j) atomicAdd(bad, 1);
  } else if (mode == 1) {
    tm_st(my, acc); tm_wait_st();
    for (int i = 0; i < iters; ++i) { double v = tm_ld(my); v = fma(v, 1.0000001, 1e-9); tm_st(my, v); tm_wait_st(); }
    acc = tm_ld(my);
  } else if (mode == 2) {
    volatile double *p = sm + tid;
    *p = acc;
    for (int i = 0; i < iters; ++i) { double v = *p; v = fma(v, 1.0000001, 1e-9); *p = v; }
    acc = *p;
  } else if (mode == 3) {
    for (int j = 0; j < NACC; ++j) tm_st(my + 2 * j, 0.0);
    tm_wait_st();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) {
        acc = fma(acc, 1.0000001, 1e-9); acc = fma(acc, 0.9999999, 1e-9);   // some FP64 work between the accesses
        const double v = tm_ld(my + 2 * j);
        tm_st(my + 2 * j, v + acc);
      }
      tm_wait_st();
    }
    for (int j = 0; j < NACC; ++j) acc += tm_ld(my + 2 * j);
  } else {
    for (int j = 0; j < NACC; ++j) sm[j * THREADS + tid] = 0.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) {
        acc = fma(acc, 1.0000001, 1e-9); acc = fma(acc, 0.9999999, 1e-9);
        sm[j * THREADS + tid] = sm[j * THREADS + tid] + acc;
      }
    }
    for (int j = 0; j < NACC; ++j) acc += sm[j * THREADS + tid];
  }
  unsigned long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * THREADS + tid] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
}

template <int THREADS, int NACC>
static void run(const char *what, int mode, int iters) {
  double *out; unsigned long long *cyc; int *bad;
  CK(cudaMalloc(&out, 148 * THREADS * 8)); CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
  auto kern = k<THREADS, NACC>;
  const int smem = 150 * 1024;   // one block per SM
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  kern<<<148, THREADS, smem>>>(mode, iters, out, cyc, bad); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); kern<<<148, THREADS, smem>>>(mode, iters, out, cyc, bad); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  unsigned long long hc; int hb; CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
  printf("%-62s threads %3d: %8.3f ms, %10llu cycles (%.1f per iteration%s), mismatches %d\n", what, THREADS, ms, hc,
         (double)hc / (iters > 0 ? iters : 1), mode >= 3 ? " of NACC accesses" : "", hb);
  CK(cudaFree(out)); CK(cudaFree(cyc)); CK(cudaFree(bad));
}

int main() {
  run<256, 32>("store / load back every column pair of the thread's slot", 0, 0);
  run<384, 32>("store / load back every column pair of the thread's slot", 0, 0);
  run<256, 32>("dependent ld-fma-st chain through TMEM", 1, 20000);
  run<256, 32>("dependent ld-fma-st chain through shared memory", 2, 20000);
  run<256, 40>("40 accumulators per thread in TMEM, 2 FMA between accesses", 3, 2000);
  run<256, 40>("40 accumulators per thread in shared memory", 4, 2000);
  run<384, 40>("40 accumulators per thread in TMEM, 2 FMA between accesses", 3, 2000);
  run<384, 40>("40 accumulators per thread in shared memory", 4, 2000);
  return 0;
}
