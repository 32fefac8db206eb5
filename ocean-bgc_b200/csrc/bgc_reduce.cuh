// bgc_reduce.cuh — deterministic reductions used by the fused inventory sums.
// No atomics anywhere: every sum is formed in a fixed order, so the inventory vector is
// bit-reproducible from run to run and independent of scheduling.
#pragma once
#include <cuda_runtime.h>

namespace bgc {

// Block-wide sum: fixed shuffle tree inside each warp, then the warp sums in index
// order.  `smem` holds blockDim.x / 32 doubles.  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += smem[w];
  }
  return r;
}

}  // namespace bgc
