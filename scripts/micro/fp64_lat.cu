// Microbenchmarks that inform bgc_math.cuh: dependent-issue latency of DFMA / DADD / DMUL /
// MUFU.RCP64H / shared-memory load on sm_100a, and the accuracy of rcp.approx.ftz.f64.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void lat_dfma(double *out, long long *cyc, double a, double b, int n) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 64; ++j) x = fma(x, b, a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_dadd(double *out, long long *cyc, double a, int n) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 64; ++j) x = x + a;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rcp(double *out, long long *cyc, double a, int n) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_lds(double *out, long long *cyc, int n) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) p = s[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: several warps, independent DFMAs
__global__ void thr_dfma(double *out, long long *cyc, double a, double b, int n) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
      x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a); }
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rcp_acc(const double *in, double *relerr, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double b = in[i], r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  relerr[i] = fabs(fma(-b, r, 1.0));
}
int main() {
  double *out; long long *cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  int n = 1000;
  lat_dfma<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n); cudaDeviceSynchronize(); printf("DFMA dependent latency: %.2f cycles\n", (double)cyc[0] / (n * 64));
  lat_dadd<<<1, 32>>>(out, cyc, 1e-9, n); cudaDeviceSynchronize(); printf("DADD dependent latency: %.2f cycles\n", (double)cyc[0] / (n * 64));
  lat_rcp<<<1, 32>>>(out, cyc, 1.7, n); cudaDeviceSynchronize(); printf("MUFU.RCP64H(+mov) dependent latency: %.2f cycles\n", (double)cyc[0] / (n * 16));
  lat_lds<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize(); printf("LDS dependent latency: %.2f cycles\n", (double)cyc[0] / (n * 32));
  for (int warps = 1; warps <= 16; warps *= 2) {
    thr_dfma<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, n); cudaDeviceSynchronize();
    printf("DFMA throughput, %2d warps x 8 independent chains on one SM: %.3f warp-DFMA/cycle/SM\n", warps, (double)n * 64 * warps / cyc[0]);
  }
  const int N = 1 << 20; double *in, *err; cudaMallocManaged(&in, N * 8); cudaMallocManaged(&err, N * 8);
  for (int i = 0; i < N; ++i) in[i] = ldexp(1.0 + (double)i / N, (i % 41) - 20) * ((i & 1) ? 1 : -1);
  rcp_acc<<<N / 256, 256>>>(in, err, N); cudaDeviceSynchronize();
  double m = 0; for (int i = 0; i < N; ++i) if (err[i] > m) m = err[i];
  printf("rcp.approx.ftz.f64 max relative error: %.3e (2^%.1f)\n", m, log2(m));
  return 0;
}
