// Stand-in for <cuda_runtime.h> so that ocean-bgc_b200/csrc/bgc_math.cuh compiles for the HOST
// (tests/test_device_math_on_host.py).  TEST INFRASTRUCTURE ONLY: it lets the accuracy claims of the
// device math be measured without a GPU; nothing in the product path uses it.
#pragma once
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __constant__ const
static inline int __double2loint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  long long b = ((long long)hi << 32) | (unsigned int)lo; double x; std::memcpy(&x, &b, 8); return x; }
// MUFU.RCP64H: a reciprocal good to 2^-20 (measured on sm_100a, scripts/micro/fp64_lat.cu).  The host twin
// keeps the 20 leading mantissa bits of the exact quotient and drops the rest - the worst seed the
// measured bound allows.
static inline double host_rcp_seed(double b) {
  double r = 1.0 / b; long long bits; std::memcpy(&bits, &r, 8); bits &= ~0xffffffffLL; std::memcpy(&r, &bits, 8); return r; }
struct HostTid { int x; };
static HostTid threadIdx = {0};
static inline void __syncthreads() {}
// warp votes of a one-lane "warp"
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline bool __all_sync(unsigned, bool p) { return p; }
