// Helpers shared by the loss translation units (losses.cu, losses_extra.cu): launch geometry, deterministic block /
// partial reductions, 128-bit streaming loads / stores, the one-exp sigmoid + softplus pair.
#pragma once

#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace tb {

constexpr int kLT = 256;  // threads per CTA for loss kernels

static inline int loss_grid(long long work_items) {
  const long long want = (work_items + kLT - 1) / kLT;
  const long long cap = (long long)sm_count() * 8;
  return (int)std::max<long long>(1, std::min<long long>(want, cap));
}

// ---- block reduction of N doubles; result valid in thread 0 -------------------------------------
template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N]) {
  __shared__ double sh[N][kLT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double x = v[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[n][warp] = x;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double x = lane < kLT / 32 ? sh[n][lane] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[n] = x;
    }
  }
  __syncthreads();
}

// sum `count` partial rows of N doubles (row stride N) in a fixed order: deterministic
template <int N>
__device__ __forceinline__ void reduce_partials(const double* __restrict__ part, int count, double (&out)[N]) {
  double v[N];
#pragma unroll
  for (int n = 0; n < N; ++n) v[n] = 0.0;
  for (int i = threadIdx.x; i < count; i += kLT) {
#pragma unroll
    for (int n = 0; n < N; ++n) v[n] += part[(size_t)i * N + n];
  }
  block_reduce<N>(v);
#pragma unroll
  for (int n = 0; n < N; ++n) out[n] = v[n];
}

template <int VW>
__device__ __forceinline__ void ld_vec(const float* p, long long v, float (&o)[VW]) {
  if (VW == 4) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p) + v);
    o[0] = q.x; o[1 % VW] = q.y; o[2 % VW] = q.z; o[3 % VW] = q.w;
  } else {
    o[0] = __ldcs(p + v);
  }
}
template <int VW>
__device__ __forceinline__ void st_vec(float* p, long long v, const float (&o)[VW]) {
  if (VW == 4)
    __stcs(reinterpret_cast<float4*>(p) + v, make_float4(o[0], o[1 % VW], o[2 % VW], o[3 % VW]));
  else
    __stcs(p + v, o[0]);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// binary_cross_entropy_with_logits(x, t) = max(x,0) - x t + log1p(exp(-|x|))
// log1p(e) for e = exp(-|x|) in (0,1]: one MUFU.LG2; absolute error ~4e-7, i.e. <= 1e-6 relative on any loss sum
__device__ __forceinline__ float softplus_from_exp(float e) { return __logf(1.f + e); }
__device__ __forceinline__ float bce_logits(float x, float t) {
  return fmaxf(x, 0.f) - x * t + softplus_from_exp(__expf(-fabsf(x)));
}
__device__ __forceinline__ float powg(float base, float gamma) {
  if (gamma == 2.f) return base * base;
  if (gamma == 1.f) return base;
  if (gamma == 0.f) return 1.f;
  return powf(base, gamma);
}
// d/dbase base^gamma with torch's convention (0 for gamma == 0)
__device__ __forceinline__ float dpowg(float base, float gamma) {
  if (gamma == 2.f) return 2.f * base;
  if (gamma == 1.f) return 1.f;
  if (gamma == 0.f) return 0.f;
  return gamma * powf(base, gamma - 1.f);
}


}  // namespace tb
