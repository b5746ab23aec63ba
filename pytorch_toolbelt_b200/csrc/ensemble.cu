// SURVEY 8f rank 2, second half: the model-ensemble average of inference/ensembling.py:71-123 (torch.stack of M model outputs
// + _deaugment_averaging) and the activation wrappers (:12-68) as single passes for sm_100a.
//
//   stack_reduce_kernel   out[i] = reduce_m f(x_m[i]) over M separately allocated tensors (pointer list by value in the kernel
//                         parameters): every input element is read once, the result written once -- no [M, ...] stack is ever
//                         materialised (the reference writes it, then reduces it: 3 passes instead of 1). Same seven reductions
//                         and the same sequential m = 0..M-1 summation order as the TTA view reduction (tta.py:63-96).
//   activation_kernel     in-place x.mul(T).sigmoid_() (ApplySigmoidTo) / x.mul(T).softmax(dim=1) (ApplySoftmaxTo) on [B,C,inner].
#include <algorithm>

#include "common.cuh"

namespace tb {

constexpr int kET = 256;
constexpr int kMaxEnsemble = 64;

struct StackPtrs {
  const float* p[kMaxEnsemble];
};

template <int VW>
__global__ void __launch_bounds__(kET) stack_reduce_kernel(const __grid_constant__ StackPtrs in, int M, long long n, int red,
                                                           float* __restrict__ out) {
  const long long nv = n / VW;
  for (long long v = (long long)blockIdx.x * kET + threadIdx.x; v < nv; v += (long long)gridDim.x * kET) {
    float acc[VW];
    for (int m = 0; m < M; ++m) {
      float x[VW];
      if (VW == 4) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(in.p[m]) + v);
        x[0] = q.x; x[1 % VW] = q.y; x[2 % VW] = q.z; x[3 % VW] = q.w;
      } else {
        x[0] = __ldcs(in.p[m] + v);
      }
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const float pv = red_pre<float>(red, x[e]);
        acc[e] = m == 0 ? pv : add_rn(acc[e], pv);
      }
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[e] = red_post<float>(red, acc[e], M);
    if (VW == 4) __stcs(reinterpret_cast<float4*>(out) + v, make_float4(acc[0], acc[1 % VW], acc[2 % VW], acc[3 % VW]));
    else out[v] = acc[0];
  }
}

// mode 0: sigmoid(x * T); mode 1: softmax(x * T) over the channel dim of [B,C,inner]; in place
__global__ void __launch_bounds__(kET) activation_kernel(float* __restrict__ x, long long B, int C, long long inner, int mode,
                                                         float T) {
  if (mode == 0) {
    const long long n = B * C * inner;
    for (long long i = (long long)blockIdx.x * kET + threadIdx.x; i < n; i += (long long)gridDim.x * kET)
      x[i] = 1.f / (1.f + expf(-(x[i] * T)));
    return;
  }
  const long long npix = B * inner;
  for (long long pix = (long long)blockIdx.x * kET + threadIdx.x; pix < npix; pix += (long long)gridDim.x * kET) {
    const long long b = pix / inner, pos = pix - b * inner;
    float* xp = x + b * C * inner + pos;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, xp[(long long)c * inner] * T);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(xp[(long long)c * inner] * T - mx);
    for (int c = 0; c < C; ++c) xp[(long long)c * inner] = expf(xp[(long long)c * inner] * T - mx) / den;
  }
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_stack_reduce(const float* const* inputs_host, int M, int64_t n, int reduction, float* out, void* stream) {
  TB_REQUIRE(inputs_host && out, TB_ERR_ARG, "tb_stack_reduce: null pointer");
  TB_REQUIRE(M > 0 && M <= kMaxEnsemble, TB_ERR_UNSUPPORTED, "tb_stack_reduce: 1 <= M <= %d inputs", kMaxEnsemble);
  TB_REQUIRE(n >= 0, TB_ERR_ARG, "tb_stack_reduce: bad length");
  TB_REQUIRE(reduction >= TB_RED_MEAN && reduction <= TB_RED_LOG1P, TB_ERR_ARG, "tb_stack_reduce: bad reduction %d", reduction);
  if (n == 0) return TB_OK;
  StackPtrs ptrs;
  bool vec4 = (n % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int m = 0; m < M; ++m) {
    TB_REQUIRE(inputs_host[m], TB_ERR_ARG, "tb_stack_reduce: input %d is null", m);
    ptrs.p[m] = inputs_host[m];
    vec4 = vec4 && (reinterpret_cast<uintptr_t>(inputs_host[m]) & 15) == 0;
  }
  const long long items = vec4 ? n / 4 : n;
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + kET - 1) / kET, (long long)sm_count() * 16));
  cudaStream_t s = (cudaStream_t)stream;
  if (vec4) stack_reduce_kernel<4><<<grid, kET, 0, s>>>(ptrs, M, n, reduction, out);
  else stack_reduce_kernel<1><<<grid, kET, 0, s>>>(ptrs, M, n, reduction, out);
  TB_LAUNCH_CHECK("stack_reduce_kernel");
  return TB_OK;
}

int tb_activation_inplace(float* x, int64_t B, int C, int64_t inner, int mode, float temperature, void* stream) {
  TB_REQUIRE(x && B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_activation_inplace: bad argument");
  TB_REQUIRE(mode == 0 || mode == 1, TB_ERR_ARG, "tb_activation_inplace: mode must be 0 (sigmoid) or 1 (softmax over dim 1)");
  const long long items = mode == 0 ? B * C * inner : B * inner;
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + kET - 1) / kET, (long long)sm_count() * 16));
  activation_kernel<<<grid, kET, 0, (cudaStream_t)stream>>>(x, B, C, inner, mode, temperature);
  TB_LAUNCH_CHECK("activation_kernel");
  return TB_OK;
}

}  // extern "C"
