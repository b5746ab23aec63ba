// the remaining element-wise losses of the reference on the fused forward+gradient skeleton (sm_100a).
//
// Reference semantics restated (file:line relative to /root/reference/pytorch_toolbelt/losses):
//   soft_bce.py:9-50             SoftBCEWithLogitsLoss (label smoothing, ignore_index, weight, pos_weight)  -> elem kind 0
//   balanced_bce.py:10-49        balanced_binary_cross_entropy_with_logits (class-count balanced weights)   -> count + kind 1
//   quality_focal_loss.py:5-53   QualityFocalLoss (|sigmoid - t|^beta * BCE, mean / sum / normalized)       -> elem kind 2
//   soft_f1.py:8-107             soft_micro_f1, BinarySoftF1Loss, SoftF1Loss                                -> f1_* kernels
//
// One streaming pass over (x, t) produces the loss partial sums and d loss / d x (two passes where the loss is a ratio of
// two global sums): 128-bit loads / stores, warp-shuffle reductions into per-CTA double partials, a one-CTA deterministic
// finalisation. Same structure as focal_kernel in losses.cu.
#include "loss_common.cuh"

namespace tb {

enum : int { EL_SOFT_BCE = 0, EL_BALANCED_BCE = 1, EL_QFL = 2 };

struct ElemArgs {
  const float* x;
  const float* t;
  long long B, inner;
  int C;
  tb_elem_loss_params p;
  const float* weight;      // soft bce: [1] or [C] or null
  const float* pos_weight;  // soft bce: [1] or [C] or null
  const double* counts;     // balanced bce: [2] = (#t==1, #t==0), device
  float* loss_elem;         // [B,C,inner] (reduction none) or null
  float* grad;              // [B,C,inner] or null
  double* partials;         // [grid,2] (sum L, sum F)
  const double* sums;       // pass 2 of the normalised QFL
  float gscale;
};

// stable log(sigmoid(x)) pieces from one exp: sp_neg = softplus(-x) = -log sigmoid(x); sig = sigmoid(x)
__device__ __forceinline__ void sigmoid_pieces(float x, float& sig, float& sp_neg) {
  const float e = __expf(-fabsf(x));
  const float inv = 1.f / (1.f + e);
  sig = x >= 0.f ? inv : e * inv;
  sp_neg = fmaxf(-x, 0.f) + __logf(1.f + e);
}

template <int KIND>
__device__ __forceinline__ void elem_loss(const ElemArgs& a, float x, float t, int c, float bal_a, float bal_b, float& L, float& F,
                                          float& dL, float& dF) {
  float sig, sp_neg;
  sigmoid_pieces(x, sig, sp_neg);
  F = 0.f;
  dF = 0.f;
  if (KIND == EL_SOFT_BCE) {
    // soft_bce.py:31-43: soft target, F.binary_cross_entropy_with_logits(weight, pos_weight), then *= (target != ignore)
    const float st = a.p.has_smooth ? (1.f - t) * a.p.smooth + t * (1.f - a.p.smooth) : t;
    const float pw = a.pos_weight ? __ldg(a.pos_weight + (a.p.pos_weight_n > 1 ? c : 0)) : 1.f;
    const float w = a.weight ? __ldg(a.weight + (a.p.weight_n > 1 ? c : 0)) : 1.f;
    const float coef = pw * st + (1.f - st);  // 1 + (pw - 1) st
    L = w * ((1.f - st) * x + coef * sp_neg);
    dL = w * ((1.f - st) - coef * (1.f - sig));
    if (a.p.has_ignore && t == a.p.ignore_index) {
      L = 0.f;
      dL = 0.f;
    }
  } else if (KIND == EL_BALANCED_BCE) {
    // balanced_bce.py:34-42: -(pw^g t log sig(x) + nw^g (1-t) log sig(-x)); masked_fill(t == ignore, 0)
    const float sp_pos = sp_neg + x;  // softplus(x) = -log sigmoid(-x)
    L = bal_a * t * sp_neg + bal_b * (1.f - t) * sp_pos;
    dL = -bal_a * t * (1.f - sig) + bal_b * (1.f - t) * sig;
    if (a.p.has_ignore && t == a.p.ignore_index) {
      L = 0.f;
      dL = 0.f;
    }
  } else {
    // quality_focal_loss.py:37-39: bce * |sigmoid(x) - t|^beta
    const float bce = (1.f - t) * x + sp_neg;
    const float d = sig - t, ad = fabsf(d);
    F = powg(ad, a.p.beta);
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    dF = dpowg(ad, a.p.beta) * sgn * sig * (1.f - sig);
    L = F * bce;
    dL = dF * bce + F * (sig - t);
  }
}

// PASS 0: single pass; PASS 1: sums only (normalised QFL); PASS 2: element losses / gradient from the sums
template <int KIND, int PASS, int VW>
__global__ void __launch_bounds__(kLT) elem_loss_kernel(ElemArgs a) {
  const long long totalv = a.B * a.C * a.inner / VW, innerv = a.inner / VW;
  double acc[2] = {0.0, 0.0};
  float bal_a = 0.f, bal_b = 0.f;
  if (KIND == EL_BALANCED_BCE) {
    // balanced_bce.py:26-32 in float32 like the reference: pos_weight = (neg / (num + 1e-7))^g, then ^g AGAIN on use
    const float pos = (float)a.counts[0], neg = (float)a.counts[1];
    const float pw = powf(neg / (pos + neg + 1e-7f), a.p.gamma);
    bal_a = powf(pw, a.p.gamma);
    bal_b = powf(1.f - pw, a.p.gamma);
  }
  float invN = 1.f, coefF = 0.f;
  if (PASS == 2) {
    invN = (float)(1.0 / a.sums[1]);
    coefF = (float)(a.sums[0] / (a.sums[1] * a.sums[1]));
  }
  for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kLT) {
    const int c = (int)((v / innerv) % a.C);
    float xs[VW], ts[VW], le[VW], ge[VW];
    ld_vec<VW>(a.x, v, xs);
    ld_vec<VW>(a.t, v, ts);
    float sumL = 0.f, sumF = 0.f;
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      float L, F, dL, dF;
      elem_loss<KIND>(a, xs[e], ts[e], c, bal_a, bal_b, L, F, dL, dF);
      if (PASS != 2) {
        sumL += L;
        sumF += F;
        le[e] = L;
        ge[e] = dL * a.gscale;
      } else {
        le[e] = L * invN;
        ge[e] = (dL * invN - coefF * dF) * a.gscale;
      }
    }
    if (PASS != 2) {
      acc[0] += (double)sumL;
      acc[1] += (double)sumF;
    }
    if (PASS != 1) {
      if (a.loss_elem) st_vec<VW>(a.loss_elem, v, le);
      if (a.grad) st_vec<VW>(a.grad, v, ge);
    }
  }
  if (PASS != 2) {
    block_reduce<2>(acc);
    if (threadIdx.x == 0) {
      a.partials[2 * blockIdx.x] = acc[0];
      a.partials[2 * blockIdx.x + 1] = acc[1];
    }
  }
}

// sums[0..1] = deterministic totals; optionally the scalar loss: sum * scale, or sum L / sum F (normalised)
__global__ void __launch_bounds__(kLT) elem_finalize_kernel(const double* partials, int count, double* sums, float* loss_out,
                                                            double scale, int normalized) {
  double s[2];
  reduce_partials<2>(partials, count, s);
  if (threadIdx.x == 0) {
    sums[0] = s[0];
    sums[1] = s[1];
    if (loss_out) *loss_out = (float)(normalized ? s[0] / s[1] : s[0] * scale);
  }
}

// balanced bce: exact counts of t == 1 and t == 0 (balanced_bce.py:26-27)
template <int VW>
__global__ void __launch_bounds__(kLT) count01_kernel(const float* __restrict__ t, long long n, double* partials) {
  double acc[2] = {0.0, 0.0};
  const long long nv = n / VW;
  unsigned c1 = 0, c0 = 0;
  for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < nv; v += (long long)gridDim.x * kLT) {
    float ts[VW];
    ld_vec<VW>(t, v, ts);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      c1 += ts[e] == 1.f ? 1u : 0u;
      c0 += ts[e] == 0.f ? 1u : 0u;
    }
    if ((c1 | c0) & 0x40000000u) {  // flush long before the 32-bit counters can wrap
      acc[0] += (double)c1; acc[1] += (double)c0;
      c1 = c0 = 0;
    }
  }
  acc[0] += (double)c1;
  acc[1] += (double)c0;
  block_reduce<2>(acc);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = acc[0];
    partials[2 * blockIdx.x + 1] = acc[1];
  }
}

__global__ void __launch_bounds__(kLT) count_finalize_kernel(const double* partials, int count, double* counts) {
  double s[2];
  reduce_partials<2>(partials, count, s);
  if (threadIdx.x == 0) {
    counts[0] = s[0];
    counts[1] = s[1];
  }
}

// -------------------------------------------------------------------------------------------------
// soft F1 (soft_f1.py): rows n = 0..N-1, classes c = 0..C-1, element (n, c) at x[n * C + c]  ([N,C] layout)
//   mode 0  BinarySoftF1Loss: C == 1, p = clamp(sigmoid(x)); targets f32 [N]; ignored rows dropped
//   mode 1  SoftF1Loss:       p = clamp(softmax(x, dim=1)); targets i64 [N] one-hot; per-class F1, mean over classes
//   mode 2  SoftF1Loss with ignore_index: elements whose ONE-HOT VALUE equals ignore_index are dropped and the rest is
//           flattened into one F1 (the reference's literal behaviour, soft_f1.py:97-106)
// stats[c] = (tp, fp, fn) double; the gradient kernel recomputes p.
// -------------------------------------------------------------------------------------------------
struct F1Args {
  const float* x;
  const void* t;
  long long N;
  int C, mode, has_ignore;
  float ignore, eps;
  double* partials;  // [grid][Cs][3], Cs = stat classes (C for mode 1, else 1)
  const double* stats;  // [Cs][4] = tp, fp, fn, kept elements
  float* grad;
  float* loss_out;
};

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// one row: probabilities (clamped) of its C classes into p[], raw (unclamped) softmax into q[] (C <= 64 in registers)
constexpr int kF1MaxC = 64;

template <bool GRAD>
__global__ void __launch_bounds__(kLT) f1_kernel(F1Args a) {
  extern __shared__ double sh[];  // stats pass: [Cs][4]
  const int Cs = a.mode == 1 ? a.C : 1;
  if (!GRAD) {
    for (int i = threadIdx.x; i < Cs * 4; i += kLT) sh[i] = 0.0;
    __syncthreads();
  }
  for (long long n = (long long)blockIdx.x * kLT + threadIdx.x; n < a.N; n += (long long)gridDim.x * kLT) {
    const float* xr = a.x + n * a.C;
    float tval = 0.f;
    long long lab = 0;
    if (a.mode == 0) tval = __ldg(static_cast<const float*>(a.t) + n);
    else lab = __ldg(static_cast<const long long*>(a.t) + n);
    if (a.mode == 0) {
      if (a.has_ignore && tval == a.ignore) {
        if (GRAD && a.grad) a.grad[n] = 0.f;
        continue;
      }
      const float s = 1.f / (1.f + expf(-xr[0]));
      const float p = clampf(s, a.eps, 1.f - a.eps);
      if (!GRAD) {
        atomicAdd(&sh[0], (double)(p * tval));
        atomicAdd(&sh[1], (double)(p * (1.f - tval)));
        atomicAdd(&sh[2], (double)((1.f - p) * tval));
        atomicAdd(&sh[3], 1.0);
      } else if (a.grad) {
        const double tp = a.stats[0], D = 2.0 * tp + a.stats[2] + a.stats[1] + (double)a.eps;
        // d(1 - 2tp/D)/dp = -(2 t D - 2 tp)/D^2  (dD/dp = 1)
        const float g = (float)(-(2.0 * tval * D - 2.0 * tp) / (D * D));
        const float pass = (s >= a.eps && s <= 1.f - a.eps) ? 1.f : 0.f;  // clamp backward
        a.grad[n] = a.stats[3] > 0.0 ? g * pass * s * (1.f - s) : 0.f;
      }
      continue;
    }
    // softmax over the row
    float q[kF1MaxC];
    float mx = -INFINITY;
    for (int c = 0; c < a.C; ++c) mx = fmaxf(mx, xr[c]);
    float den = 0.f;
    for (int c = 0; c < a.C; ++c) den += expf(xr[c] - mx);
    const float inv = 1.f / den;
    float gsum = 0.f;  // sum_k g_k q_k for the softmax backward
    for (int c = 0; c < a.C; ++c) {
      q[c] = expf(xr[c] - mx) * inv;
      const float oh = (lab == c) ? 1.f : 0.f;
      const bool kept = !(a.mode == 2 && a.has_ignore && oh == a.ignore);
      const float p = clampf(q[c], a.eps, 1.f - a.eps);
      const int sc = a.mode == 1 ? c : 0;
      if (!GRAD) {
        if (kept) {
          atomicAdd(&sh[4 * sc + 0], (double)(p * oh));
          atomicAdd(&sh[4 * sc + 1], (double)(p * (1.f - oh)));
          atomicAdd(&sh[4 * sc + 2], (double)((1.f - p) * oh));
          atomicAdd(&sh[4 * sc + 3], 1.0);
        }
      } else {
        const double tp = a.stats[4 * sc], D = 2.0 * tp + a.stats[4 * sc + 2] + a.stats[4 * sc + 1] + (double)a.eps;
        float g = (float)(-(2.0 * oh * D - 2.0 * tp) / (D * D)) / (float)(a.mode == 1 ? a.C : 1);
        const bool pass = q[c] >= a.eps && q[c] <= 1.f - a.eps;
        if (!kept || !pass || !(a.stats[4 * sc + 3] > 0.0)) g = 0.f;
        gsum += g * q[c];
        q[c] = g;  // reuse the slot for g_c; q_c is recomputed below
      }
    }
    if (GRAD && a.grad) {
      for (int c = 0; c < a.C; ++c) {
        const float pc = expf(xr[c] - mx) * inv;
        a.grad[n * a.C + c] = pc * (q[c] - gsum);
      }
    }
  }
  if (!GRAD) {
    __syncthreads();
    for (int i = threadIdx.x; i < Cs * 4; i += kLT) a.partials[(size_t)blockIdx.x * Cs * 4 + i] = sh[i];
  }
}

// stats[c][0..3] = sum over CTAs (fixed order); loss = mean_c (1 - 2 tp / (2 tp + fn + fp + eps)); 0 if nothing was kept
__global__ void __launch_bounds__(kLT) f1_finalize_kernel(const double* partials, int grid, int Cs, float eps, double* stats,
                                                          float* loss_out) {
  __shared__ double sh_loss;
  if (threadIdx.x == 0) sh_loss = 0.0;
  __syncthreads();
  for (int c = 0; c < Cs; ++c) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < grid; i += kLT) {
      const double* p = partials + ((size_t)i * Cs + c) * 4;
      v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += p[3];
    }
    block_reduce<4>(v);
    if (threadIdx.x == 0) {
      stats[4 * c] = v[0]; stats[4 * c + 1] = v[1]; stats[4 * c + 2] = v[2]; stats[4 * c + 3] = v[3];
      const double D = 2.0 * v[0] + v[2] + v[1] + (double)eps;
      if (v[3] > 0.0) sh_loss += (1.0 - 2.0 * v[0] / D) / Cs;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = (float)sh_loss;
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_elem_loss_fwd_bwd(const float* x, const float* t, int64_t B, int C, int64_t inner, const tb_elem_loss_params* p,
                         const float* weight, const float* pos_weight, float* loss_out, float* grad, void* ws, size_t ws_bytes,
                         void* stream) {
  TB_REQUIRE(x && t && p && loss_out && ws, TB_ERR_ARG, "tb_elem_loss_fwd_bwd: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_elem_loss_fwd_bwd: bad shape");
  TB_REQUIRE(p->kind >= EL_SOFT_BCE && p->kind <= EL_QFL, TB_ERR_ARG, "tb_elem_loss_fwd_bwd: bad kind %d", p->kind);
  TB_REQUIRE(p->reduction == TB_LOSS_NONE || p->reduction == TB_LOSS_MEAN || p->reduction == TB_LOSS_SUM ||
                 (p->reduction == TB_LOSS_NORMALIZED && p->kind == EL_QFL),
             TB_ERR_ARG, "tb_elem_loss_fwd_bwd: bad reduction %d for kind %d", p->reduction, p->kind);
  TB_REQUIRE(!weight || p->weight_n == 1 || p->weight_n == C, TB_ERR_UNSUPPORTED, "tb_elem_loss_fwd_bwd: weight must have 1 or C entries");
  TB_REQUIRE(!pos_weight || p->pos_weight_n == 1 || p->pos_weight_n == C, TB_ERR_UNSUPPORTED,
             "tb_elem_loss_fwd_bwd: pos_weight must have 1 or C entries");
  TB_REQUIRE(ws_bytes >= tb_loss_workspace_bytes(B, C, inner), TB_ERR_WORKSPACE, "tb_elem_loss_fwd_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const double numel = (double)B * C * inner;
  const bool vec4 = (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0 &&
                    (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0) && (reinterpret_cast<uintptr_t>(loss_out) & 15) == 0;
  ElemArgs a{};
  a.x = x; a.t = t; a.B = B; a.inner = inner; a.C = C; a.p = *p; a.weight = weight; a.pos_weight = pos_weight;
  a.grad = grad;
  a.gscale = p->reduction == TB_LOSS_MEAN ? (float)(1.0 / numel) : 1.f;
  a.loss_elem = p->reduction == TB_LOSS_NONE ? loss_out : nullptr;
  const int grid = loss_grid((long long)(numel / (vec4 ? 4 : 1)));
  double* partials = static_cast<double*>(ws);
  double* sums = partials + 2 * (size_t)grid;
  double* counts = sums + 2;
  a.partials = partials; a.sums = nullptr; a.counts = counts;
  float* scalar_out = p->reduction == TB_LOSS_NONE ? nullptr : loss_out;
  const double scale = p->reduction == TB_LOSS_MEAN ? 1.0 / numel : 1.0;
#define TB_ELEM(KIND, PASS)                                                    \
  do {                                                                         \
    if (vec4) elem_loss_kernel<KIND, PASS, 4><<<grid, kLT, 0, s>>>(a);          \
    else elem_loss_kernel<KIND, PASS, 1><<<grid, kLT, 0, s>>>(a);               \
  } while (0)
  if (p->kind == EL_SOFT_BCE) {
    TB_ELEM(EL_SOFT_BCE, 0);
    elem_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, scale, 0);
  } else if (p->kind == EL_BALANCED_BCE) {
    const long long n = (long long)numel;
    if (vec4) count01_kernel<4><<<grid, kLT, 0, s>>>(t, n, partials);
    else count01_kernel<1><<<grid, kLT, 0, s>>>(t, n, partials);
    count_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, counts);
    TB_ELEM(EL_BALANCED_BCE, 0);
    elem_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, scale, 0);
  } else if (p->reduction != TB_LOSS_NORMALIZED) {
    TB_ELEM(EL_QFL, 0);
    elem_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, scale, 0);
  } else {
    TB_ELEM(EL_QFL, 1);
    elem_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, 1.0, 1);
    if (grad) {
      a.sums = sums;
      TB_ELEM(EL_QFL, 2);
    }
  }
#undef TB_ELEM
  TB_LAUNCH_CHECK("elem_loss_kernel");
  return TB_OK;
}

size_t tb_soft_f1_workspace_bytes(int64_t N, int C) {
  (void)N;
  const size_t grid = (size_t)sm_count() * 8;
  return (grid * (size_t)C * 4 + (size_t)C * 4 + 8) * sizeof(double) + 256;
}

int tb_soft_f1_fwd_bwd(const float* x, const void* t, int64_t N, int C, int mode, int has_ignore, float ignore, float eps,
                       float* loss_out, float* grad, void* ws, size_t ws_bytes, void* stream) {
  TB_REQUIRE(x && t && loss_out && ws, TB_ERR_ARG, "tb_soft_f1_fwd_bwd: null pointer");
  TB_REQUIRE(N > 0 && C > 0 && mode >= 0 && mode <= 2, TB_ERR_ARG, "tb_soft_f1_fwd_bwd: bad argument");
  TB_REQUIRE(mode != 0 || C == 1, TB_ERR_ARG, "tb_soft_f1_fwd_bwd: the binary variant has one logit per sample");
  TB_REQUIRE(C <= kF1MaxC, TB_ERR_UNSUPPORTED, "tb_soft_f1_fwd_bwd: supports C <= %d", kF1MaxC);
  TB_REQUIRE(ws_bytes >= tb_soft_f1_workspace_bytes(N, C), TB_ERR_WORKSPACE, "tb_soft_f1_fwd_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int Cs = mode == 1 ? C : 1;
  F1Args a{};
  a.x = x; a.t = t; a.N = N; a.C = C; a.mode = mode; a.has_ignore = has_ignore; a.ignore = ignore; a.eps = eps;
  a.partials = static_cast<double*>(ws);
  const int grid = loss_grid(N);
  double* stats = a.partials + (size_t)grid * Cs * 4;
  a.stats = stats; a.grad = grad; a.loss_out = loss_out;
  f1_kernel<false><<<grid, kLT, Cs * 4 * sizeof(double), s>>>(a);
  f1_finalize_kernel<<<1, kLT, 0, s>>>(a.partials, grid, Cs, eps, stats, loss_out);
  if (grad) f1_kernel<true><<<grid, kLT, 0, s>>>(a);
  TB_LAUNCH_CHECK("f1_kernel");
  return TB_OK;
}

}  // extern "C"
