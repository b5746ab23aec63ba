// Segmentation-loss kernels for sm_100a: focal (sigmoid / softmax), dice / jaccard, wing.
//
// Every loss is a streaming pass over the logits that produces the loss value AND the gradient w.r.t. the
// logits in the same read of (x, target): HBM-bound, 128-bit loads/stores, grid of 8 CTAs per SM looping
// over the tensor, reductions by warp shuffles -> per-CTA double partials -> one tiny deterministic
// finalisation kernel (no float atomics, so loss values are reproducible run to run).
//
// Reference semantics restated (file:line relative to /root/reference/pytorch_toolbelt/losses):
//   functional.py:19-107   focal_loss_with_logits            -> focal_kernel / focal_softmax_act via pixel kernel
//   functional.py:110-173  softmax_focal_loss_with_logits    -> softmax_focal_kernel
//   functional.py:188-247  soft_jaccard_score / soft_dice_score, dice.py:59-131, jaccard.py:48-103
//                                                            -> region_stats_* / region_grad_* kernels
//   functional.py:250-277  wing_loss                         -> wing_kernel
//   (Lovasz: csrc/lovasz.cu)
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "loss_common.cuh"

namespace tb {

// =================================================================================================
// sigmoid focal
// =================================================================================================
struct FocalDev {
  tb_focal_params p;
  const float* cw;  // [C] or null
};

// One element: loss L, focal term F (as used by the normalisation), dL/dx and dF/dx (0 where masked).
// sig = sigmoid(x) and softplus = log1p(exp(-|x|)) are passed in so that one exp serves both
__device__ __forceinline__ void focal_elem(const FocalDev& f, float x, float t, float prob, float dprob_dx, float sig,
                                           float softplus, int c, float& L, float& F, float& dL, float& dF) {
  const tb_focal_params& p = f.p;
  const float ce = fmaxf(x, 0.f) - x * t + softplus;
  const float pt = prob * t + (1.f - prob) * (1.f - t);
  const float dpt = dprob_dx * (2.f * t - 1.f);
  float focal, dfocal;
  if (!p.has_reduced_threshold) {
    const float base = 1.f - pt;
    focal = powg(base, p.gamma);
    dfocal = -dpowg(base, p.gamma) * dpt;
  } else {
    if (pt < p.reduced_threshold) {
      focal = 1.f;
      dfocal = 0.f;
    } else {
      const float inv = 1.f / (1.f - p.reduced_threshold);
      const float base = (1.f - pt) * inv;
      focal = powg(base, p.gamma);
      dfocal = -dpowg(base, p.gamma) * inv * dpt;
    }
  }
  float scale = 1.f;
  if (p.has_alpha) scale = p.alpha * t + (1.f - p.alpha) * (1.f - t);
  if (f.cw) scale *= __ldg(f.cw + c);
  L = focal * ce * scale;
  dL = (dfocal * ce + focal * (sig - t)) * scale;
  F = focal;
  dF = dfocal;
  if (p.has_ignore && t == p.ignore_index) {
    L = 0.f; dL = 0.f; F = 0.f; dF = 0.f;
  }
}

// FAST: the default BinaryFocalLoss() of BASELINE config 3 (gamma == 2, no alpha / class weights / ignore / reduced threshold):
// closed forms, one ex2 + one rcp + one lg2 per element, no option branches. Same formulas as focal_elem, so the value
// agrees with the general path to rounding (division by the MUFU reciprocal: <= 2 ulp on sigmoid).
__device__ __forceinline__ void joint_elem_fast(float x, float t, float& pr, float& L, float& dL) {
  const float e = __expf(-fabsf(x));
  const float r = __fdividef(1.f, 1.f + e);
  pr = x >= 0.f ? r : e * r;
  const float ce = fmaxf(x, 0.f) - x * t + __logf(1.f + e);
  const float base = pr + t - 2.f * pr * t;       // 1 - pt,  pt = p t + (1 - p)(1 - t)
  const float focal = base * base;
  const float dpt = pr * (1.f - pr) * (2.f * t - 1.f);
  L = focal * ce;
  dL = -2.f * base * dpt * ce + focal * (pr - t);
}

struct FocalArgs {
  const float* x;
  const void* t;
  int t_kind;
  long long B, inner;
  int C;
  FocalDev f;
  float* loss_elem;  // [B,C,inner] or null
  float* grad;       // [B,C,inner] or null
  double* partials;  // [grid,2] (sum L, sum F)
  // second pass of the normalised loss: grad = gs * (dL/N - sumL/N^2 * dF)
  const double* sums;  // [2] (sum L, sum F) or null
  float gscale;        // 1/numel for mean, 1 for sum / none
};

// PASS: 0 = single pass (not normalised): partial sums + loss_elem + grad
//       1 = normalised pass 1: partial sums only
//       2 = normalised pass 2: loss_elem / grad using the sums
template <int PASS, bool LABELS, int VW>
__global__ void __launch_bounds__(kLT) focal_kernel(FocalArgs a) {
  const long long totalv = a.B * a.C * a.inner / VW;  // host guarantees inner % VW == 0
  const long long innerv = a.inner / VW;
  double acc[2] = {0.0, 0.0};
  float invN = 1.f, coefF = 0.f;
  if (PASS == 2) {
    const double N = fmax(a.sums[1], (double)a.f.p.eps);
    invN = (float)(1.0 / N);
    coefF = (a.sums[1] > (double)a.f.p.eps) ? (float)(a.sums[0] / (N * N)) : 0.f;  // clamp_min kills the gradient
  }
  for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kLT) {
    const long long row = v / innerv;  // b*C + c
    const long long posv = v - row * innerv;
    const int c = (int)(row % a.C);
    float xs[VW], tv[VW];
    ld_vec<VW>(a.x, v, xs);
    if (LABELS) {
      const long long b = row / a.C;
      const long long* lp = static_cast<const long long*>(a.t) + b * a.inner + posv * VW;
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const long long lab = __ldg(lp + e);
        // BinaryFocalLoss._one_hot_targets[_with_ignore], focal.py:94-105
        if (a.f.p.has_ignore && (float)lab == a.f.p.ignore_index)
          tv[e] = a.f.p.ignore_index;
        else
          tv[e] = (lab == c) ? 1.f : 0.f;
      }
    } else {
      ld_vec<VW>(static_cast<const float*>(a.t), v, tv);
    }
    float le[VW], ge[VW];
    float sumL = 0.f, sumF = 0.f;  // per-vector float partials: one f32->f64 conversion per vector, not per element
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const float ex = __expf(-fabsf(xs[e]));
      const float inv = 1.f / (1.f + ex);
      const float pr = xs[e] >= 0.f ? inv : ex * inv;  // sigmoid(x) from the same exp as the softplus term
      float L, F, dL, dF;
      focal_elem(a.f, xs[e], tv[e], pr, pr * (1.f - pr), pr, softplus_from_exp(ex), c, L, F, dL, dF);
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

// FAST single pass of the sigmoid focal loss for the default option set (gamma == 2, no class weights / ignore / reduced
// threshold / normalisation; alpha optional): the closed forms of joint_elem_fast instead of focal_elem's option branches,
// pow() and full-precision division -- the generic kernel is issue-bound (0.55 of the HBM peak at 134 M elements), this one
// is back on the memory roofline. Float partial sums are flushed into the double accumulator every 4th vector.
template <bool LABELS, bool ALPHA>
__global__ void __launch_bounds__(kLT) focal_fast_kernel(FocalArgs a) {
  const long long totalv = a.B * a.C * a.inner / 4;
  const long long innerv = a.inner / 4;
  const float al = a.f.p.alpha;
  double acc[2] = {0.0, 0.0};
  float part = 0.f;
  int trips = 0;
  for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kLT) {
    float xs[4], tv[4], le[4], ge[4];
    ld_vec<4>(a.x, v, xs);
    if (LABELS) {
      const long long row = v / innerv;
      const long long posv = v - row * innerv;
      const int c = (int)(row % a.C);
      const long long* lp = static_cast<const long long*>(a.t) + (row / a.C) * a.inner + posv * 4;
      const longlong2 l01 = __ldg(reinterpret_cast<const longlong2*>(lp)), l23 = __ldg(reinterpret_cast<const longlong2*>(lp) + 1);
      tv[0] = l01.x == c ? 1.f : 0.f; tv[1] = l01.y == c ? 1.f : 0.f;
      tv[2] = l23.x == c ? 1.f : 0.f; tv[3] = l23.y == c ? 1.f : 0.f;
    } else {
      ld_vec<4>(static_cast<const float*>(a.t), v, tv);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float pr, L, dL;
      joint_elem_fast(xs[e], tv[e], pr, L, dL);
      if (ALPHA) {
        const float sc = al * tv[e] + (1.f - al) * (1.f - tv[e]);
        L *= sc;
        dL *= sc;
      }
      part += L;
      le[e] = L;
      ge[e] = dL * a.gscale;
    }
    if (a.loss_elem) st_vec<4>(a.loss_elem, v, le);
    if (a.grad) st_vec<4>(a.grad, v, ge);
    if ((++trips & 3) == 0) {
      acc[0] += (double)part;
      part = 0.f;
    }
  }
  acc[0] += (double)part;
  block_reduce<2>(acc);
  if (threadIdx.x == 0) {
    a.partials[2 * blockIdx.x] = acc[0];
    a.partials[2 * blockIdx.x + 1] = 0.0;  // the focal-term sum only matters to the normalised loss (generic kernel)
  }
}

// sums[0..1] = deterministic totals; optionally the scalar loss
__global__ void __launch_bounds__(kLT) focal_finalize_kernel(const double* partials, int count, double* sums,
                                                             float* loss_out, int reduction, int normalized, float eps,
                                                             double numel) {
  double s[2];
  reduce_partials<2>(partials, count, s);
  if (threadIdx.x == 0) {
    sums[0] = s[0];
    sums[1] = s[1];
    if (loss_out) {
      double L = s[0];
      if (normalized) L /= fmax(s[1], (double)eps);
      if (reduction == TB_LOSS_MEAN) L /= numel;
      *loss_out = (float)L;
    }
  }
}

// loss.sum(dim=0): out[c,pos] = sum_b elem[b,c,pos]  (reduction="batchwise_mean", functional.py:104-105)
__global__ void __launch_bounds__(kLT) sum_dim0_kernel(const float* __restrict__ elem, long long B, long long rest,
                                                       float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * kLT + threadIdx.x; i < rest; i += (long long)gridDim.x * kLT) {
    float s = 0.f;
    for (long long b = 0; b < B; ++b) s += elem[b * rest + i];
    out[i] = s;
  }
}

// =================================================================================================
// pixel-wise kernels over the channel dim: softmax focal, sigmoid-focal with softmax activation
// =================================================================================================
constexpr int kMaxC = 64;  // channels held in registers / local memory per pixel

struct PixelFocalArgs {
  const float* x;
  const void* t;  // i64 [B,inner] (softmax focal) or dense f32 / i64 labels (softmax-activated sigmoid focal)
  int t_kind;
  long long B, inner;
  int C;
  int mode;  // 0: softmax_focal_loss_with_logits; 1: focal_loss_with_logits(activation="softmax", softmax_dim=1)
  FocalDev f;             // mode 1 (and gamma / reduced threshold / eps / normalized / reduction for mode 0)
  long long ignore_label; // mode 0
  float* loss_elem;       // mode 0: [B,inner]; mode 1: [B,C,inner]
  float* grad;            // [B,C,inner]
  double* partials;       // [grid,2]
  const double* sums;
  float gscale;
};

template <int PASS, int CB>
__global__ void __launch_bounds__(kLT) pixel_focal_kernel(PixelFocalArgs a) {
  const long long npix = a.B * a.inner;
  const tb_focal_params& p = a.f.p;
  double acc[2] = {0.0, 0.0};
  float invN = 1.f, coefF = 0.f;
  if (PASS == 2) {
    const double N = fmax(a.sums[1], (double)p.eps);
    invN = (float)(1.0 / N);
    coefF = (a.sums[1] > (double)p.eps) ? (float)(a.sums[0] / (N * N)) : 0.f;
  }
  for (long long pix = (long long)blockIdx.x * kLT + threadIdx.x; pix < npix; pix += (long long)gridDim.x * kLT) {
    const long long b = pix / a.inner, pos = pix - b * a.inner;
    const float* xp = a.x + b * a.C * a.inner + pos;
    float xv[CB];
    float mx = -INFINITY;
    #pragma unroll
    for (int c = 0; c < CB; ++c) if (c < a.C) {
      xv[c] = __ldcs(xp + (long long)c * a.inner);
      mx = fmaxf(mx, xv[c]);
    }
    float den = 0.f;
    #pragma unroll
    for (int c = 0; c < CB; ++c) if (c < a.C) den += __expf(xv[c] - mx);
    const float inv_den = 1.f / den;

    if (a.mode == 0) {
      // ---- softmax_focal_loss_with_logits, functional.py:139-173 ----
      const long long lab = __ldg(static_cast<const long long*>(a.t) + pix);
      const bool valid = lab != a.ignore_label;
      const int label = valid ? (int)lab : 0;  // masked_fill(ignore, 0) then one_hot
      float Lpix = 0.f, Fpix = 0.f, S_G = 0.f, S_F = 0.f;
      float G[CB], Fd[CB], direct[CB];
      #pragma unroll
    for (int c = 0; c < CB; ++c) if (c < a.C) {
        const float prob = __expf(xv[c] - mx) * inv_den;
        const float oh = (c == label) ? 1.f : 0.f;
        const float pt = (1.f - oh) * prob + oh * (1.f - prob);
        const float dpt_dp = 1.f - 2.f * oh;
        float focal, dfocal_dp;
        if (!p.has_reduced_threshold) {
          focal = powg(pt, p.gamma);
          dfocal_dp = dpowg(pt, p.gamma) * dpt_dp;
        } else if (pt < p.reduced_threshold) {
          focal = 1.f;
          dfocal_dp = 0.f;
        } else {
          const float inv = 1.f / p.reduced_threshold;
          focal = powg(pt * inv, p.gamma);
          dfocal_dp = dpowg(pt * inv, p.gamma) * inv * dpt_dp;
        }
        const float bce = bce_logits(xv[c], oh);
        const float w = a.f.cw ? __ldg(a.f.cw + c) : 1.f;
        Lpix += focal * bce * w;
        Fpix += focal;  // the reference normaliser sums focal over ALL positions, ignored ones included (:163)
        G[c] = valid ? w * bce * dfocal_dp : 0.f;            // d Lpix / d prob_c
        direct[c] = valid ? w * focal * (sigmoidf_(xv[c]) - oh) : 0.f;  // through the BCE term
        Fd[c] = dfocal_dp;                                   // d Fpix / d prob_c
        S_G += G[c] * prob;
        S_F += Fd[c] * prob;
      }
      if (!valid) Lpix = 0.f;
      if (PASS != 2) {
        acc[0] += (double)Lpix;
        acc[1] += (double)Fpix;
      }
      if (PASS != 1) {
        if (a.loss_elem) a.loss_elem[pix] = (PASS == 2) ? Lpix * invN : Lpix;
        if (a.grad) {
          float* gp = a.grad + b * a.C * a.inner + pos;
          #pragma unroll
    for (int c = 0; c < CB; ++c) if (c < a.C) {
            const float prob = __expf(xv[c] - mx) * inv_den;
            const float dL = prob * (G[c] - S_G) + direct[c];
            float g = dL;
            if (PASS == 2) g = dL * invN - coefF * prob * (Fd[c] - S_F);
            gp[(long long)c * a.inner] = g * a.gscale;
          }
        }
      }
    } else {
      // ---- focal_loss_with_logits(activation="softmax"): p = softmax, CE stays element-wise BCE (:61-67) ----
      float L[CB], F[CB], A[CB], Bq[CB], dLx[CB];
      float S_A = 0.f, S_B = 0.f;
      #pragma unroll
    for (int c = 0; c < CB; ++c) if (c < a.C) {
        const float prob = __expf(xv[c] - mx) * inv_den;
        float t;
        if (a.t_kind == TB_T_LABELS_I64) {
          const long long lab = __ldg(static_cast<const long long*>(a.t) + pix);
          if (p.has_ignore && (float)lab == p.ignore_index) t = p.ignore_index;
          else t = (lab == c) ? 1.f : 0.f;
        } else {
          t = __ldcs(static_cast<const float*>(a.t) + b * a.C * a.inner + (long long)c * a.inner + pos);
        }
        // split dL into the part through prob (dprob = 1 here) and the direct BCE part
        float Lc, Fc, dLp, dFp;
        const float sg = sigmoidf_(xv[c]);
        focal_elem(a.f, xv[c], t, prob, 1.f, sg, softplus_from_exp(__expf(-fabsf(xv[c]))), c, Lc, Fc, dLp, dFp);
        // focal_elem folds focal*(sigmoid-t)*scale into dL: recover the prob-path derivative separately
        float scale = 1.f;
        if (p.has_alpha) scale = p.alpha * t + (1.f - p.alpha) * (1.f - t);
        if (a.f.cw) scale *= __ldg(a.f.cw + c);
        const bool ign = p.has_ignore && t == p.ignore_index;
        const float direct = ign ? 0.f : Fc * (sg - t) * scale;
        const float via_p = dLp - direct;  // = dfocal/dprob * ce * scale
        L[c] = Lc; F[c] = Fc; A[c] = via_p; Bq[c] = dFp; dLx[c] = direct;
        S_A += via_p * prob;
        S_B += dFp * prob;
        if (PASS != 2) {
          acc[0] += (double)Lc;
          acc[1] += (double)Fc;
        }
      }
      if (PASS != 1) {
        #pragma unroll
    for (int c = 0; c < CB; ++c) if (c < a.C) {
          const long long o = b * a.C * a.inner + (long long)c * a.inner + pos;
          const float prob = __expf(xv[c] - mx) * inv_den;
          const float dL = prob * (A[c] - S_A) + dLx[c];
          if (a.loss_elem) a.loss_elem[o] = (PASS == 2) ? L[c] * invN : L[c];
          if (a.grad) {
            float g = dL;
            if (PASS == 2) g = dL * invN - coefF * prob * (Bq[c] - S_B);
            a.grad[o] = g * a.gscale;
          }
        }
      }
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

// softmax_focal_loss_with_logits for ANY class count (C > 64: ADE20K-sized label sets): nothing is held per channel, the
// logits of a pixel are re-read from L1/L2 in four sweeps (max, denominator, loss + the two softmax-backward sums, gradient).
template <int PASS>
__global__ void __launch_bounds__(kLT) pixel_focal_anyc_kernel(PixelFocalArgs a) {
  const long long npix = a.B * a.inner;
  const tb_focal_params& p = a.f.p;
  double acc[2] = {0.0, 0.0};
  float invN = 1.f, coefF = 0.f;
  if (PASS == 2) {
    const double N = fmax(a.sums[1], (double)p.eps);
    invN = (float)(1.0 / N);
    coefF = (a.sums[1] > (double)p.eps) ? (float)(a.sums[0] / (N * N)) : 0.f;
  }
  for (long long pix = (long long)blockIdx.x * kLT + threadIdx.x; pix < npix; pix += (long long)gridDim.x * kLT) {
    const long long b = pix / a.inner, pos = pix - b * a.inner;
    const float* xp = a.x + b * a.C * a.inner + pos;
    const long long lab = __ldg(static_cast<const long long*>(a.t) + pix);
    const bool valid = lab != a.ignore_label;
    const int label = valid ? (int)lab : 0;
    float mx = -INFINITY;
    for (int c = 0; c < a.C; ++c) mx = fmaxf(mx, __ldg(xp + (long long)c * a.inner));
    float den = 0.f;
    for (int c = 0; c < a.C; ++c) den += __expf(__ldg(xp + (long long)c * a.inner) - mx);
    const float inv_den = 1.f / den;
    auto terms = [&](int c, float& prob, float& G, float& direct, float& Fd, float& focal, float& bce_w) {
      const float x = __ldg(xp + (long long)c * a.inner);
      prob = __expf(x - mx) * inv_den;
      const float oh = (c == label) ? 1.f : 0.f;
      const float pt = (1.f - oh) * prob + oh * (1.f - prob);
      const float dpt_dp = 1.f - 2.f * oh;
      float dfocal_dp;
      if (!p.has_reduced_threshold) {
        focal = powg(pt, p.gamma);
        dfocal_dp = dpowg(pt, p.gamma) * dpt_dp;
      } else if (pt < p.reduced_threshold) {
        focal = 1.f;
        dfocal_dp = 0.f;
      } else {
        const float inv = 1.f / p.reduced_threshold;
        focal = powg(pt * inv, p.gamma);
        dfocal_dp = dpowg(pt * inv, p.gamma) * inv * dpt_dp;
      }
      const float bce = bce_logits(x, oh);
      const float w = a.f.cw ? __ldg(a.f.cw + c) : 1.f;
      bce_w = focal * bce * w;
      G = valid ? w * bce * dfocal_dp : 0.f;
      direct = valid ? w * focal * (sigmoidf_(x) - oh) : 0.f;
      Fd = dfocal_dp;
    };
    float Lpix = 0.f, Fpix = 0.f, S_G = 0.f, S_F = 0.f;
    for (int c = 0; c < a.C; ++c) {
      float prob, G, direct, Fd, focal, bw;
      terms(c, prob, G, direct, Fd, focal, bw);
      Lpix += bw;
      Fpix += focal;
      S_G += G * prob;
      S_F += Fd * prob;
    }
    if (!valid) Lpix = 0.f;
    if (PASS != 2) {
      acc[0] += (double)Lpix;
      acc[1] += (double)Fpix;
    }
    if (PASS != 1) {
      if (a.loss_elem) a.loss_elem[pix] = (PASS == 2) ? Lpix * invN : Lpix;
      if (a.grad) {
        float* gp = a.grad + b * a.C * a.inner + pos;
        for (int c = 0; c < a.C; ++c) {
          float prob, G, direct, Fd, focal, bw;
          terms(c, prob, G, direct, Fd, focal, bw);
          const float dL = prob * (G - S_G) + direct;
          float g = dL;
          if (PASS == 2) g = dL * invN - coefF * prob * (Fd - S_F);
          gp[(long long)c * a.inner] = g * a.gscale;
        }
      }
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

// =================================================================================================
// dice / jaccard
// =================================================================================================
struct RegionArgs {
  const float* x;
  const void* t;
  int t_kind;
  long long B, inner;
  int C;
  tb_dice_params p;
  double* partials;       // stats pass: [rows][chunks][3] (elementwise) or [grid][C][3] (multiclass)
  const double* stats;    // grad pass: [C,3]
  const int* classes;
  int n_classes;
  float* grad;
  float* loss_out;
  int chunks;             // CTAs per (b,c) row in the elementwise stats pass
};

// binary / multilabel: rows are (b,c) pairs of `inner` contiguous elements
template <int VW>
__global__ void __launch_bounds__(kLT) region_stats_elem_kernel(RegionArgs a) {
  // 1-D grid: CTA = (row, chunk); rows = B*C may exceed the 65535 limit of gridDim.y
  const long long row = blockIdx.x / a.chunks;  // b*C + c
  const int chunk = (int)(blockIdx.x - row * a.chunks);
  const long long innerv = a.inner / VW;
  const float* xr = a.x + row * a.inner;
  const float* tr = static_cast<const float*>(a.t) + row * a.inner;
  double acc[3] = {0.0, 0.0, 0.0};
  for (long long v = (long long)chunk * kLT + threadIdx.x; v < innerv; v += (long long)a.chunks * kLT) {
    float xs[VW], ts[VW];
    ld_vec<VW>(xr, v, xs);
    ld_vec<VW>(tr, v, ts);
    float I = 0.f, S = 0.f, T = 0.f;
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      float pr = a.p.from_logits ? sigmoidf_(xs[e]) : xs[e];
      float t = ts[e];
      if (a.p.has_ignore && t == a.p.ignore_index) {  // y_pred * mask, y_true * mask (dice.py:85-88,107-111)
        pr = 0.f;
        t = 0.f;
      }
      I += pr * t;
      S += pr + t;
      T += t;
    }
    acc[0] += (double)I;
    acc[1] += (double)S;
    acc[2] += (double)T;
  }
  block_reduce<3>(acc);
  if (threadIdx.x == 0) {
    double* o = a.partials + ((size_t)row * a.chunks + chunk) * 3;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
  }
}

// stats[c] = sum over b and chunks, fixed order
__global__ void __launch_bounds__(kLT) region_stats_elem_finalize(const double* partials, long long B, int C, int chunks,
                                                                  double* stats) {
  const int c = blockIdx.x;
  double v[3] = {0.0, 0.0, 0.0};
  const long long count = B * chunks;
  for (long long i = threadIdx.x; i < count; i += kLT) {
    const long long b = i / chunks, ch = i - b * chunks;
    const double* p = partials + ((size_t)(b * C + c) * chunks + ch) * 3;
    v[0] += p[0]; v[1] += p[1]; v[2] += p[2];
  }
  block_reduce<3>(v);
  if (threadIdx.x == 0) {
    stats[3 * c] = v[0]; stats[3 * c + 1] = v[1]; stats[3 * c + 2] = v[2];
  }
}

// multiclass: one thread per pixel, softmax over C. CB <= 8: per-thread register accumulators, one block reduction
// at the end; larger C: warp-aggregated shared-memory accumulation per pixel.
template <int CB>
__global__ void __launch_bounds__(kLT) region_stats_mc_small_kernel(RegionArgs a) {
  const long long npix = a.B * a.inner;
  const long long* labels = static_cast<const long long*>(a.t);
  float accI[CB], accS[CB], accT[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) accI[c] = accS[c] = accT[c] = 0.f;
  double dI[CB], dS[CB], dT[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) dI[c] = dS[c] = dT[c] = 0.0;
  int since_flush = 0;
  for (long long pix = (long long)blockIdx.x * kLT + threadIdx.x; pix < npix; pix += (long long)gridDim.x * kLT) {
    const long long b = pix / a.inner, pos = pix - b * a.inner;
    const float* xp = a.x + b * a.C * a.inner + pos;
    const long long lab = __ldg(labels + pix);
    const bool valid = !(a.p.has_ignore && (float)lab == a.p.ignore_index);
    float xv[CB];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < a.C) {
        xv[c] = __ldcs(xp + (long long)c * a.inner);
        mx = fmaxf(mx, xv[c]);
      }
    float inv_den = 1.f;
    if (a.p.from_logits) {
      float den = 0.f;
#pragma unroll
      for (int c = 0; c < CB; ++c)
        if (c < a.C) {
          xv[c] = __expf(xv[c] - mx);
          den += xv[c];
        }
      inv_den = 1.f / den;
    }
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < a.C) {
        const float pr = valid ? xv[c] * inv_den : 0.f;
        const float t = (valid && lab == c) ? 1.f : 0.f;
        accI[c] += pr * t;
        accS[c] += pr + t;
        accT[c] += t;
      }
    if (++since_flush == 64) {  // keep the float partials short: flush into doubles
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        dI[c] += accI[c]; dS[c] += accS[c]; dT[c] += accT[c];
        accI[c] = accS[c] = accT[c] = 0.f;
      }
      since_flush = 0;
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    dI[c] += accI[c]; dS[c] += accS[c]; dT[c] += accT[c];
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    double v[3] = {dI[c], dS[c], dT[c]};
    block_reduce<3>(v);
    if (threadIdx.x == 0 && c < a.C) {
      double* o = a.partials + ((size_t)blockIdx.x * a.C + c) * 3;
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
  }
}

__global__ void __launch_bounds__(kLT) region_stats_mc_kernel(RegionArgs a) {
  extern __shared__ double sh_stats[];  // [C][3]
  for (int i = threadIdx.x; i < 3 * a.C; i += kLT) sh_stats[i] = 0.0;
  __syncthreads();
  const long long npix = a.B * a.inner;
  const long long* labels = static_cast<const long long*>(a.t);
  // uniform trip count: the warp shuffles below need every lane of the warp in the loop
  for (long long base = (long long)blockIdx.x * kLT; base < npix; base += (long long)gridDim.x * kLT) {
    const long long pix_raw = base + threadIdx.x;
    const bool in_range = pix_raw < npix;
    const long long pix = in_range ? pix_raw : npix - 1;
    const long long b = pix / a.inner, pos = pix - b * a.inner;
    const float* xp = a.x + b * a.C * a.inner + pos;
    const long long lab = __ldg(labels + pix);
    const bool valid = in_range && !(a.p.has_ignore && (float)lab == a.p.ignore_index);
    float mx = -INFINITY;
    if (a.p.from_logits) {
      for (int c = 0; c < a.C; ++c) mx = fmaxf(mx, __ldcs(xp + (long long)c * a.inner));
    }
    float den = 0.f;
    if (a.p.from_logits) {
      for (int c = 0; c < a.C; ++c) den += __expf(__ldg(xp + (long long)c * a.inner) - mx);
    }
    const float inv_den = a.p.from_logits ? 1.f / den : 1.f;
    for (int c = 0; c < a.C; ++c) {
      const float xc = __ldg(xp + (long long)c * a.inner);
      float pr = a.p.from_logits ? __expf(xc - mx) * inv_den : xc;
      float t = (valid && lab == c) ? 1.f : 0.f;
      if (!valid) pr = 0.f;  // ignored pixel (dice.py:93-96) or padding lane
      // warp-aggregate before touching shared memory
      float I = pr * t, S = pr + t, T = t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        I += __shfl_xor_sync(0xffffffffu, I, o);
        S += __shfl_xor_sync(0xffffffffu, S, o);
        T += __shfl_xor_sync(0xffffffffu, T, o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_stats[3 * c], (double)I);
        atomicAdd(&sh_stats[3 * c + 1], (double)S);
        atomicAdd(&sh_stats[3 * c + 2], (double)T);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * a.C; i += kLT) a.partials[(size_t)blockIdx.x * 3 * a.C + i] = sh_stats[i];
}

__global__ void __launch_bounds__(kLT) region_stats_mc_finalize(const double* partials, int grid, int C, double* stats) {
  const int c = blockIdx.x;
  double v[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < grid; i += kLT) {
    const double* p = partials + ((size_t)i * C + c) * 3;
    v[0] += p[0]; v[1] += p[1]; v[2] += p[2];
  }
  block_reduce<3>(v);
  if (threadIdx.x == 0) {
    stats[3 * c] = v[0]; stats[3 * c + 1] = v[1]; stats[3 * c + 2] = v[2];
  }
}

// Per-class loss term and the affine gradient coefficients  dLoss/dp_{c,i} = alpha_c * t_i + beta_c
__device__ __forceinline__ void region_class_terms(const tb_dice_params& p, const double* st, double& loss_c,
                                                   float& alpha, float& beta) {
  const double I = st[0], S = st[1], T = st[2];
  const double sm = p.smooth, eps = p.eps;
  double score, ds_t, ds_0;  // dscore/dp = ds_t * t + ds_0
  if (!p.jaccard) {
    const double raw = S + sm, D = fmax(raw, eps);
    score = (2.0 * I + sm) / D;
    const double uncl = raw > eps ? 1.0 : 0.0;
    ds_t = 2.0 / D;
    ds_0 = -(2.0 * I + sm) * uncl / (D * D);
  } else {
    const double raw = S - I + sm, U = fmax(raw, eps);
    score = (I + sm) / U;
    const double uncl = raw > eps ? 1.0 : 0.0;
    // d/dp [(I+sm)/U] with dI = t, dU = (1 - t) when unclamped
    ds_t = 1.0 / U + (I + sm) * uncl / (U * U);
    ds_0 = -(I + sm) * uncl / (U * U);
  }
  double dl;  // d loss_c / d score
  if (p.log_loss) {
    loss_c = -log(fmax(score, eps));
    dl = score > eps ? -1.0 / score : 0.0;
  } else {
    loss_c = 1.0 - score;
    dl = -1.0;
  }
  const double m = T > 0.0 ? 1.0 : 0.0;  // classes without true pixels contribute nothing (dice.py:124-125)
  loss_c *= m;
  alpha = (float)(m * dl * ds_t);
  beta = (float)(m * dl * ds_0);
}

}  // namespace tb
#include "pixel_vec4.cuh"
namespace tb {

// VW: 1 / 4 = elementwise scalar / vector; 8 / 12 = multiclass with 4 pixels per thread (pixel_vec4.cuh), C <= 8 / C <= 4
template <int VW>
__global__ void __launch_bounds__(kLT) region_grad_kernel(RegionArgs a) {
  extern __shared__ float sh_coef[];  // [C][2] alpha, beta (already divided by the number of averaged classes)
  __shared__ double sh_loss;
  if (threadIdx.x == 0) sh_loss = 0.0;
  __syncthreads();
  const int nsel = a.classes ? a.n_classes : a.C;
  for (int c = threadIdx.x; c < a.C; c += kLT) {
    double lc;
    float al, be;
    region_class_terms(a.p, a.stats + 3 * c, lc, al, be);
    int mult = 1;
    if (a.classes) {
      mult = 0;
      for (int k = 0; k < a.n_classes; ++k) mult += (a.classes[k] == c);
    }
    sh_coef[2 * c] = al * (float)mult / (float)nsel;
    sh_coef[2 * c + 1] = be * (float)mult / (float)nsel;
    if (blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(&sh_loss, lc * mult / nsel);
  }
  __syncthreads();
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && a.loss_out) *a.loss_out = (float)sh_loss;
  if (!a.grad) return;

  if (VW == 8 || VW == 12) {
    region_grad_mc_vec4<VW == 8 ? 8 : 4>(a, sh_coef);
    return;
  }
  if (a.p.mode != 1) {
    // binary / multilabel: elementwise
    constexpr int EV = VW >= 8 ? 1 : VW;
    const long long totalv = a.B * a.C * a.inner / EV, innerv = a.inner / EV;
    for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kLT) {
      const int c = (int)((v / innerv) % a.C);
      const float al = sh_coef[2 * c], be = sh_coef[2 * c + 1];
      float xs[EV], ts[EV], g[EV];
      ld_vec<EV>(a.x, v, xs);
      ld_vec<EV>(static_cast<const float*>(a.t), v, ts);
#pragma unroll
      for (int e = 0; e < EV; ++e) {
        const float t = ts[e];
        const bool ign = a.p.has_ignore && t == a.p.ignore_index;
        float dp = 1.f;
        if (a.p.from_logits) {
          const float pr = sigmoidf_(xs[e]);
          dp = pr * (1.f - pr);
        }
        g[e] = ign ? 0.f : (al * t + be) * dp;
      }
      st_vec<EV>(a.grad, v, g);
    }
  } else {
    const long long npix = a.B * a.inner;
    const long long* labels = static_cast<const long long*>(a.t);
    for (long long pix = (long long)blockIdx.x * kLT + threadIdx.x; pix < npix; pix += (long long)gridDim.x * kLT) {
      const long long b = pix / a.inner, pos = pix - b * a.inner;
      const float* xp = a.x + b * a.C * a.inner + pos;
      float* gp = a.grad + b * a.C * a.inner + pos;
      const long long lab = __ldg(labels + pix);
      const bool valid = !(a.p.has_ignore && (float)lab == a.p.ignore_index);
      if (!valid) {
        for (int c = 0; c < a.C; ++c) gp[(long long)c * a.inner] = 0.f;
        continue;
      }
      if (!a.p.from_logits) {
        for (int c = 0; c < a.C; ++c) gp[(long long)c * a.inner] = sh_coef[2 * c] * (lab == c ? 1.f : 0.f) + sh_coef[2 * c + 1];
        continue;
      }
      if (a.C <= 8) {
        float pv[8];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < a.C) {
            pv[c] = __ldcs(xp + (long long)c * a.inner);
            mx = fmaxf(mx, pv[c]);
          }
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < a.C) {
            pv[c] = __expf(pv[c] - mx);
            den += pv[c];
          }
        const float inv_den = 1.f / den;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < a.C) {
            pv[c] *= inv_den;
            dot += (sh_coef[2 * c] * (lab == c ? 1.f : 0.f) + sh_coef[2 * c + 1]) * pv[c];
          }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < a.C) {
            const float gc = sh_coef[2 * c] * (lab == c ? 1.f : 0.f) + sh_coef[2 * c + 1];
            __stcs(gp + (long long)c * a.inner, pv[c] * (gc - dot));
          }
        continue;
      }
      float mx = -INFINITY;
      for (int c = 0; c < a.C; ++c) mx = fmaxf(mx, __ldcs(xp + (long long)c * a.inner));
      float den = 0.f;
      for (int c = 0; c < a.C; ++c) den += __expf(__ldg(xp + (long long)c * a.inner) - mx);
      const float inv_den = 1.f / den;
      float dot = 0.f;  // sum_k g_k p_k
      for (int c = 0; c < a.C; ++c) {
        const float pr = __expf(__ldg(xp + (long long)c * a.inner) - mx) * inv_den;
        dot += (sh_coef[2 * c] * (lab == c ? 1.f : 0.f) + sh_coef[2 * c + 1]) * pr;
      }
      for (int c = 0; c < a.C; ++c) {
        const float pr = __expf(__ldg(xp + (long long)c * a.inner) - mx) * inv_den;
        const float gc = sh_coef[2 * c] * (lab == c ? 1.f : 0.f) + sh_coef[2 * c + 1];
        gp[(long long)c * a.inner] = pr * (gc - dot);
      }
    }
  }
}

// =================================================================================================
// joint sigmoid-focal + dice / jaccard on a dense target (BASELINE config 3: BinaryFocalLoss + DiceLoss)
// Two passes over (x, t) and ONE gradient write: 2*(N+N)*4 + N*4 bytes instead of the 3N*4 (focal) + 5N*4
// (dice) + 3N*4 (adding the two gradients) of the separate kernels. One exp per element per pass serves the
// sigmoid, the softplus of the BCE term and the dice probability.
// =================================================================================================
struct JointArgs {
  const float* x;
  const float* t;
  long long B, inner;
  int C;
  FocalDev f;        // sigmoid activation, not normalised, reduction MEAN or SUM
  tb_dice_params d;  // mode 0 / 2, from_logits
  const int* classes;
  int n_classes;
  float w_focal, w_dice;
  float gscale;      // focal: 1/numel for MEAN, 1 for SUM
  double* partials;  // pass 1: [rows][chunks][4] = (I, S, T, focal sum)
  double* stats;     // [C][4], written by the finalize kernel, read by pass 2
  float* grad;
  float* loss_out;   // [3]: total, focal, dice
  int chunks;
};

// shared by both passes so that they see identical p / loss terms (joint_elem_fast: above, next to focal_elem)
__device__ __forceinline__ void joint_elem(const JointArgs& a, float x, float t, int c, float& pr, float& L, float& dL) {
  const float ex = __expf(-fabsf(x));
  const float inv = 1.f / (1.f + ex);
  pr = x >= 0.f ? inv : ex * inv;
  float F, dF;
  focal_elem(a.f, x, t, pr, pr * (1.f - pr), pr, softplus_from_exp(ex), c, L, F, dL, dF);
}

template <int VW, bool FAST>
__global__ void __launch_bounds__(kLT) joint_stats_kernel(JointArgs a) {
  const long long row = blockIdx.x / a.chunks;  // b*C + c (1-D grid: rows may exceed the gridDim.y limit)
  const int chunk = (int)(blockIdx.x - row * a.chunks);
  const int c = (int)(row % a.C);
  const long long innerv = a.inner / VW;
  const float* xr = a.x + row * a.inner;
  const float* tr = a.t + row * a.inner;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  float part[4] = {0.f, 0.f, 0.f, 0.f};  // float partials over a few vectors, flushed into the doubles every 4th trip
  int trips = 0;
  for (long long v = (long long)chunk * kLT + threadIdx.x; v < innerv; v += (long long)a.chunks * kLT) {
    float xs[VW], ts[VW];
    ld_vec<VW>(xr, v, xs);
    ld_vec<VW>(tr, v, ts);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      float pr, L, dL;
      if (FAST) joint_elem_fast(xs[e], ts[e], pr, L, dL);
      else joint_elem(a, xs[e], ts[e], c, pr, L, dL);
      part[3] += L;
      float t = ts[e];
      if (!FAST && a.d.has_ignore && t == a.d.ignore_index) {  // dice.py:85-88,107-111
        pr = 0.f;
        t = 0.f;
      }
      part[0] += pr * t;
      part[1] += pr + t;
      part[2] += t;
    }
    if ((++trips & 3) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[q] += (double)part[q];
        part[q] = 0.f;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] += (double)part[q];
  block_reduce<4>(acc);
  if (threadIdx.x == 0) {
    double* o = a.partials + ((size_t)row * a.chunks + chunk) * 4;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
  }
}

// stats[c] = sum over b and chunks in a fixed order (deterministic)
__global__ void __launch_bounds__(kLT) joint_finalize_kernel(const double* partials, long long B, int C, int chunks,
                                                             double* stats) {
  const int c = blockIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  const long long count = B * chunks;
  for (long long i = threadIdx.x; i < count; i += kLT) {
    const long long b = i / chunks, ch = i - b * chunks;
    const double* p = partials + ((size_t)(b * C + c) * chunks + ch) * 4;
    v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += p[3];
  }
  block_reduce<4>(v);
  if (threadIdx.x == 0) {
    stats[4 * c] = v[0]; stats[4 * c + 1] = v[1]; stats[4 * c + 2] = v[2]; stats[4 * c + 3] = v[3];
  }
}

template <int VW, bool FAST>
__global__ void __launch_bounds__(kLT) joint_grad_kernel(JointArgs a) {
  extern __shared__ float sh_coef[];  // [C][2]: w_dice * (alpha, beta) / number of averaged classes
  const int nsel = a.classes ? a.n_classes : a.C;
  for (int c = threadIdx.x; c < a.C; c += kLT) {
    double lc;
    float al, be;
    region_class_terms(a.d, a.stats + 4 * c, lc, al, be);
    int mult = 1;
    if (a.classes) {
      mult = 0;
      for (int k = 0; k < a.n_classes; ++k) mult += (a.classes[k] == c);
    }
    sh_coef[2 * c] = a.w_dice * al * (float)mult / (float)nsel;
    sh_coef[2 * c + 1] = a.w_dice * be * (float)mult / (float)nsel;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.loss_out) {  // serial, fixed order: deterministic
    double dice = 0.0, focal = 0.0;
    for (int c = 0; c < a.C; ++c) {
      double lc;
      float al, be;
      region_class_terms(a.d, a.stats + 4 * c, lc, al, be);
      int mult = 1;
      if (a.classes) {
        mult = 0;
        for (int k = 0; k < a.n_classes; ++k) mult += (a.classes[k] == c);
      }
      dice += lc * mult / nsel;
      focal += a.stats[4 * c + 3];
    }
    focal *= (double)a.gscale;
    a.loss_out[0] = (float)((double)a.w_focal * focal + (double)a.w_dice * dice);
    a.loss_out[1] = (float)focal;
    a.loss_out[2] = (float)dice;
  }
  __syncthreads();
  if (!a.grad) return;
  const long long totalv = a.B * a.C * a.inner / VW, innerv = a.inner / VW;
  const float fscale = a.w_focal * a.gscale;
  for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kLT) {
    const int c = (int)((v / innerv) % a.C);
    const float al = sh_coef[2 * c], be = sh_coef[2 * c + 1];
    float xs[VW], ts[VW], g[VW];
    ld_vec<VW>(a.x, v, xs);
    ld_vec<VW>(a.t, v, ts);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      float pr, L, dL;
      if (FAST) joint_elem_fast(xs[e], ts[e], pr, L, dL);
      else joint_elem(a, xs[e], ts[e], c, pr, L, dL);
      const float t = ts[e];
      const bool ign = !FAST && a.d.has_ignore && t == a.d.ignore_index;
      const float gd = ign ? 0.f : (al * t + be) * (pr * (1.f - pr));
      g[e] = fscale * dL + gd;
    }
    st_vec<VW>(a.grad, v, g);
  }
}

// =================================================================================================
// wing
// =================================================================================================
struct WingArgs {
  const float* x;
  const float* t;
  long long n;
  float width, curvature, inv_curvature, cconst, gscale;
  float* loss_elem;
  float* grad;
  double* partials;
};

__device__ __forceinline__ void wing_elem(const WingArgs& a, float x, float t, float& L, float& g) {
  const float diff = t - x;
  const float d = fabsf(diff);
  const float sgn = (diff > 0.f) ? 1.f : (diff < 0.f ? -1.f : 0.f);
  if (d < a.width) {
    L = a.width * __logf(1.f + d * a.inv_curvature);
    g = -sgn * __fdividef(a.width, a.curvature + d);
  } else {
    L = d - a.cconst;
    g = -sgn;
  }
}

__global__ void __launch_bounds__(kLT) wing_kernel(WingArgs a) {
  double acc[1] = {0.0};
  const long long n4 = a.n / 4;
  for (long long v = (long long)blockIdx.x * kLT + threadIdx.x; v < n4; v += (long long)gridDim.x * kLT) {
    const float4 x4 = __ldcs(reinterpret_cast<const float4*>(a.x) + v), t4 = __ldcs(reinterpret_cast<const float4*>(a.t) + v);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ts[4] = {t4.x, t4.y, t4.z, t4.w};
    float L[4], g[4];
    float sumL = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      wing_elem(a, xs[e], ts[e], L[e], g[e]);
      sumL += L[e];
      g[e] *= a.gscale;
    }
    acc[0] += (double)sumL;
    if (a.loss_elem) __stcs(reinterpret_cast<float4*>(a.loss_elem) + v, make_float4(L[0], L[1], L[2], L[3]));
    if (a.grad) __stcs(reinterpret_cast<float4*>(a.grad) + v, make_float4(g[0], g[1], g[2], g[3]));
  }
  // tail (n % 4 elements) handled by the first threads of CTA 0
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    float L, g;
    wing_elem(a, a.x[i], a.t[i], L, g);
    acc[0] += (double)L;
    if (a.loss_elem) a.loss_elem[i] = L;
    if (a.grad) a.grad[i] = g * a.gscale;
  }
  block_reduce<1>(acc);
  if (threadIdx.x == 0) a.partials[blockIdx.x] = acc[0];
}

__global__ void __launch_bounds__(kLT) scalar_finalize_kernel(const double* partials, int count, double scale,
                                                              float* loss_out) {
  double s[1];
  reduce_partials<1>(partials, count, s);
  if (threadIdx.x == 0 && loss_out) *loss_out = (float)(s[0] * scale);
}

// data[i] *= *scale unless *scale == 1 (autograd hands the upstream gradient as a device scalar; no host sync)
__global__ void __launch_bounds__(kLT) scale_by_device_scalar_kernel(float* data, long long n, const float* scale) {
  const float sc = __ldg(scale);
  if (sc == 1.f) return;
  for (long long i = (long long)blockIdx.x * kLT + threadIdx.x; i < n; i += (long long)gridDim.x * kLT) data[i] *= sc;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace tb

using namespace tb;

extern "C" {

size_t tb_loss_workspace_bytes(int64_t B, int C, int64_t inner) {
  const size_t grid = (size_t)sm_count() * 8;
  // focal: partials [grid,2] + sums [2]; dice elementwise stats: [B*C*chunks,3]; multiclass: [grid,C,3]
  const size_t chunks = 64;
  size_t bytes = 0;
  bytes = std::max(bytes, (grid * 2 + 8) * sizeof(double));
  bytes = std::max(bytes, ((size_t)B * C * chunks * 3 + 8) * sizeof(double));
  bytes = std::max(bytes, (grid * (size_t)C * 3 + 8) * sizeof(double));
  bytes = std::max(bytes, ((size_t)B * C * chunks * 4 + (size_t)C * 4 + 8) * sizeof(double));  // joint focal + dice
  (void)inner;
  return align_up(bytes, 256) + 256;
}

}  // extern "C"

template <int PASS>
static void launch_pixel_focal(const PixelFocalArgs& a, int grid, cudaStream_t s) {
  if (a.C > kMaxC && a.mode == 0) {
    pixel_focal_anyc_kernel<PASS><<<grid, kLT, 0, s>>>(a);
    return;
  }
  if (a.C <= 4) pixel_focal_kernel<PASS, 4><<<grid, kLT, 0, s>>>(a);
  else if (a.C <= 8) pixel_focal_kernel<PASS, 8><<<grid, kLT, 0, s>>>(a);
  else if (a.C <= 16) pixel_focal_kernel<PASS, 16><<<grid, kLT, 0, s>>>(a);
  else if (a.C <= 32) pixel_focal_kernel<PASS, 32><<<grid, kLT, 0, s>>>(a);
  else pixel_focal_kernel<PASS, kMaxC><<<grid, kLT, 0, s>>>(a);
}

extern "C" {

static int reduction_ok(int reduction) { return reduction >= TB_LOSS_NONE && reduction <= TB_LOSS_BATCHWISE; }

int tb_focal_fwd_bwd(const float* x, const void* t, int t_kind, int64_t B, int C, int64_t inner,
                     const tb_focal_params* p, const float* class_w, float* loss_out, float* grad, void* ws,
                     size_t ws_bytes, void* stream) {
  TB_REQUIRE(x && t && p && loss_out && ws, TB_ERR_ARG, "tb_focal_fwd_bwd: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_focal_fwd_bwd: bad shape");
  TB_REQUIRE(reduction_ok(p->reduction), TB_ERR_ARG, "tb_focal_fwd_bwd: bad reduction %d", p->reduction);
  TB_REQUIRE(t_kind == TB_T_DENSE_F32 || t_kind == TB_T_LABELS_I64, TB_ERR_ARG, "tb_focal_fwd_bwd: bad t_kind");
  TB_REQUIRE(ws_bytes >= tb_loss_workspace_bytes(B, C, inner), TB_ERR_WORKSPACE, "tb_focal_fwd_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const double numel = (double)B * C * inner;
  double* partials = static_cast<double*>(ws);
  const bool elem_out = p->reduction == TB_LOSS_NONE;
  const bool batchwise = p->reduction == TB_LOSS_BATCHWISE;
  const float gscale = p->reduction == TB_LOSS_MEAN ? (float)(1.0 / numel) : 1.f;

  if (p->softmax) {
    TB_REQUIRE(C <= kMaxC, TB_ERR_UNSUPPORTED, "tb_focal_fwd_bwd: softmax activation supports C <= %d", kMaxC);
    TB_REQUIRE(!batchwise, TB_ERR_UNSUPPORTED, "tb_focal_fwd_bwd: batchwise reduction with softmax activation");
    PixelFocalArgs a;
    a.x = x; a.t = t; a.t_kind = t_kind; a.B = B; a.inner = inner; a.C = C; a.mode = 1;
    a.f.p = *p; a.f.cw = class_w; a.ignore_label = 0;
    a.loss_elem = elem_out ? loss_out : nullptr; a.grad = grad; a.gscale = gscale;
    const int grid = loss_grid(B * inner);
    a.partials = partials; a.sums = nullptr;
    double* sums = partials + 2 * (size_t)grid;
    if (!p->normalized) {
      launch_pixel_focal<0>(a, grid, s);
      focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, elem_out ? nullptr : loss_out, p->reduction, 0, p->eps, numel);
    } else {
      launch_pixel_focal<1>(a, grid, s);
      focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, elem_out ? nullptr : loss_out, p->reduction, 1, p->eps, numel);
      if (elem_out || grad) {
        a.sums = sums;
        launch_pixel_focal<2>(a, grid, s);
      }
    }
    TB_LAUNCH_CHECK("pixel_focal_kernel");
    return TB_OK;
  }

  const bool vec4 = (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (t_kind == TB_T_LABELS_I64 || (reinterpret_cast<uintptr_t>(t) & 15) == 0) &&
                    (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0) &&
                    (reinterpret_cast<uintptr_t>(loss_out) & 15) == 0;
  FocalArgs a;
  a.x = x; a.t = t; a.t_kind = t_kind; a.B = B; a.inner = inner; a.C = C;
  a.f.p = *p; a.f.cw = class_w;
  a.grad = grad; a.gscale = gscale; a.sums = nullptr;
  const int grid = loss_grid(B * C * inner / (vec4 ? 4 : 1));
  a.partials = partials;
  double* sums = partials + 2 * (size_t)grid;
  float* elem_buf = nullptr;
  if (elem_out) elem_buf = loss_out;
  if (batchwise) elem_buf = loss_out + (size_t)C * inner;  // scratch for the element losses behind the result (tb200.h)
  a.loss_elem = elem_buf;
  const bool labels = t_kind == TB_T_LABELS_I64;
#define TB_FOCAL_LAUNCH(PASS)                                            \
  do {                                                                   \
    if (labels && vec4)                                                  \
      focal_kernel<PASS, true, 4><<<grid, kLT, 0, s>>>(a);               \
    else if (labels)                                                     \
      focal_kernel<PASS, true, 1><<<grid, kLT, 0, s>>>(a);               \
    else if (vec4)                                                       \
      focal_kernel<PASS, false, 4><<<grid, kLT, 0, s>>>(a);              \
    else                                                                 \
      focal_kernel<PASS, false, 1><<<grid, kLT, 0, s>>>(a);              \
  } while (0)
  float* scalar_out = (elem_out || batchwise) ? nullptr : loss_out;
  // default option set -> closed-form kernel (TB200_FOCAL_GENERIC=1 keeps the generic one: A/B and parity tests)
  static const bool force_generic = getenv("TB200_FOCAL_GENERIC") && atoi(getenv("TB200_FOCAL_GENERIC")) != 0;
  const bool labels16 = !labels || (inner % 4 == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0);
  const bool fast = vec4 && labels16 && !p->normalized && p->gamma == 2.f && !class_w && !p->has_ignore && !p->has_reduced_threshold &&
                    !force_generic;
  if (fast) {
    if (labels) {
      if (p->has_alpha) focal_fast_kernel<true, true><<<grid, kLT, 0, s>>>(a);
      else focal_fast_kernel<true, false><<<grid, kLT, 0, s>>>(a);
    } else {
      if (p->has_alpha) focal_fast_kernel<false, true><<<grid, kLT, 0, s>>>(a);
      else focal_fast_kernel<false, false><<<grid, kLT, 0, s>>>(a);
    }
    focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, p->reduction, 0, p->eps, numel);
  } else if (!p->normalized) {
    TB_FOCAL_LAUNCH(0);
    focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, p->reduction, 0, p->eps, numel);
  } else {
    TB_FOCAL_LAUNCH(1);
    focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, p->reduction, 1, p->eps, numel);
    if (elem_buf || grad) {
      a.sums = sums;
      TB_FOCAL_LAUNCH(2);
    }
  }
#undef TB_FOCAL_LAUNCH
  if (batchwise) {
    const long long rest = (long long)C * inner;
    sum_dim0_kernel<<<loss_grid(rest), kLT, 0, s>>>(elem_buf, B, rest, loss_out);
  }
  TB_LAUNCH_CHECK("focal_kernel");
  return TB_OK;
}

int tb_softmax_focal_fwd_bwd(const float* x, const int64_t* t, int64_t B, int C, int64_t inner, float gamma,
                             float reduced_threshold, int has_reduced_threshold, int64_t ignore_index, int normalized,
                             float eps, int reduction, const float* class_w, float* loss_out, float* grad, void* ws,
                             size_t ws_bytes, void* stream) {
  TB_REQUIRE(x && t && loss_out && ws, TB_ERR_ARG, "tb_softmax_focal_fwd_bwd: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_softmax_focal_fwd_bwd: bad shape");
  TB_REQUIRE(reduction_ok(reduction), TB_ERR_ARG, "tb_softmax_focal_fwd_bwd: bad reduction %d", reduction);
  TB_REQUIRE(ws_bytes >= tb_loss_workspace_bytes(B, C, inner), TB_ERR_WORKSPACE, "tb_softmax_focal_fwd_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const double npos = (double)B * inner;  // loss.mean() runs over [B, *] after sum(dim=1) (functional.py:159,166-167)
  PixelFocalArgs a;
  a.x = x; a.t = t; a.t_kind = TB_T_LABELS_I64; a.B = B; a.inner = inner; a.C = C; a.mode = 0;
  a.f.p = tb_focal_params{};
  a.f.p.gamma = gamma; a.f.p.reduced_threshold = reduced_threshold; a.f.p.has_reduced_threshold = has_reduced_threshold;
  a.f.p.normalized = normalized; a.f.p.eps = eps; a.f.p.reduction = reduction;
  a.f.cw = class_w; a.ignore_label = ignore_index;
  const bool elem_out = reduction == TB_LOSS_NONE, batchwise = reduction == TB_LOSS_BATCHWISE;
  float* elem_buf = elem_out ? loss_out : (batchwise ? loss_out + inner : nullptr);
  a.loss_elem = elem_buf; a.grad = grad;
  a.gscale = reduction == TB_LOSS_MEAN ? (float)(1.0 / npos) : 1.f;
  const int grid = loss_grid(B * inner);
  double* partials = static_cast<double*>(ws);
  double* sums = partials + 2 * (size_t)grid;
  a.partials = partials; a.sums = nullptr;
  float* scalar_out = (elem_out || batchwise) ? nullptr : loss_out;
  const bool quads = !normalized && !has_reduced_threshold && C <= 8 && (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(t) & 15) == 0 && (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0) &&
                     (!elem_buf || (reinterpret_cast<uintptr_t>(elem_buf) & 15) == 0);
  if (quads) {
    const int gq = loss_grid(B * inner / 4);
    a.partials = partials;
    double* sums_q = partials + 2 * (size_t)gq;
    const bool g2 = gamma == 2.f && !class_w;  // the defaults: no pow(), smaller register budget
    static const int minb = getenv("TB200_SFOCAL_MINB") ? atoi(getenv("TB200_SFOCAL_MINB")) : 3;  // A/B: register budget
    if (C <= 4 && g2 && minb == 4) softmax_focal_vec4_kernel<4, true, 4><<<gq, kLT, 0, s>>>(a);
    else if (C <= 4 && g2) softmax_focal_vec4_kernel<4, true, 3><<<gq, kLT, 0, s>>>(a);
    else if (C <= 4) softmax_focal_vec4_kernel<4, false, 1><<<gq, kLT, 0, s>>>(a);
    else if (g2) softmax_focal_vec4_kernel<8, true, 1><<<gq, kLT, 0, s>>>(a);
    else softmax_focal_vec4_kernel<8, false, 1><<<gq, kLT, 0, s>>>(a);
    focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, gq, sums_q, scalar_out, reduction, 0, eps, npos);
  } else if (!normalized) {
    launch_pixel_focal<0>(a, grid, s);
    focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, reduction, 0, eps, npos);
  } else {
    launch_pixel_focal<1>(a, grid, s);
    focal_finalize_kernel<<<1, kLT, 0, s>>>(partials, grid, sums, scalar_out, reduction, 1, eps, npos);
    if (elem_buf || grad) {
      a.sums = sums;
      launch_pixel_focal<2>(a, grid, s);
    }
  }
  if (batchwise) sum_dim0_kernel<<<loss_grid(inner), kLT, 0, s>>>(elem_buf, B, inner, loss_out);
  TB_LAUNCH_CHECK("pixel_focal_kernel(softmax)");
  return TB_OK;
}

int tb_dice_stats(const float* x, const void* t, int t_kind, int64_t B, int C, int64_t inner, const tb_dice_params* p,
                  double* stats_out, void* ws, size_t ws_bytes, void* stream) {
  TB_REQUIRE(x && t && p && stats_out && ws, TB_ERR_ARG, "tb_dice_stats: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_dice_stats: bad shape");
  TB_REQUIRE(p->mode >= 0 && p->mode <= 2, TB_ERR_ARG, "tb_dice_stats: bad mode %d", p->mode);
  TB_REQUIRE(ws_bytes >= tb_loss_workspace_bytes(B, C, inner), TB_ERR_WORKSPACE, "tb_dice_stats: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  RegionArgs a{};
  a.x = x; a.t = t; a.t_kind = t_kind; a.B = B; a.inner = inner; a.C = C; a.p = *p;
  a.partials = static_cast<double*>(ws);
  if (p->mode == 1) {
    TB_REQUIRE(t_kind == TB_T_LABELS_I64, TB_ERR_ARG, "tb_dice_stats: multiclass needs int64 labels");
    TB_REQUIRE(C <= 1024, TB_ERR_UNSUPPORTED, "tb_dice_stats: multiclass supports C <= 1024");
    const bool quads = (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0 && C <= 8;
    const int grid = loss_grid(quads ? B * inner / 4 : B * inner);
    if (quads && C <= 4)
      region_stats_mc_vec4_kernel<4, true><<<grid, kLT, 0, s>>>(a);
    else if (quads)
      region_stats_mc_vec4_kernel<8, false><<<grid, kLT, 0, s>>>(a);
    else if (C <= 4)
      region_stats_mc_small_kernel<4><<<grid, kLT, 0, s>>>(a);
    else if (C <= 8)
      region_stats_mc_small_kernel<8><<<grid, kLT, 0, s>>>(a);
    else
      region_stats_mc_kernel<<<grid, kLT, 3 * C * sizeof(double), s>>>(a);
    region_stats_mc_finalize<<<C, kLT, 0, s>>>(a.partials, grid, C, stats_out);
  } else {
    TB_REQUIRE(t_kind == TB_T_DENSE_F32, TB_ERR_ARG, "tb_dice_stats: binary/multilabel need a dense f32 target");
    const bool vec4 = (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0;
    const long long innerv = inner / (vec4 ? 4 : 1);
    const int chunks = (int)std::max<long long>(1, std::min<long long>(64, (innerv + kLT * 4 - 1) / (kLT * 4)));
    a.chunks = chunks;
    TB_REQUIRE((long long)B * C * chunks < (1ll << 31), TB_ERR_UNSUPPORTED, "tb_dice_stats: too many (batch, class) rows");
    const unsigned grid = (unsigned)((long long)B * C * chunks);
    if (vec4)
      region_stats_elem_kernel<4><<<grid, kLT, 0, s>>>(a);
    else
      region_stats_elem_kernel<1><<<grid, kLT, 0, s>>>(a);
    region_stats_elem_finalize<<<C, kLT, 0, s>>>(a.partials, B, C, chunks, stats_out);
  }
  TB_LAUNCH_CHECK("region_stats_kernel");
  return TB_OK;
}

int tb_dice_loss_grad(const float* x, const void* t, int t_kind, int64_t B, int C, int64_t inner,
                      const tb_dice_params* p, const double* stats, const int32_t* classes, int n_classes,
                      float* loss_out, float* grad, void* stream) {
  TB_REQUIRE(x && t && p && stats && loss_out, TB_ERR_ARG, "tb_dice_loss_grad: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_dice_loss_grad: bad shape");
  TB_REQUIRE(C <= 4096, TB_ERR_UNSUPPORTED, "tb_dice_loss_grad: supports C <= 4096");
  TB_REQUIRE(!classes || n_classes > 0, TB_ERR_ARG, "tb_dice_loss_grad: empty class list");
  cudaStream_t s = (cudaStream_t)stream;
  RegionArgs a{};
  a.x = x; a.t = t; a.t_kind = t_kind; a.B = B; a.inner = inner; a.C = C; a.p = *p;
  a.stats = stats; a.classes = classes; a.n_classes = n_classes; a.grad = grad; a.loss_out = loss_out;
  int grid = 1;
  bool vec4 = false, mc_quads = false;
  if (grad) {
    if (p->mode == 1) {
      mc_quads = p->from_logits && C <= 8 && (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(t) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
      grid = loss_grid(mc_quads ? B * inner / 4 : B * inner);
    } else {
      vec4 = (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
      grid = loss_grid(B * C * inner / (vec4 ? 4 : 1));
    }
  }
  if (mc_quads && C <= 4)
    region_grad_kernel<12><<<grid, kLT, 2 * C * sizeof(float), s>>>(a);
  else if (mc_quads)
    region_grad_kernel<8><<<grid, kLT, 2 * C * sizeof(float), s>>>(a);
  else if (vec4)
    region_grad_kernel<4><<<grid, kLT, 2 * C * sizeof(float), s>>>(a);
  else
    region_grad_kernel<1><<<grid, kLT, 2 * C * sizeof(float), s>>>(a);
  TB_LAUNCH_CHECK("region_grad_kernel");
  return TB_OK;
}

int tb_focal_dice_fwd_bwd(const float* x, const float* t, int64_t B, int C, int64_t inner, const tb_focal_params* fp,
                          const float* class_w, const tb_dice_params* dp, const int32_t* classes, int n_classes,
                          float w_focal, float w_dice, float* loss_out, float* grad, void* ws, size_t ws_bytes,
                          void* stream) {
  TB_REQUIRE(x && t && fp && dp && loss_out && ws, TB_ERR_ARG, "tb_focal_dice_fwd_bwd: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_focal_dice_fwd_bwd: bad shape");
  TB_REQUIRE(!fp->softmax && !fp->normalized && (fp->reduction == TB_LOSS_MEAN || fp->reduction == TB_LOSS_SUM),
             TB_ERR_UNSUPPORTED, "tb_focal_dice_fwd_bwd: focal term must be sigmoid, not normalised, mean or sum reduced");
  TB_REQUIRE((dp->mode == 0 || dp->mode == 2) && dp->from_logits, TB_ERR_UNSUPPORTED,
             "tb_focal_dice_fwd_bwd: region term must be binary / multilabel on logits");
  TB_REQUIRE(C <= 4096, TB_ERR_UNSUPPORTED, "tb_focal_dice_fwd_bwd: C <= 4096");
  TB_REQUIRE(!classes || n_classes > 0, TB_ERR_ARG, "tb_focal_dice_fwd_bwd: empty class list");
  TB_REQUIRE(ws_bytes >= tb_loss_workspace_bytes(B, C, inner), TB_ERR_WORKSPACE, "tb_focal_dice_fwd_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  JointArgs a{};
  a.x = x; a.t = t; a.B = B; a.inner = inner; a.C = C;
  a.f.p = *fp; a.f.cw = class_w; a.d = *dp; a.classes = classes; a.n_classes = n_classes;
  a.w_focal = w_focal; a.w_dice = w_dice;
  a.gscale = fp->reduction == TB_LOSS_MEAN ? (float)(1.0 / ((double)B * C * inner)) : 1.f;
  a.grad = grad; a.loss_out = loss_out;
  const bool vec4 = (inner % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0 &&
                    (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0);
  const long long innerv = inner / (vec4 ? 4 : 1);
  const int chunks = (int)std::max<long long>(1, std::min<long long>(64, (innerv + kLT * 4 - 1) / (kLT * 4)));
  a.chunks = chunks;
  a.partials = static_cast<double*>(ws);
  a.stats = a.partials + (size_t)B * C * chunks * 4;
  TB_REQUIRE((long long)B * C * chunks < (1ll << 31), TB_ERR_UNSUPPORTED, "tb_focal_dice_fwd_bwd: too many (batch, class) rows");
  const unsigned grid1 = (unsigned)((long long)B * C * chunks);
  // the common configuration (BinaryFocalLoss() + DiceLoss / JaccardLoss without ignore_index) takes the branch-free kernels
  const bool fast = vec4 && fp->gamma == 2.f && !fp->has_alpha && !class_w && !fp->has_ignore && !fp->has_reduced_threshold &&
                    !dp->has_ignore;
  if (fast)
    joint_stats_kernel<4, true><<<grid1, kLT, 0, s>>>(a);
  else if (vec4)
    joint_stats_kernel<4, false><<<grid1, kLT, 0, s>>>(a);
  else
    joint_stats_kernel<1, false><<<grid1, kLT, 0, s>>>(a);
  joint_finalize_kernel<<<C, kLT, 0, s>>>(a.partials, B, C, chunks, a.stats);
  const int grid2 = grad ? loss_grid(B * C * inner / (vec4 ? 4 : 1)) : 1;
  if (fast)
    joint_grad_kernel<4, true><<<grid2, kLT, 2 * C * sizeof(float), s>>>(a);
  else if (vec4)
    joint_grad_kernel<4, false><<<grid2, kLT, 2 * C * sizeof(float), s>>>(a);
  else
    joint_grad_kernel<1, false><<<grid2, kLT, 2 * C * sizeof(float), s>>>(a);
  TB_LAUNCH_CHECK("joint_focal_dice_kernel");
  return TB_OK;
}

int tb_wing_fwd_bwd(const float* x, const float* t, int64_t n, float width, float curvature, int reduction,
                    float* loss_out, float* grad, void* ws, size_t ws_bytes, void* stream) {
  TB_REQUIRE(x && t && loss_out && ws, TB_ERR_ARG, "tb_wing_fwd_bwd: null pointer");
  TB_REQUIRE(n > 0, TB_ERR_ARG, "tb_wing_fwd_bwd: empty input");
  TB_REQUIRE(reduction == TB_LOSS_NONE || reduction == TB_LOSS_MEAN || reduction == TB_LOSS_SUM, TB_ERR_ARG,
             "tb_wing_fwd_bwd: bad reduction %d", reduction);
  TB_REQUIRE(ws_bytes >= tb_loss_workspace_bytes(1, 1, n), TB_ERR_WORKSPACE, "tb_wing_fwd_bwd: workspace too small");
  TB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0 &&
                 (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0),
             TB_ERR_ALIGN, "tb_wing_fwd_bwd: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  WingArgs a;
  a.x = x; a.t = t; a.n = n; a.width = width; a.curvature = curvature; a.inv_curvature = 1.f / curvature;
  a.cconst = (float)((double)width - (double)width * log(1.0 + (double)width / (double)curvature));
  a.gscale = reduction == TB_LOSS_MEAN ? (float)(1.0 / (double)n) : 1.f;
  a.loss_elem = reduction == TB_LOSS_NONE ? loss_out : nullptr;
  a.grad = grad;
  a.partials = static_cast<double*>(ws);
  const int grid = loss_grid(std::max<long long>(1, n / 4));
  wing_kernel<<<grid, kLT, 0, s>>>(a);
  if (reduction != TB_LOSS_NONE)
    scalar_finalize_kernel<<<1, kLT, 0, s>>>(a.partials, grid, reduction == TB_LOSS_MEAN ? 1.0 / (double)n : 1.0, loss_out);
  TB_LAUNCH_CHECK("wing_kernel");
  return TB_OK;
}

int tb_scale_by_device_scalar(float* data, int64_t n, const float* scale_dev, void* stream) {
  TB_REQUIRE(data && scale_dev && n >= 0, TB_ERR_ARG, "tb_scale_by_device_scalar: bad argument");
  if (n == 0) return TB_OK;
  scale_by_device_scalar_kernel<<<loss_grid(n), kLT, 0, (cudaStream_t)stream>>>(data, n, scale_dev);
  TB_LAUNCH_CHECK("scale_by_device_scalar_kernel");
  return TB_OK;
}

}  // extern "C"
