// 4-pixels-per-thread variants of the channel-coupled loss kernels (softmax focal, multiclass dice / jaccard) for C <= 8
// and 16-byte aligned [B,C,inner] logits with inner % 4 == 0: every channel of a pixel quad is ONE 128-bit load / store,
// the channel softmax of the four pixels lives in registers, address arithmetic is paid once per quad.
// ncu on the one-pixel-per-thread kernels (profiles/r1_final_ncu_summary.md): 757 M instructions for 134 M elements,
// issue-active 73 %, DRAM 17 % -- issue-bound by per-element index math and recomputed exponentials.
// Included by losses.cu after PixelFocalArgs / RegionArgs.
#pragma once

namespace tb {

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void ld_labels4(const long long* p, long long (&lab)[4]) {
  const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(p));
  const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(p) + 1);
  lab[0] = a.x; lab[1] = a.y; lab[2] = b.x; lab[3] = b.y;
}

// channel softmax of 4 pixels in place: xv[c][e] -> probabilities
template <int CB>
__device__ __forceinline__ void softmax4(float (&xv)[CB][4], int C) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < C) mx = fmaxf(mx, xv[c][e]);
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < C) {
        xv[c][e] = __expf(xv[c][e] - mx);
        den += xv[c][e];
      }
    const float inv = 1.f / den;
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < C) xv[c][e] *= inv;
  }
}

// ---- softmax_focal_loss_with_logits (functional.py:110-173), not normalised, no reduced threshold ----------------
// G2: gamma == 2 and no class weights (the defaults): pow() and its register footprint are compiled out (ncu on the
// generic instance: 112 registers -> 2 CTAs/SM, 24 % of the warps resident, DRAM 33 %), MINB CTAs per SM are asked for.
template <int CB, bool G2, int MINB>
__global__ void __launch_bounds__(kLT, MINB) softmax_focal_vec4_kernel(PixelFocalArgs a) {
  const long long innerq = a.inner >> 2, nquads = a.B * innerq;
  const long long* labels = static_cast<const long long*>(a.t);
  const float gamma = a.f.p.gamma;
  double acc[2] = {0.0, 0.0};
  for (long long q = (long long)blockIdx.x * kLT + threadIdx.x; q < nquads; q += (long long)gridDim.x * kLT) {
    const long long b = q / innerq, pos = (q - b * innerq) << 2;
    const float* xp = a.x + b * a.C * a.inner + pos;
    float xv[CB][4], pr[CB][4];
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < a.C) {
        ld4(xp + (long long)c * a.inner, xv[c]);
#pragma unroll
        for (int e = 0; e < 4; ++e) pr[c][e] = xv[c][e];
      }
    long long lab[4];
    ld_labels4(labels + b * a.inner + pos, lab);
    softmax4<CB>(pr, a.C);
    float Lq[4], SG[4] = {0.f, 0.f, 0.f, 0.f};
    float Fsum = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool valid = lab[e] != a.ignore_label;
      const int label = valid ? (int)lab[e] : 0;  // masked_fill(ignore, 0) then one_hot (:141-146)
      float Lpix = 0.f;
#pragma unroll
      for (int c = 0; c < CB; ++c)
        if (c < a.C) {
          const float x = xv[c][e], p = pr[c][e];
          const float oh = (c == label) ? 1.f : 0.f;
          const float pt = (c == label) ? 1.f - p : p;
          const float focal = G2 ? pt * pt : powg(pt, gamma);
          const float dfocal_dp = (G2 ? 2.f * pt : dpowg(pt, gamma)) * (1.f - 2.f * oh);
          const float ex = __expf(-fabsf(x));
          const float r = __fdividef(1.f, 1.f + ex);
          const float sig = x >= 0.f ? r : ex * r;
          const float bce = fmaxf(x, 0.f) - x * oh + __logf(1.f + ex);
          const float w = (!G2 && a.f.cw) ? __ldg(a.f.cw + c) : 1.f;
          Lpix += focal * bce * w;
          Fsum += focal;
          const float G = valid ? w * bce * dfocal_dp : 0.f;           // d Lpix / d prob_c
          const float direct = valid ? w * focal * (sig - oh) : 0.f;   // through the BCE term
          SG[e] += G * p;
          xv[c][e] = p * G + direct;  // dL_c = p (G - S_G) + direct = (p G + direct) - p S_G
        }
      Lq[e] = valid ? Lpix : 0.f;
    }
    acc[0] += (double)(Lq[0] + Lq[1] + Lq[2] + Lq[3]);
    acc[1] += (double)Fsum;
    if (a.loss_elem) __stcs(reinterpret_cast<float4*>(a.loss_elem + b * a.inner + pos), make_float4(Lq[0], Lq[1], Lq[2], Lq[3]));
    if (a.grad) {
      float* gp = a.grad + b * a.C * a.inner + pos;
#pragma unroll
      for (int c = 0; c < CB; ++c)
        if (c < a.C) {
          float g[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) g[e] = (xv[c][e] - pr[c][e] * SG[e]) * a.gscale;
          __stcs(reinterpret_cast<float4*>(gp + (long long)c * a.inner), make_float4(g[0], g[1], g[2], g[3]));
        }
    }
  }
  block_reduce<2>(acc);
  if (threadIdx.x == 0) {
    a.partials[2 * blockIdx.x] = acc[0];
    a.partials[2 * blockIdx.x + 1] = acc[1];
  }
}

// ---- multiclass dice / jaccard statistics (dice.py:90-99, jaccard.py:76-82) ---------------------------------------
// Per-thread float partials over 16 quads (64 pixels), then into double accumulators. SMEM_ACC (CB <= 4): the 3*CB doubles
// of a thread live in shared memory ([3*CB][kLT], conflict-free) instead of 6*CB registers -- the register version sat at
// 92 registers = 2 CTAs/SM (ncu: 24 % of the warps resident, DRAM 40 %); a flush is 3*CB LDS/DADD/STS per 16 quads.
template <int CB, bool SMEM_ACC>
__global__ void __launch_bounds__(kLT, SMEM_ACC ? 4 : 1) region_stats_mc_vec4_kernel(RegionArgs a) {
  const long long innerq = a.inner >> 2, nquads = a.B * innerq;
  const long long* labels = static_cast<const long long*>(a.t);
  __shared__ double sh_acc[SMEM_ACC ? 3 * CB : 1][SMEM_ACC ? kLT : 1];
  float fI[CB], fS[CB], fT[CB];
  double dI[SMEM_ACC ? 1 : CB], dS[SMEM_ACC ? 1 : CB], dT[SMEM_ACC ? 1 : CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    fI[c] = fS[c] = fT[c] = 0.f;
    if constexpr (SMEM_ACC) {
      sh_acc[3 * c][threadIdx.x] = 0.0;
      sh_acc[(3 * c + 1)][threadIdx.x] = 0.0;
      sh_acc[(3 * c + 2)][threadIdx.x] = 0.0;
    } else {
      dI[c] = dS[c] = dT[c] = 0.0;
    }
  }
  int since_flush = 0;
  for (long long q = (long long)blockIdx.x * kLT + threadIdx.x; q < nquads; q += (long long)gridDim.x * kLT) {
    const long long b = q / innerq, pos = (q - b * innerq) << 2;
    const float* xp = a.x + b * a.C * a.inner + pos;
    float pr[CB][4];
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < a.C) ld4(xp + (long long)c * a.inner, pr[c]);
    long long lab[4];
    ld_labels4(labels + b * a.inner + pos, lab);
    if (a.p.from_logits) softmax4<CB>(pr, a.C);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool valid = !(a.p.has_ignore && (float)lab[e] == a.p.ignore_index);
#pragma unroll
      for (int c = 0; c < CB; ++c)
        if (c < a.C) {
          const float p = valid ? pr[c][e] : 0.f;
          const float t = (valid && lab[e] == c) ? 1.f : 0.f;
          fI[c] += p * t;
          fS[c] += p + t;
          fT[c] += t;
        }
    }
    if (++since_flush == 16) {  // 64 pixels per float partial, then into the doubles
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        if constexpr (SMEM_ACC) {
          sh_acc[3 * c][threadIdx.x] += (double)fI[c];
          sh_acc[(3 * c + 1)][threadIdx.x] += (double)fS[c];
          sh_acc[(3 * c + 2)][threadIdx.x] += (double)fT[c];
        } else {
          dI[c] += fI[c]; dS[c] += fS[c]; dT[c] += fT[c];
        }
        fI[c] = fS[c] = fT[c] = 0.f;
      }
      since_flush = 0;
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    double v[3];
    if constexpr (SMEM_ACC) {
      v[0] = sh_acc[3 * c][threadIdx.x] + (double)fI[c];
      v[1] = sh_acc[(3 * c + 1)][threadIdx.x] + (double)fS[c];
      v[2] = sh_acc[(3 * c + 2)][threadIdx.x] + (double)fT[c];
    } else {
      v[0] = dI[c] + fI[c]; v[1] = dS[c] + fS[c]; v[2] = dT[c] + fT[c];
    }
    block_reduce<3>(v);
    if (threadIdx.x == 0 && c < a.C) {
      double* o = a.partials + ((size_t)blockIdx.x * a.C + c) * 3;
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
  }
}

// ---- multiclass dice / jaccard gradient: dL/dx_c = p_c (g_c - sum_k g_k p_k), g_c = alpha_c [label == c] + beta_c --------
// coef: shared [C][2] (alpha, beta) prepared by the caller kernel prologue (region_grad_kernel's)
template <int CB>
__device__ __forceinline__ void region_grad_mc_vec4(const RegionArgs& a, const float* sh_coef) {
  const long long innerq = a.inner >> 2, nquads = a.B * innerq;
  const long long* labels = static_cast<const long long*>(a.t);
  float al[CB], be[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    al[c] = c < a.C ? sh_coef[2 * c] : 0.f;
    be[c] = c < a.C ? sh_coef[2 * c + 1] : 0.f;
  }
  for (long long q = (long long)blockIdx.x * kLT + threadIdx.x; q < nquads; q += (long long)gridDim.x * kLT) {
    const long long b = q / innerq, pos = (q - b * innerq) << 2;
    const float* xp = a.x + b * a.C * a.inner + pos;
    float* gp = a.grad + b * a.C * a.inner + pos;
    float pr[CB][4];
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < a.C) ld4(xp + (long long)c * a.inner, pr[c]);
    long long lab[4];
    ld_labels4(labels + b * a.inner + pos, lab);
    softmax4<CB>(pr, a.C);
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    bool valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      valid[e] = !(a.p.has_ignore && (float)lab[e] == a.p.ignore_index);
#pragma unroll
      for (int c = 0; c < CB; ++c)
        if (c < a.C) dot[e] += (al[c] * (lab[e] == c ? 1.f : 0.f) + be[c]) * pr[c][e];
    }
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < a.C) {
        float g[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float gc = al[c] * (lab[e] == c ? 1.f : 0.f) + be[c];
          g[e] = valid[e] ? pr[c][e] * (gc - dot[e]) : 0.f;
        }
        __stcs(reinterpret_cast<float4*>(gp + (long long)c * a.inner), make_float4(g[0], g[1], g[2], g[3]));
      }
  }
}

}  // namespace tb
