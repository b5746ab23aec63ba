// Lovasz hinge / Lovasz-softmax as ONE launch sequence over all (image, class) segments -- sm_100a.
//
// Reference semantics (file:line relative to /root/reference/pytorch_toolbelt/losses/lovasz.py):
//   :75-86,143-160  flatten + drop ignored pixels        -> build kernel (ignored pixels get the sentinel key, no gather)
//   :65-68,125-137  errors, descending sort, gt[perm]    -> build kernel + segmented LSD radix sort (this file)
//   :23-34          lovasz_grad (Jaccard increments)     -> tile count / tile scan / apply kernels
//   :71,137         dot(relu?(errors_sorted), grad)      -> apply kernel (+ finalize)
//   :45-49,101-109,131-140  per-image / per-class loops, "present" classes, means -> segment weights in the finalize and
//                                                            apply kernels (no host loop, no host synchronisation)
//   backward                                             -> the apply kernel scatters d loss / d input through the
//                                                            sorted permutation (payload) in the same pass
//
// Sort: keys are the errors mapped to order-preserving uint32 (descending), payload = fg<<31 | neg<<30 | index. 4 passes of
// 8 bits, each pass = per-tile digit histogram -> per-(segment,digit) offsets -> stable scatter (warp match_any ranking,
// CTA-local staging through shared memory so that global writes are digit-contiguous runs). 20 B moved per key and pass
// (4 B histogram read + 8 B read + 8 B write) against the 16 B of an ideal pass.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace tb {

constexpr int kRT = 256;                        // threads per CTA
constexpr int kRadixKPT = 16;                   // keys per thread
constexpr int kRadixTile = kRT * kRadixKPT;     // 4096 keys per CTA
constexpr int kLovItems = 8;                    // apply: consecutive sorted elements per thread
constexpr int kLovTile = kRT * kLovItems;       // 2048
constexpr uint32_t kIdxMask = 0x3FFFFFFFu;      // payload bits [0,30): element index inside the segment
constexpr uint32_t kSentinelKey = 0xFFFFFFFFu;  // sorts after every finite error (descending order)

// float -> uint32 whose ASCENDING order is the DESCENDING order of the floats (and back)
__device__ __forceinline__ uint32_t desc_key(float e) {
  const uint32_t b = __float_as_uint(e);
  const uint32_t u = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
  return ~u;
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  const uint32_t u = ~k;
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

struct LovaszArgs2 {
  const float* x;      // hinge: logits [B,inner]; softmax: probabilities [B,C,inner]
  const void* labels;  // [B,inner] f32 or i64
  int t_kind;
  long long B, inner, P;  // P = elements per segment: inner (per_image) or B*inner
  int C, mode, per_image, has_ignore;
  float ignore;
  int groups, segs, T, TA;  // T = radix tiles per segment, TA = apply tiles per segment
  uint32_t *keys, *pay;     // [segs][P] (the buffers that hold the SORTED result)
  unsigned long long* n_fg;     // [segs]
  unsigned long long* n_valid;  // [groups]
  uint32_t* tile_counts;        // [segs][TA]
  double* partials;             // [segs][TA]
  float* grad;                  // same layout as x, or null
  float* loss_out;
};

__device__ __forceinline__ float label_value(const LovaszArgs2& a, long long i) {
  return a.t_kind == TB_T_LABELS_I64 ? (float)__ldg(static_cast<const long long*>(a.labels) + i)
                                     : __ldg(static_cast<const float*>(a.labels) + i);
}

// offset of element j of segment (g, c) in the input / gradient tensor
__device__ __forceinline__ long long input_offset(const LovaszArgs2& a, int g, int c, long long j, long long& lab_off) {
  long long b, pix;
  if (a.per_image) {
    b = g;
    pix = j;
  } else {
    b = j / a.inner;
    pix = j - b * a.inner;
  }
  lab_off = b * a.inner + pix;
  return (b * a.C + c) * a.inner + pix;
}

// ---- 1. errors -> keys / payload, positives and valid counts -------------------------------------------------------
__global__ void __launch_bounds__(kRT) lovasz_build_kernel(LovaszArgs2 a) {
  const int seg = blockIdx.y, g = seg / a.C, c = seg - g * a.C;
  const long long base = (long long)blockIdx.x * kRadixTile;
  const size_t sbase = (size_t)seg * a.P;
  unsigned nfg = 0, nvalid = 0;
#pragma unroll 4
  for (int k = 0; k < kRadixKPT; ++k) {
    const long long j = base + (long long)k * kRT + threadIdx.x;
    if (j >= a.P) break;
    long long lo;
    const long long off = input_offset(a, g, c, j, lo);
    const float y = label_value(a, lo);
    const bool valid = !(a.has_ignore && y == a.ignore);
    const float xv = __ldcs(a.x + off);
    float e;
    bool fg, neg = false;
    if (a.mode == 0) {  // hinge: e = 1 - x * (2y - 1)   (:65-66)
      fg = valid && (y != 0.f);
      e = 1.f - xv * (2.f * y - 1.f);
    } else {  // softmax: e = |fg - p_c|   (:125,133)
      fg = valid && (y == (float)c);
      const float d = (fg ? 1.f : 0.f) - xv;
      neg = d < 0.f;
      e = fabsf(d);
    }
    a.keys[sbase + j] = valid ? desc_key(e) : kSentinelKey;
    a.pay[sbase + j] = (fg ? 0x80000000u : 0u) | (neg ? 0x40000000u : 0u) | (uint32_t)j;
    nfg += fg ? 1u : 0u;
    nvalid += valid ? 1u : 0u;
  }
  // integer block reduction -> one atomic per CTA (exact, order-independent)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nfg += __shfl_xor_sync(0xffffffffu, nfg, o);
    nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
  }
  __shared__ unsigned sh[2][kRT / 32];
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = nfg;
    sh[1][threadIdx.x >> 5] = nvalid;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned f = 0, v = 0;
    for (int w = 0; w < kRT / 32; ++w) {
      f += sh[0][w];
      v += sh[1][w];
    }
    if (f) atomicAdd(a.n_fg + seg, (unsigned long long)f);
    if (v && c == 0) atomicAdd(a.n_valid + g, (unsigned long long)v);
  }
}

// ---- 2. segmented LSD radix sort (one pass = hist + offsets + scatter) ----------------------------------------------
__global__ void __launch_bounds__(kRT) radix_hist_kernel(const uint32_t* __restrict__ keys, long long P, int shift, int T,
                                                         uint32_t* __restrict__ counts, uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int seg = blockIdx.y;
  const long long base = (long long)blockIdx.x * kRadixTile;
  const uint32_t* k = keys + (size_t)seg * P;
#pragma unroll 4
  for (int it = 0; it < kRadixKPT; ++it) {
    const long long i = base + (long long)it * kRT + threadIdx.x;
    if (i < P) atomicAdd(&h[(__ldg(k + i) >> shift) & 255u], 1u);
  }
  __syncthreads();
  const uint32_t c = h[threadIdx.x];
  counts[((size_t)seg * 256 + threadIdx.x) * T + blockIdx.x] = c;
  if (c) atomicAdd(&digit_total[seg * 256 + threadIdx.x], c);
}

// inclusive scan of one value per thread over the CTA (256 threads); returns the inclusive prefix, *total = CTA sum
__device__ __forceinline__ uint32_t block_scan_incl(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[kRT / 32];
  __shared__ uint32_t tot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = lane < kRT / 32 ? wsum[lane] : 0u;
#pragma unroll
    for (int o = 1; o < kRT / 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < kRT / 32) wsum[lane] = s;  // inclusive over warps
    if (lane == kRT / 32 - 1) tot = s;
  }
  __syncthreads();
  const uint32_t res = x + (warp ? wsum[warp - 1] : 0u);
  *total = tot;
  __syncthreads();  // wsum / tot are reused by the next call
  return res;
}

// grid (256 digits, segs): counts[seg][d][0..T) -> exclusive global offsets (digit base + prefix over tiles), in place
__global__ void __launch_bounds__(kRT) radix_offsets_kernel(uint32_t* __restrict__ counts, const uint32_t* __restrict__ digit_total,
                                                            int T) {
  const int d = blockIdx.x, seg = blockIdx.y;
  uint32_t below = 0;
  (void)block_scan_incl(threadIdx.x < d ? digit_total[seg * 256 + threadIdx.x] : 0u, &below);
  uint32_t* row = counts + ((size_t)seg * 256 + d) * T;
  uint32_t carry = below;
  for (int i0 = 0; i0 < T; i0 += kRT) {
    const int i = i0 + threadIdx.x;
    const uint32_t v = i < T ? row[i] : 0u;
    uint32_t total;
    const uint32_t incl = block_scan_incl(v, &total);
    if (i < T) row[i] = carry + incl - v;
    carry += total;
  }
}

__global__ void __launch_bounds__(kRT) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ pay_in,
                                                            uint32_t* __restrict__ keys_out, uint32_t* __restrict__ pay_out,
                                                            long long P, int shift, int T, const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t whist[kRT / 32][256];  // per-warp digit counts, then per-warp exclusive base inside the CTA's digit run
  __shared__ uint32_t dstart[256];           // CTA-local start of each digit run
  __shared__ uint32_t gbase[256];            // global start of this CTA's run of each digit
  __shared__ uint32_t sk[kRadixTile], sp[kRadixTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.y;
  const size_t sbase = (size_t)seg * P;
  const long long base = (long long)blockIdx.x * kRadixTile;
  for (int i = threadIdx.x; i < (kRT / 32) * 256; i += kRT) (&whist[0][0])[i] = 0u;
  gbase[threadIdx.x] = offsets[((size_t)seg * 256 + threadIdx.x) * T + blockIdx.x];
  __syncthreads();

  // element order inside the tile = (warp, iteration, lane): monotone in the global index, so ranks are stable
  uint32_t key[kRadixKPT], pay[kRadixKPT];
  uint16_t rank[kRadixKPT];
  const long long wbase = base + (long long)warp * (32 * kRadixKPT);
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kRadixKPT; ++it) {
    const long long i = wbase + it * 32 + lane;
    const bool in = i < P;
    key[it] = in ? __ldg(keys_in + sbase + i) : kSentinelKey;  // padding ranks last (digit 255, end of the last tile)
    pay[it] = in ? __ldg(pay_in + sbase + i) : 0u;
    const uint32_t d = (key[it] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t r = __popc(peers & lt_mask);
    const uint32_t prev = whist[warp][d];
    __syncwarp();
    if (r == 0) whist[warp][d] = prev + __popc(peers);
    __syncwarp();
    rank[it] = (uint16_t)(prev + r);
  }
  __syncthreads();
  uint32_t total_d = 0;
  {
    const int d = threadIdx.x;
#pragma unroll
    for (int w = 0; w < kRT / 32; ++w) {
      const uint32_t c = whist[w][d];
      whist[w][d] = total_d;
      total_d += c;
    }
  }
  uint32_t tile_total;
  const uint32_t incl = block_scan_incl(total_d, &tile_total);  // (contains the barriers that publish whist)
  dstart[threadIdx.x] = incl - total_d;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kRadixKPT; ++it) {
    const uint32_t d = (key[it] >> shift) & 255u;
    const uint32_t lp = dstart[d] + whist[warp][d] + rank[it];
    sk[lp] = key[it];
    sp[lp] = pay[it];
  }
  __syncthreads();
  const long long left = P - base;
  const int n_in = (int)(left < kRadixTile ? left : kRadixTile);
#pragma unroll 4
  for (int k = 0; k < kRadixKPT; ++k) {
    const int lp = threadIdx.x + k * kRT;
    if (lp < n_in) {
      const uint32_t kk = sk[lp];
      const uint32_t d = (kk >> shift) & 255u;
      const size_t gpos = sbase + gbase[d] + (uint32_t)(lp - dstart[d]);
      keys_out[gpos] = kk;
      pay_out[gpos] = sp[lp];
    }
  }
}

// ---- 3. Lovasz gradient on the sorted segments --------------------------------------------------------------------
__global__ void __launch_bounds__(kRT) lovasz_tile_count_kernel(LovaszArgs2 a) {
  const int seg = blockIdx.y, g = seg / a.C;
  const long long n = (long long)a.n_valid[g];
  const long long base = (long long)blockIdx.x * kLovTile;
  const uint32_t* pay = a.pay + (size_t)seg * a.P;
  unsigned cnt = 0;
  for (int k = 0; k < kLovItems; ++k) {
    const long long i = base + (long long)k * kRT + threadIdx.x;
    if (i < n) cnt += __ldg(pay + i) >> 31;
  }
  uint32_t total;
  (void)block_scan_incl(cnt, &total);
  if (threadIdx.x == 0) a.tile_counts[(size_t)seg * a.TA + blockIdx.x] = total;
}

// one CTA per segment: exclusive scan of the tile counts, in place
__global__ void __launch_bounds__(kRT) lovasz_tile_scan_kernel(uint32_t* tile_counts, int TA) {
  uint32_t* row = tile_counts + (size_t)blockIdx.x * TA;
  uint32_t carry = 0;
  for (int i0 = 0; i0 < TA; i0 += kRT) {
    const int i = i0 + threadIdx.x;
    const uint32_t v = i < TA ? row[i] : 0u;
    uint32_t total;
    const uint32_t incl = block_scan_incl(v, &total);
    if (i < TA) row[i] = carry + incl - v;
    carry += total;
  }
}

// weight of a segment in the final loss: 1 / (number of averaged images) [/ number of present classes of its image]
__device__ __forceinline__ float segment_weight(const LovaszArgs2& a, int seg) {
  const int g = seg / a.C;
  if (a.mode == 0) return 1.f / (float)a.groups;
  if (a.n_fg[seg] == 0ull) return 0.f;  // class absent from this image / batch: skipped (:127-128)
  int present = 0;
  for (int c = 0; c < a.C; ++c) present += a.n_fg[(size_t)g * a.C + c] != 0ull ? 1 : 0;
  return 1.f / ((float)present * (float)a.groups);
}

__global__ void __launch_bounds__(kRT) lovasz_apply_kernel2(LovaszArgs2 a) {
  const int seg = blockIdx.y, g = seg / a.C, c = seg - g * a.C;
  const long long n = (long long)a.n_valid[g];
  const long long my_base = (long long)blockIdx.x * kLovTile + (long long)threadIdx.x * kLovItems;
  const uint32_t* keys = a.keys + (size_t)seg * a.P;
  const uint32_t* pay = a.pay + (size_t)seg * a.P;
  __shared__ float sh_w;
  if (threadIdx.x == 0) sh_w = segment_weight(a, seg);
  uint32_t kk[kLovItems], pp[kLovItems];
  unsigned local = 0;
#pragma unroll
  for (int k = 0; k < kLovItems; ++k) {
    const long long i = my_base + k;
    kk[k] = i < a.P ? __ldg(keys + i) : kSentinelKey;
    pp[k] = i < a.P ? __ldg(pay + i) : 0u;
    local += (i < n) ? (pp[k] >> 31) : 0u;
  }
  uint32_t tile_total;
  const uint32_t incl = block_scan_incl(local, &tile_total);  // (its barriers also publish sh_w)
  unsigned run = a.tile_counts[(size_t)seg * a.TA + blockIdx.x] + incl - local;  // positives strictly before my first element
  const float w = sh_w;
  const long long gts = (long long)a.n_fg[seg];
  double dot = 0.0;
#pragma unroll
  for (int k = 0; k < kLovItems; ++k) {
    const long long i = my_base + k;
    if (i >= a.P) break;
    float gx = 0.f;
    if (i < n) {
      const unsigned bit = pp[k] >> 31;
      // lovasz.py:27-34 at sorted position i (inclusive cumsums) minus the value at i-1. Intersection and union are exact
      // integers converted to float ONCE, with the same expression for "the value at i-1" here and "the value at i" of the
      // previous position, so the increments telescope exactly even beyond 2^24 elements (a float position counter does not:
      // (float)(i-1) + 1 != (float)i there, and the error accumulates over the whole segment)
      const long long cum = (long long)run + bit;
      float gk = 1.f - (float)(gts - cum) / (float)(gts + (i + 1) - cum);
      if (i > 0) gk -= 1.f - (float)(gts - (long long)run) / (float)(gts + i - (long long)run);
      run += bit;
      const float e = key_to_float(kk[k]);
      if (a.mode == 0) {
        dot += (double)fmaxf(e, 0.f) * (double)gk;                 // dot(relu(errors_sorted), grad)  (:71)
        gx = e > 0.f ? (bit ? -gk : gk) * w : 0.f;                 // d/dx [1 - x s] = -s, s = +1 (fg) / -1
      } else {
        dot += (double)e * (double)gk;                             // (:137)
        const float sgn = e == 0.f ? 0.f : ((pp[k] & 0x40000000u) ? -1.f : 1.f);  // sign(fg - p); d|fg-p|/dp = -sign
        gx = -sgn * gk * w;
      }
    }
    if (a.grad) {
      long long lo;
      a.grad[input_offset(a, g, c, (long long)(pp[k] & kIdxMask), lo)] = gx;
    }
  }
  // block reduction of the dot partial (double)
  __shared__ double shd[kRT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wv = 0; wv < kRT / 32; ++wv) s += shd[wv];
    a.partials[(size_t)seg * a.TA + blockIdx.x] = s;
  }
}

// loss = sum_seg weight_seg * dot_seg, fixed order (deterministic)
__global__ void __launch_bounds__(kRT) lovasz_finalize_kernel2(LovaszArgs2 a) {
  __shared__ double shd[kRT / 32];
  double total = 0.0;
  for (int seg = 0; seg < a.segs; ++seg) {
    double v = 0.0;
    for (int i = threadIdx.x; i < a.TA; i += kRT) v += a.partials[(size_t)seg * a.TA + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int wv = 0; wv < kRT / 32; ++wv) s += shd[wv];
      total += s * (double)segment_weight(a, seg);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.loss_out = (float)total;
}

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

struct LovaszLayout {
  size_t keys0, pay0, keys1, pay1, counts, digit_total, n_fg, n_valid, tile_counts, partials, total;
};

static LovaszLayout lovasz_layout(long long P, int segs, int groups) {
  LovaszLayout L;
  const size_t n = (size_t)segs * (size_t)P;
  const size_t T = (size_t)((P + kRadixTile - 1) / kRadixTile), TA = (size_t)((P + kLovTile - 1) / kLovTile);
  size_t o = 0;
  L.keys0 = o; o += up256(n * 4);
  L.pay0 = o; o += up256(n * 4);
  L.keys1 = o; o += up256(n * 4);
  L.pay1 = o; o += up256(n * 4);
  L.counts = o; o += up256((size_t)segs * 256 * T * 4);
  L.digit_total = o; o += up256((size_t)4 * segs * 256 * 4);
  L.n_fg = o; o += up256((size_t)segs * 8);
  L.n_valid = o; o += up256((size_t)groups * 8);
  L.tile_counts = o; o += up256((size_t)segs * TA * 4);
  L.partials = o; o += up256((size_t)segs * TA * 8);
  L.total = o + 256;
  return L;
}

}  // namespace tb

using namespace tb;

extern "C" {

size_t tb_lovasz_fused_workspace_bytes(int64_t B, int C, int64_t inner, int per_image) {
  if (B <= 0 || C <= 0 || inner <= 0) return 0;
  const long long P = per_image ? inner : B * inner;
  const int groups = per_image ? (int)B : 1;
  return lovasz_layout(P, groups * C, groups).total;
}

int tb_lovasz_fwd_bwd(const float* x, const void* labels, int t_kind, int64_t B, int C, int64_t inner, int mode, int per_image,
                      int has_ignore, float ignore, float* loss_out, float* grad, void* ws, size_t ws_bytes, void* stream) {
  TB_REQUIRE(x && labels && loss_out && ws, TB_ERR_ARG, "tb_lovasz_fwd_bwd: null pointer");
  TB_REQUIRE(B > 0 && C > 0 && inner > 0, TB_ERR_ARG, "tb_lovasz_fwd_bwd: bad shape");
  TB_REQUIRE(mode == 0 || mode == 1, TB_ERR_ARG, "tb_lovasz_fwd_bwd: mode must be 0 (hinge) or 1 (softmax)");
  TB_REQUIRE(mode == 1 || C == 1, TB_ERR_ARG, "tb_lovasz_fwd_bwd: the hinge variant takes one logit plane (C == 1)");
  TB_REQUIRE(t_kind == TB_T_DENSE_F32 || t_kind == TB_T_LABELS_I64, TB_ERR_ARG, "tb_lovasz_fwd_bwd: bad t_kind");
  const long long P = per_image ? inner : B * inner;
  const long long groups_ll = per_image ? B : 1;
  TB_REQUIRE(P <= (long long)kIdxMask, TB_ERR_UNSUPPORTED, "tb_lovasz_fwd_bwd: more than 2^30 elements per segment");
  TB_REQUIRE(groups_ll * C <= 65535, TB_ERR_UNSUPPORTED, "tb_lovasz_fwd_bwd: more than 65535 (image, class) segments");
  const int groups = (int)groups_ll, segs = groups * C;
  const LovaszLayout L = lovasz_layout(P, segs, groups);
  TB_REQUIRE(ws_bytes >= L.total, TB_ERR_WORKSPACE, "tb_lovasz_fwd_bwd: workspace too small (%zu < %zu)", ws_bytes, L.total);
  TB_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, TB_ERR_ALIGN, "tb_lovasz_fwd_bwd: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* base = static_cast<unsigned char*>(ws);
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(base + L.keys0), reinterpret_cast<uint32_t*>(base + L.keys1)};
  uint32_t* pay[2] = {reinterpret_cast<uint32_t*>(base + L.pay0), reinterpret_cast<uint32_t*>(base + L.pay1)};
  uint32_t* counts = reinterpret_cast<uint32_t*>(base + L.counts);
  uint32_t* digit_total = reinterpret_cast<uint32_t*>(base + L.digit_total);

  LovaszArgs2 a{};
  a.x = x; a.labels = labels; a.t_kind = t_kind; a.B = B; a.inner = inner; a.P = P; a.C = C; a.mode = mode;
  a.per_image = per_image ? 1 : 0; a.has_ignore = has_ignore ? 1 : 0; a.ignore = ignore;
  a.groups = groups; a.segs = segs;
  a.T = (int)((P + kRadixTile - 1) / kRadixTile);
  a.TA = (int)((P + kLovTile - 1) / kLovTile);
  a.keys = keys[0]; a.pay = pay[0];
  a.n_fg = reinterpret_cast<unsigned long long*>(base + L.n_fg);
  a.n_valid = reinterpret_cast<unsigned long long*>(base + L.n_valid);
  a.tile_counts = reinterpret_cast<uint32_t*>(base + L.tile_counts);
  a.partials = reinterpret_cast<double*>(base + L.partials);
  a.grad = grad; a.loss_out = loss_out;

  // counters: digit totals of the 4 passes, positives per segment, valid pixels per group (contiguous in the layout)
  TB_CUDA(cudaMemsetAsync(base + L.digit_total, 0, L.tile_counts - L.digit_total, s));
  const dim3 grid_r((unsigned)a.T, (unsigned)segs), grid_a((unsigned)a.TA, (unsigned)segs);
  lovasz_build_kernel<<<grid_r, kRT, 0, s>>>(a);
  for (int pass = 0; pass < 4; ++pass) {
    const int src = pass & 1, dst = src ^ 1, shift = 8 * pass;
    uint32_t* dt = digit_total + (size_t)pass * segs * 256;
    radix_hist_kernel<<<grid_r, kRT, 0, s>>>(keys[src], P, shift, a.T, counts, dt);
    radix_offsets_kernel<<<dim3(256, (unsigned)segs), kRT, 0, s>>>(counts, dt, a.T);
    radix_scatter_kernel<<<grid_r, kRT, 0, s>>>(keys[src], pay[src], keys[dst], pay[dst], P, shift, a.T, counts);
  }
  // 4 passes: the sorted result is back in buffer 0 (= a.keys / a.pay)
  lovasz_tile_count_kernel<<<grid_a, kRT, 0, s>>>(a);
  lovasz_tile_scan_kernel<<<segs, kRT, 0, s>>>(a.tile_counts, a.TA);
  lovasz_apply_kernel2<<<grid_a, kRT, 0, s>>>(a);
  lovasz_finalize_kernel2<<<1, kRT, 0, s>>>(a);
  TB_LAUNCH_CHECK("lovasz kernels");
  return TB_OK;
}

}  // extern "C"
