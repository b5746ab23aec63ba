// Single-view (no TTA) slice and merge kernels for sm_100a -- BASELINE configs 1 and 5.
//
// With one view there is nothing to permute and nothing to share between threads, so the TMA ring of the multi-view
// merge only adds a producer lane (index arithmetic for every 16 KB it feeds) and barrier round trips: ncu on cfg 5
// (profiles/r2_cfg5_ncu_summary.md) shows that kernel at 27 % DRAM throughput with the consumers spinning on mbarriers.
// Here every thread streams its own data with 128-bit loads, and the memory system is kept busy by occupancy:
//
//   merge   one CTA = a 32-row x 128-column canvas unit; lane = 4 consecutive pixels, warp = 128 consecutive pixels of
//           one row (512 B per request), 4 row passes per thread. The covering tile rows / columns are the same for the
//           whole unit (tile size and stride are multiples of 32 / 128), computed once per CTA. Up to four covering tiles
//           are loaded (prediction: streaming LDG.128; weight: L2-resident table or separable in-register gaussian)
//           before any is consumed; tiles are accumulated in the reference's tile order with separately rounded mul / add
//           and an IEEE division, so the result is bit-identical to the TMA / generic kernels and to TileMerger.
//   slice   lane = 4 consecutive pixels of a tile row, 8 rows in flight per thread (loads first, then the streaming
//           float4 stores); the grid is sized to >= 8 waves of resident CTAs so the tail wave costs little.
//
// Reference: inference/tiles.py:177-204 (split), :321-350,271-280 (integrate_batch + merge + crop).
#include <stdlib.h>

#include <algorithm>

#include "tiles_args.cuh"

namespace tb {

constexpr int kSV = 256;  // threads per CTA

__device__ __forceinline__ void ldcs4(const float* p, float v[4]) {
  const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

// ---------------------------------------------------------------------------------------------------------------
// merge
// ---------------------------------------------------------------------------------------------------------------
// UW = unit width in pixels (128: Tw and Sx multiples of 128; 32 otherwise).
// WK = how the tile weight is obtained: 0 table (L2-resident), 1 separable gaussian (column factors in registers), 2 weight_at()
//      per pixel, 3 separable gaussian with the unit's 1-D factors in shared memory (default for gaussian weights).
// PLAIN: mean / sum reduction and no pre-activation (one view: the value itself): no per-element switch.
// gridDim.y > 1: one row pass per CTA (small problems: 4x the CTAs, a quarter of the dependent load rounds per CTA).
//
// The covering tiles of a unit are the same for all of its pixels (lattice geometry), so thread 0 lists them once per CTA
// in shared memory -- base pointer of the tile's first plane, canvas origin -- in the reference's accumulation order; the
// inner loop is then: address = base + k*plane + (y - ty)*Tw + (x - tx), no per-thread coverage test, no index arithmetic.
constexpr int kSvMaxTiles = 16;  // covering tiles per unit the list holds (4x4); larger covering sets -> general kernels

template <int UW, int WK, bool PLAIN, int MINB = 4, int CH = 4>
__global__ void __launch_bounds__(kSV, MINB) merge_single_view_kernel(MergeAllArgs a, const __grid_constant__ BatchPtrs ptrs,
                                                                     int nux, int ux0) {
  constexpr int LPR = UW / 4;        // lanes per unit row
  constexpr int RPP = kSV / LPR;     // rows per pass
  constexpr int PASSES = 32 / RPP;
  __shared__ const float* s_base[kSvMaxTiles];
  __shared__ int s_ty[kSvMaxTiles], s_tx[kSvMaxTiles];
  __shared__ int s_nt, s_cls;
  const tb_grid& g = a.grid;
  const int uy = blockIdx.x / nux, ux = blockIdx.x - uy * nux;
  const int Y0 = a.y0 - (a.y0 % 32) + uy * 32;
  const int X0 = (ux0 + ux) * UW;
  const size_t plane = (size_t)g.Th * g.Tw;
  if (threadIdx.x == 0) {
    int ra = (Y0 - g.Th + g.Sy) / g.Sy;
    if (Y0 - g.Th + 1 <= 0) ra = 0;
    const int rb_all = min((Y0 + 31) / g.Sy, g.ny - 1);
    int ca = (X0 - g.Tw + g.Sx) / g.Sx;
    if (X0 - g.Tw + 1 <= 0) ca = 0;
    const int cb = min((X0 + UW - 1) / g.Sx, g.nx - 1);
    int cls = 0;
    bool skip = false;
    if (a.peer.phase_mask) {  // multi-GPU ragged strips: is this block ours, seeded, a spill? (all covering tiles count)
      cls = peer_block_class(a.peer, g.nx, ra, rb_all, ca, cb);
      skip = cls == 0;
    }
    int nt = 0;
    if (!skip) {
      const int r_lo = max(ra, a.r0), r_hi = min(rb_all, a.r1 - 1);
      for (int r = r_lo; r <= r_hi; ++r)
        for (int c = ca; c <= cb; ++c) {
          const int rel = r * g.nx + c - a.first_tile;
          if (rel < 0 || rel >= a.covered_tiles) continue;  // tile not resident (band mode / ragged strips)
          int bi = 0, pos = rel;
          if (a.batch_tiles < a.covered_tiles) {
            bi = rel / a.batch_tiles;
            pos = rel - bi * a.batch_tiles;
          }
          if (nt < kSvMaxTiles) {
            s_base[nt] = ptrs.p[bi] + (size_t)pos * a.K * plane;
            s_ty[nt] = r * g.Sy;
            s_tx[nt] = c * g.Sx;
          }
          ++nt;
        }
    }
    s_nt = skip ? -1 : nt;
    s_cls = cls;
  }
  __syncthreads();
  const int nt = s_nt, cls = s_cls;
  if (nt < 0) return;
  // WK == 3: separable gaussian with the 1-D factors of the unit's covering tiles in shared memory -- per listed tile 32 row
  // factors exp(-di^2 c) and UW column factors exp(-dj^2 c), computed once per CTA by all threads (<= 3 ex2 each); the inner
  // loop then costs one LDS.32 + one LDS.128 + 4 FMUL per tile instead of an L2 round trip (table) or 16 live registers of
  // column factors + one ex2 per tile (WK == 1, which pushed the 64-register instance into spills).
  __shared__ float s_wi[WK == 3 ? 4 : 1][WK == 3 ? 32 : 1];
  __shared__ __align__(16) float s_wj[WK == 3 ? 4 : 1][WK == 3 ? UW : 1];
  if constexpr (WK == 3) {
    const int ntl = min(nt, 4);
    for (int idx = threadIdx.x; idx < ntl * (32 + UW); idx += kSV) {
      const int u = idx / (32 + UW), r = idx - u * (32 + UW);
      if (r < 32) {
        const float di = (float)(Y0 + r - s_ty[u]) - a.w.hc;
        s_wi[u][r] = __expf(-(di * di) * a.w.param);
      } else {
        const float dj = (float)(X0 + (r - 32) - s_tx[u]) - a.w.wc;
        s_wj[u][r - 32] = __expf(-(dj * dj) * a.w.param);
      }
    }
    __syncthreads();
  }
  const bool seeded = cls & (TB_PH_SEEDED_FINAL | TB_PH_SEEDED_SPILL), spill = cls & (TB_PH_SPILL | TB_PH_SEEDED_SPILL);
  const int lc = threadIdx.x % LPR, row0 = threadIdx.x / LPR;
  const int x = X0 + 4 * lc;
  if (x >= a.Wp) return;  // (after the last barrier)
  const int rows = a.y1 - a.y0;

  // separable gaussian: w(i,j) = exp(-di^2 c) * exp(-dj^2 c); the column factors of the first four listed tiles live in
  // registers (static indexing only: a runtime index would push the array to local memory)
  float wj[4][4];
  if (WK == 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tx = u < nt ? s_tx[u] : 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dj = (float)(x - tx + e) - a.w.wc;
        wj[u][e] = __expf(-(dj * dj) * a.w.param);
      }
    }
  }

  // row pass outside, channel inside: the running argmax of the label epilogue then lives in one LabelState
  const int p_begin = gridDim.y > 1 ? (int)blockIdx.y : 0, p_end = gridDim.y > 1 ? p_begin + 1 : PASSES;
#pragma unroll 1
  for (int p = p_begin; p < p_end; ++p) {
    const int y = Y0 + row0 + p * RPP;
    if (y < a.y0 || y >= a.y1) continue;
    LabelState ls;
#pragma unroll 1
    for (int k = 0; k < a.K; ++k) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.init_num) {
        const size_t o = (size_t)(y - a.y0) * a.Wp + x;
        ldcs4(a.init_num + (size_t)k * rows * a.Wp + o, acc);
        ldcs4(a.init_den + o, den);
      }
      if (seeded) {
        const size_t o = (size_t)(y - a.peer.seed_y0) * a.Wp + x;
        const float4 n4 = __ldcg(reinterpret_cast<const float4*>(a.peer.seed_num + (size_t)k * a.peer.seed_rows * a.Wp + o));
        const float4 d4 = __ldcg(reinterpret_cast<const float4*>(a.peer.seed_den + o));
        acc[0] = n4.x; acc[1] = n4.y; acc[2] = n4.z; acc[3] = n4.w;
        den[0] = d4.x; den[1] = d4.y; den[2] = d4.z; den[3] = d4.w;
      }
#pragma unroll 1
      for (int u0 = 0; u0 < nt; u0 += CH) {  // CH tiles' loads in flight before any is consumed
        float pv[CH][4], wv[CH][4];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          if (u0 + u < nt) {
            const int i = y - s_ty[u0 + u], j = x - s_tx[u0 + u];
            const int off = i * g.Tw + j;
            ldcs4(s_base[u0 + u] + (size_t)k * plane + off, pv[u]);
            if (WK == 0) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w.table + off));
              wv[u][0] = w4.x; wv[u][1] = w4.y; wv[u][2] = w4.z; wv[u][3] = w4.w;
            } else if (WK == 3) {  // (only launched with nt <= 4)
              const float wi = s_wi[u0 + u][y - Y0];
              const float4 w4 = *reinterpret_cast<const float4*>(&s_wj[u0 + u][4 * lc]);
              wv[u][0] = wi * w4.x; wv[u][1] = wi * w4.y; wv[u][2] = wi * w4.z; wv[u][3] = wi * w4.w;
            } else if (WK == 1) {
              const float di = (float)i - a.w.hc;
              const float wi = __expf(-(di * di) * a.w.param);
#pragma unroll
              for (int e = 0; e < 4; ++e) {  // (WK == 1 is only launched with nt <= 4; static index: u0 is 0 or CH)
                float wjc = wj[u][e];
                if (CH == 2 && u0 == 2) wjc = wj[(u + 2) & 3][e];
                wv[u][e] = wi * wjc;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) wv[u][e] = weight_at<float>(a.w, i, j + e);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          if (u0 + u < nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float val = PLAIN ? pv[u][e]
                                      : red_post<float>(a.red, red_pre<float>(a.red, pre_act(a.pre, a.temperature, pv[u][e])), 1);
              acc[e] = add_rn(acc[e], mul_rn(val, wv[u][e]));
              den[e] = add_rn(den[e], wv[u][e]);
            }
          }
        }
      }
      if (spill) {
        const size_t o = (size_t)(y - a.peer.spill_y0) * a.Wp + x;
        *reinterpret_cast<float4*>(a.peer.spill_num + (size_t)k * a.peer.spill_rows * a.Wp + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (k == 0) *reinterpret_cast<float4*>(a.peer.spill_den + o) = make_float4(den[0], den[1], den[2], den[3]);
      } else if (a.raw) {
        const size_t o = (size_t)(y - a.y0) * a.Wp + x;
        *reinterpret_cast<float4*>(a.out_num + (size_t)k * rows * a.Wp + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (k == 0) *reinterpret_cast<float4*>(a.out_den + o) = make_float4(den[0], den[1], den[2], den[3]);
      } else {
        const float q[4] = {acc[0] / den[0], acc[1] / den[1], acc[2] / den[2], acc[3] / den[3]};
        merge_epilogue(a, k, y - a.top, x - a.left, q, ls);
      }
    }
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

cudaError_t launch_merge_single_view(const MergeAllArgs& a, const BatchPtrs& ptrs, cudaStream_t s) {
  const tb_grid& g = a.grid;
  if (a.views.count != 1) return cudaErrorNotSupported;
  if (a.peer.fused) return cudaErrorInvalidValue;  // in-kernel flag handling needs the persistent (ordered) TMA kernel
  if ((g.Th % 32) || (g.Tw % 32) || (g.Sx % 32) || (g.Sy % 32) || (a.Wp % 4)) return cudaErrorNotSupported;
  if (!a.vec_ok && a.out_mode == TB_OUT_LOGITS && !a.raw) return cudaErrorNotSupported;
  if (a.w.kind == TB_W_TABLE && !al16(a.w.table)) return cudaErrorNotSupported;
  if (a.init_num && (!al16(a.init_num) || !al16(a.init_den))) return cudaErrorNotSupported;
  if (a.raw && (!al16(a.out_num) || !al16(a.out_den))) return cudaErrorNotSupported;
  const bool wide = (g.Tw % 128 == 0) && (g.Sx % 128 == 0);
  const int uw = wide ? 128 : 32;
  const int cover = ((g.Tw + g.Sx - 1) / g.Sx) * ((g.Th + g.Sy - 1) / g.Sy);
  if (cover > kSvMaxTiles) return cudaErrorNotSupported;
  static const int gauss_variant = getenv("TB200_SV_GAUSS") ? atoi(getenv("TB200_SV_GAUSS")) : 3;  // 3: smem factors, 1: registers
  const int wk = a.w.kind == TB_W_TABLE ? 0 : ((a.w.kind == TB_W_GAUSSIAN && cover <= 4) ? (gauss_variant == 1 ? 1 : 3) : 2);
  const bool plain = (a.red == TB_RED_MEAN || a.red == TB_RED_SUM) && a.pre == TB_PRE_NONE;
  const int xa = a.raw ? 0 : a.left, xb = a.raw ? a.Wp : a.left + a.W;
  const int ux0 = xa / uw, nux = (xb + uw - 1) / uw - ux0;
  const long long units = (long long)nux * a.nby;
  if (units <= 0 || units >= (1ll << 31)) return cudaErrorNotSupported;
  // few units (small images): one row pass per CTA, so that the grid still fills the GPU several times over
  const int passes = wide ? 4 : 1;
  const bool split = passes > 1 && units < (long long)sm_count() * 3 * 8;
  const dim3 grid((unsigned)units, split ? passes : 1);
  // A/B switch on the register budget / loads in flight of the common instances: TB200_SV_MERGE = 10 * CH + resident CTAs
  // per SM (default 44: 4 tiles in flight, 4 CTAs/SM = 64 registers; measured on cfg 5, profiles/r2_single_view_variants.md)
  const char* mv = getenv("TB200_SV_MERGE");
  const int var = mv ? atoi(mv) : 44;
  if (wide && plain && (wk <= 1 || wk == 3) && var != 44) {
#define TB_SV_MB(WK_, MB, CH_) merge_single_view_kernel<128, WK_, true, MB, CH_><<<grid, kSV, 0, s>>>(a, ptrs, nux, ux0)
#define TB_SV_VAR(WK_)                                                          \
    switch (var) {                                                              \
      case 43: TB_SV_MB(WK_, 3, 4); break;                                      \
      case 24: TB_SV_MB(WK_, 4, 2); break;                                      \
      case 25: TB_SV_MB(WK_, 5, 2); break;                                      \
      case 26: TB_SV_MB(WK_, 6, 2); break;                                      \
      default: TB_SV_MB(WK_, 4, 4); break;                                      \
    }
    if (wk == 0) { TB_SV_VAR(0) } else if (wk == 3) { TB_SV_VAR(3) } else { TB_SV_VAR(1) }
#undef TB_SV_VAR
#undef TB_SV_MB
    return cudaGetLastError();
  }
#define TB_SV(UW, WK_, PLAIN_) merge_single_view_kernel<UW, WK_, PLAIN_><<<grid, kSV, 0, s>>>(a, ptrs, nux, ux0)
#define TB_SV_WK(UW, PLAIN_) do { if (wk == 0) TB_SV(UW, 0, PLAIN_); else if (wk == 1) TB_SV(UW, 1, PLAIN_); else if (wk == 3) TB_SV(UW, 3, PLAIN_); else TB_SV(UW, 2, PLAIN_); } while (0)
  if (wide) {
    if (plain) TB_SV_WK(128, true); else TB_SV_WK(128, false);
  } else {
    if (plain) TB_SV_WK(32, true); else TB_SV_WK(32, false);
  }
#undef TB_SV_WK
#undef TB_SV
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// slice
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sv_to_f32(unsigned char v) { return (float)v; }
__device__ __forceinline__ float sv_to_f32(unsigned short v) { return (float)v; }
__device__ __forceinline__ float sv_to_f32(short v) { return (float)v; }
__device__ __forceinline__ float sv_to_f32(float v) { return v; }

constexpr int kRowsInFlight = 8;

// grid (bands, n): a thread keeps a 4-pixel column group of the tile and walks down the band's rows kRowsInFlight at a time
template <typename SrcT>
__global__ void __launch_bounds__(kSV) slice_single_view_kernel(SliceArgs a) {
  const int groups = a.Tw >> 2;            // host guarantees Tw % 4 == 0
  const int gpt = min(groups, kSV);        // column groups side by side
  const int rsub = threadIdx.x / gpt, rstep = kSV / gpt;
  if (rsub >= rstep) return;
  const int t = blockIdx.y;
  const SrcT* img = static_cast<const SrcT*>(a.img);
  const size_t plane = (size_t)a.Th * a.Tw;
  const int x0 = __ldg(a.xy + 2 * t), y0 = __ldg(a.xy + 2 * t + 1);
  const int rows_per_band = (a.Th + gridDim.x - 1) / gridDim.x;
  const int i_begin = blockIdx.x * rows_per_band, i_end = min(a.Th, i_begin + rows_per_band);
  float* out_tile = a.out + (size_t)t * a.C * plane;
  for (int cg = threadIdx.x % gpt; cg < groups; cg += gpt) {
    const int j = cg << 2;
    int sx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sx[e] = border_remap(x0 + j + e, 0, a.W, a.border);
    for (int i0 = i_begin + rsub; i0 < i_end; i0 += kRowsInFlight * rstep) {
      for (int m = 0; m < kRowsInFlight; ++m) {
        const int i = i0 + m * rstep;
        if (i >= i_end) break;
        const int sy = border_remap(y0 + i, 0, a.H, a.border);
        float* o = out_tile + (size_t)i * a.Tw + j;
        for (int c = 0; c < a.C; ++c) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (sy < 0 || sx[e] < 0)
              v[e] = c < 4 ? a.bval[c] : 0.f;
            else if (a.layout == TB_HWC)
              v[e] = sv_to_f32(img[((size_t)sy * a.W + sx[e]) * a.C + c]);
            else
              v[e] = sv_to_f32(img[((size_t)c * a.H + sy) * a.W + sx[e]]);
            if (a.has_affine && c < 8) v[e] = __fadd_rn(__fmul_rn(v[e], a.scale[c]), a.bias[c]);
          }
          __stcs(reinterpret_cast<float4*>(o + (size_t)c * plane), make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    }
  }
}

// u8, one channel (cfg 5): 4 pixels = one 32-bit load -> one streaming float4 store. RIF rows in flight per thread; PIPE: the
// loads of batch n+1 are issued before batch n is stored; MINB: resident CTAs per SM the register budget is cut for.
template <int RIF, bool PIPE, int MINB>
__global__ void __launch_bounds__(kSV, MINB) slice_u8c1_kernel(SliceArgs a) {
  const int groups = a.Tw >> 2;
  const int gpt = min(groups, kSV);
  const int rsub = threadIdx.x / gpt, rstep = kSV / gpt;
  if (rsub >= rstep) return;
  const int t = blockIdx.y;
  const unsigned char* img = static_cast<const unsigned char*>(a.img);
  const int x0 = __ldg(a.xy + 2 * t), y0 = __ldg(a.xy + 2 * t + 1);
  const int rows_per_band = (a.Th + gridDim.x - 1) / gridDim.x;
  const int i_begin = blockIdx.x * rows_per_band, i_end = min(a.Th, i_begin + rows_per_band);
  float* out_tile = a.out + (size_t)t * a.Th * a.Tw;
  const bool tile_words = ((a.W & 3) == 0) && ((x0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.img) & 3) == 0);
  const int step = RIF * rstep;
  for (int cg = threadIdx.x % gpt; cg < groups; cg += gpt) {
    const int j = cg << 2;
    int sx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sx[e] = border_remap(x0 + j + e, 0, a.W, a.border);
    const bool word_ok = tile_words && sx[0] >= 0 && sx[3] == sx[0] + 3;
    float* o_col = out_tile + j;
    auto load_batch = [&](int ib, uchar4 (&q)[RIF], unsigned& fillmask) {
      fillmask = 0;
#pragma unroll
      for (int m = 0; m < RIF; ++m) {
        const int i = ib + m * rstep;
        const int sy = i < i_end ? border_remap(y0 + i, 0, a.H, a.border) : -1;
        q[m] = make_uchar4(0, 0, 0, 0);
        if (sy < 0) {
          fillmask |= 1u << m;
        } else if (word_ok) {
          q[m] = __ldg(reinterpret_cast<const uchar4*>(img + (size_t)sy * a.W + sx[0]));
        } else {  // tile column not 4-aligned in the image, or the 4-group crosses the border: byte loads / fill value
          const unsigned char* row = img + (size_t)sy * a.W;
          const unsigned char fv = (unsigned char)0;
          q[m].x = sx[0] >= 0 ? row[sx[0]] : fv; q[m].y = sx[1] >= 0 ? row[sx[1]] : fv;
          q[m].z = sx[2] >= 0 ? row[sx[2]] : fv; q[m].w = sx[3] >= 0 ? row[sx[3]] : fv;
          if (sx[0] < 0 || sx[1] < 0 || sx[2] < 0 || sx[3] < 0) fillmask |= 0x10000u << m;  // per-pixel fill needed
        }
      }
    };
    auto store_batch = [&](int ib, const uchar4 (&q)[RIF], unsigned fillmask) {
#pragma unroll
      for (int m = 0; m < RIF; ++m) {
        const int i = ib + m * rstep;
        if (i < i_end) {
          float v[4] = {(float)q[m].x, (float)q[m].y, (float)q[m].z, (float)q[m].w};
          if (fillmask & (1u << m)) v[0] = v[1] = v[2] = v[3] = a.bval[0];
          if (fillmask & (0x10000u << m)) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (sx[e] < 0) v[e] = a.bval[0];
          }
          if (a.has_affine) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __fadd_rn(__fmul_rn(v[e], a.scale[0]), a.bias[0]);
          }
          __stcs(reinterpret_cast<float4*>(o_col + (size_t)i * a.Tw), make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    };
    int i0 = i_begin + rsub;
    if (PIPE) {
      uchar4 qa[RIF], qb[RIF];
      unsigned fa, fb;
      load_batch(i0, qa, fa);
      while (i0 < i_end) {
        load_batch(i0 + step, qb, fb);
        store_batch(i0, qa, fa);
        i0 += step;
        if (i0 >= i_end) break;
        load_batch(i0 + step, qa, fa);
        store_batch(i0, qb, fb);
        i0 += step;
      }
    } else {
      for (; i0 < i_end; i0 += step) {
        uchar4 q[RIF];
        unsigned f;
        load_batch(i0, q, f);
        store_batch(i0, q, f);
      }
    }
  }
}

static int sv_slice_variant() {
  const char* v = getenv("TB200_SV_SLICE");
  return v ? atoi(v) : 8;  // measured on cfg 5 (profiles/r2_single_view_variants.md): occupancy beats rows in flight
}

cudaError_t launch_slice_single_view(const SliceArgs& a, int img_dtype, cudaStream_t s) {
  if (a.views.count != 1 || (a.Tw % 4) || !al16(a.out) || a.n > 65535 || a.n <= 0) return cudaErrorNotSupported;
  // enough CTAs for >= 8 waves of the ~5-8 resident CTAs per SM (a tail wave then costs <= ~1/8 of a wave), but bands no
  // shorter than kRowsInFlight rows per row-lane
  const int groups = a.Tw / 4, gpt = std::min(groups, kSV), rstep = kSV / gpt;
  const int min_rows = kRowsInFlight * rstep;
  const int want = (sm_count() * 8 * 8 + a.n - 1) / a.n;
  const int bands = std::max(1, std::min(std::max(1, a.Th / min_rows), want));
  dim3 grid((unsigned)bands, (unsigned)a.n);
  if (img_dtype == TB_U8 && a.C == 1) {
    switch (sv_slice_variant()) {  // A/B switch (TB200_SV_SLICE); default = the fastest measured on cfg 5
      case 1: slice_u8c1_kernel<8, false, 4><<<grid, kSV, 0, s>>>(a); break;
      case 6: slice_u8c1_kernel<2, false, 8><<<grid, kSV, 0, s>>>(a); break;
      case 7: slice_u8c1_kernel<4, false, 8><<<grid, kSV, 0, s>>>(a); break;
      case 2: slice_u8c1_kernel<8, false, 6><<<grid, kSV, 0, s>>>(a); break;
      case 3: slice_u8c1_kernel<16, false, 3><<<grid, kSV, 0, s>>>(a); break;
      case 4: slice_u8c1_kernel<4, true, 6><<<grid, kSV, 0, s>>>(a); break;
      case 5: slice_u8c1_kernel<8, true, 4><<<grid, kSV, 0, s>>>(a); break;
      case 0: slice_u8c1_kernel<4, false, 6><<<grid, kSV, 0, s>>>(a); break;
      default: slice_u8c1_kernel<1, false, 8><<<grid, kSV, 0, s>>>(a); break;
    }
    return cudaGetLastError();
  }
  switch (img_dtype) {
    case TB_U8: slice_single_view_kernel<unsigned char><<<grid, kSV, 0, s>>>(a); break;
    case TB_U16: slice_single_view_kernel<unsigned short><<<grid, kSV, 0, s>>>(a); break;
    case TB_I16: slice_single_view_kernel<short><<<grid, kSV, 0, s>>>(a); break;
    case TB_F32: slice_single_view_kernel<float><<<grid, kSV, 0, s>>>(a); break;
    default: return cudaErrorNotSupported;  // f16: generic kernel
  }
  return cudaGetLastError();
}

}  // namespace tb
