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
// UW = unit width in pixels (128: Tw and Sx multiples of 128; 32 otherwise). SEP: separable in-register gaussian.
// PLAIN: mean / sum reduction and no pre-activation (one view: the value itself): no per-element switch.
// gridDim.y > 1: one row pass per CTA (small problems: 4x the CTAs, a quarter of the dependent load rounds per CTA).
template <int UW, bool SEP, bool PLAIN>
__global__ void __launch_bounds__(kSV, 3) merge_single_view_kernel(MergeAllArgs a, const __grid_constant__ BatchPtrs ptrs,
                                                                  int nux, int ux0) {
  constexpr int LPR = UW / 4;        // lanes per unit row
  constexpr int RPP = kSV / LPR;     // rows per pass
  constexpr int PASSES = 32 / RPP;
  __shared__ int sh[4];
  const tb_grid& g = a.grid;
  const int uy = blockIdx.x / nux, ux = blockIdx.x - uy * nux;
  const int Y0 = a.y0 - (a.y0 % 32) + uy * 32;
  const int X0 = (ux0 + ux) * UW;
  if (threadIdx.x == 0) {
    int ra = (Y0 - g.Th + g.Sy) / g.Sy;
    if (Y0 - g.Th + 1 <= 0) ra = 0;
    int rb = (Y0 + 31) / g.Sy;
    sh[0] = max(ra, a.r0);
    sh[1] = min(rb, min(g.ny, a.r1) - 1);
    int ca = (X0 - g.Tw + g.Sx) / g.Sx;
    if (X0 - g.Tw + 1 <= 0) ca = 0;
    sh[2] = ca;
    sh[3] = min((X0 + UW - 1) / g.Sx, g.nx - 1);
  }
  __syncthreads();
  const int ra = sh[0], rb = sh[1], ca = sh[2], cb = sh[3];
  int cls = 0;
  if (a.peer.phase_mask) {  // multi-GPU ragged strips: is this block ours, seeded, a spill? (uniform over the CTA)
    cls = peer_block_class(a.peer, g.nx, ra, rb, ca, cb);
    if (!cls) return;
  }
  const bool seeded = cls & (TB_PH_SEEDED_FINAL | TB_PH_SEEDED_SPILL), spill = cls & (TB_PH_SPILL | TB_PH_SEEDED_SPILL);
  const int lc = threadIdx.x % LPR, row0 = threadIdx.x / LPR;
  const int x = X0 + 4 * lc;
  if (x >= a.Wp) return;  // (after the only barrier)
  const int rows = a.y1 - a.y0;
  const size_t plane = (size_t)g.Th * g.Tw;
  const bool one_batch = a.batch_tiles >= a.covered_tiles;

  // separable gaussian: w(i,j) = exp(-di^2 c) * exp(-dj^2 c); the column factors of the (<= 2) covering tile columns
  // live in registers (selected by compare, never indexed: a runtime index would push the array to local memory)
  float wj0[4] = {0.f, 0.f, 0.f, 0.f}, wj1[4] = {0.f, 0.f, 0.f, 0.f};
  if (SEP) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d0 = (float)(x - ca * g.Sx + e) - a.w.wc;
      const float d1 = (float)(x - (ca + 1) * g.Sx + e) - a.w.wc;
      wj0[e] = __expf(-(d0 * d0) * a.w.param);
      wj1[e] = __expf(-(d1 * d1) * a.w.param);
    }
  }

  // row pass outside, channel inside: the running argmax of the label epilogue then lives in one LabelState
  const int p_begin = gridDim.y > 1 ? (int)blockIdx.y : 0, p_end = gridDim.y > 1 ? p_begin + 1 : PASSES;
#pragma unroll 1
  for (int p = p_begin; p < p_end; ++p) {
    const int y = Y0 + row0 + p * RPP;
    if (y < a.y0 || y >= a.y1) continue;
    LabelState ls;
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
      int r = ra, c = ca;
      while (r <= rb) {
        float pv[4][4], wv[4][4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ok[u] = false;
          if (r <= rb) {
            const int rel = r * g.nx + c - a.first_tile;
            const int i = y - r * g.Sy, j = x - c * g.Sx;
            if (rel >= 0 && rel < a.covered_tiles && i >= 0 && i < g.Th && j >= 0 && j + 3 < g.Tw) {
              int bi = 0, pos = rel;
              if (!one_batch) {
                bi = rel / a.batch_tiles;
                pos = rel - bi * a.batch_tiles;
              }
              ok[u] = true;
              ldcs4(ptrs.p[bi] + ((size_t)pos * a.K + k) * plane + (size_t)i * g.Tw + j, pv[u]);
              if (SEP) {
                const float di = (float)i - a.w.hc;
                const float wi = __expf(-(di * di) * a.w.param);
                const bool first_col = (c == ca);  // <= 2 covering columns on this path (host checks)
#pragma unroll
                for (int e = 0; e < 4; ++e) wv[u][e] = wi * (first_col ? wj0[e] : wj1[e]);
              } else if (a.w.kind == TB_W_TABLE) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w.table + (size_t)i * g.Tw + j));
                wv[u][0] = w4.x; wv[u][1] = w4.y; wv[u][2] = w4.z; wv[u][3] = w4.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) wv[u][e] = weight_at<float>(a.w, i, j + e);
              }
            }
            if (++c > cb) {
              c = ca;
              ++r;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (ok[u]) {
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
  if ((g.Th % 32) || (g.Tw % 32) || (g.Sx % 32) || (g.Sy % 32) || (a.Wp % 4)) return cudaErrorNotSupported;
  if (!a.vec_ok && a.out_mode == TB_OUT_LOGITS && !a.raw) return cudaErrorNotSupported;
  if (a.w.kind == TB_W_TABLE && !al16(a.w.table)) return cudaErrorNotSupported;
  if (a.init_num && (!al16(a.init_num) || !al16(a.init_den))) return cudaErrorNotSupported;
  if (a.raw && (!al16(a.out_num) || !al16(a.out_den))) return cudaErrorNotSupported;
  const bool wide = (g.Tw % 128 == 0) && (g.Sx % 128 == 0);
  const int uw = wide ? 128 : 32;
  const int cover_cols = (g.Tw + g.Sx - 1) / g.Sx;
  const bool sep = a.w.kind == TB_W_GAUSSIAN && cover_cols <= 2;
  const bool plain = (a.red == TB_RED_MEAN || a.red == TB_RED_SUM) && a.pre == TB_PRE_NONE;
  const int xa = a.raw ? 0 : a.left, xb = a.raw ? a.Wp : a.left + a.W;
  const int ux0 = xa / uw, nux = (xb + uw - 1) / uw - ux0;
  const long long units = (long long)nux * a.nby;
  if (units <= 0 || units >= (1ll << 31)) return cudaErrorNotSupported;
  // few units (small images): one row pass per CTA, so that the grid still fills the GPU several times over
  const int passes = wide ? 4 : 1;
  const bool split = passes > 1 && units < (long long)sm_count() * 4 * 8;
  const dim3 grid((unsigned)units, split ? passes : 1);
#define TB_SV(UW, SEP_, PLAIN_) merge_single_view_kernel<UW, SEP_, PLAIN_><<<grid, kSV, 0, s>>>(a, ptrs, nux, ux0)
  if (wide) {
    if (sep) { if (plain) TB_SV(128, true, true); else TB_SV(128, true, false); }
    else { if (plain) TB_SV(128, false, true); else TB_SV(128, false, false); }
  } else {
    if (sep) { if (plain) TB_SV(32, true, true); else TB_SV(32, true, false); }
    else { if (plain) TB_SV(32, false, true); else TB_SV(32, false, false); }
  }
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
  // 4 consecutive bytes of one row as a single 32-bit load: u8, one channel, the 4-group inside the image and 4-aligned
  const bool tile_words = sizeof(SrcT) == 1 && a.C == 1 && ((a.W & 3) == 0) && ((x0 & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(a.img) & 3) == 0);
  for (int cg = threadIdx.x % gpt; cg < groups; cg += gpt) {
    const int j = cg << 2;
    int sx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sx[e] = border_remap(x0 + j + e, 0, a.W, a.border);
    const bool word_ok = tile_words && sx[0] >= 0 && sx[3] == sx[0] + 3;
    for (int i0 = i_begin + rsub; i0 < i_end; i0 += kRowsInFlight * rstep) {
      if (word_ok) {
        uchar4 q[kRowsInFlight];
        bool fill[kRowsInFlight];
#pragma unroll
        for (int m = 0; m < kRowsInFlight; ++m) {
          const int i = i0 + m * rstep;
          const int sy = i < i_end ? border_remap(y0 + i, 0, a.H, a.border) : -1;
          fill[m] = sy < 0;
          q[m] = make_uchar4(0, 0, 0, 0);
          if (!fill[m])
            q[m] = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(img) + (size_t)sy * a.W + sx[0]));
        }
#pragma unroll
        for (int m = 0; m < kRowsInFlight; ++m) {
          const int i = i0 + m * rstep;
          if (i >= i_end) break;
          float v[4] = {(float)q[m].x, (float)q[m].y, (float)q[m].z, (float)q[m].w};
          if (fill[m]) v[0] = v[1] = v[2] = v[3] = a.bval[0];
          if (a.has_affine) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __fadd_rn(__fmul_rn(v[e], a.scale[0]), a.bias[0]);
          }
          __stcs(reinterpret_cast<float4*>(out_tile + (size_t)i * a.Tw + j), make_float4(v[0], v[1], v[2], v[3]));
        }
        continue;
      }
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

cudaError_t launch_slice_single_view(const SliceArgs& a, int img_dtype, cudaStream_t s) {
  if (a.views.count != 1 || (a.Tw % 4) || !al16(a.out) || a.n > 65535 || a.n <= 0) return cudaErrorNotSupported;
  // enough CTAs for >= 8 waves of the ~5-8 resident CTAs per SM (a tail wave then costs <= ~1/8 of a wave), but bands no
  // shorter than kRowsInFlight rows per row-lane
  const int groups = a.Tw / 4, gpt = std::min(groups, kSV), rstep = kSV / gpt;
  const int min_rows = kRowsInFlight * rstep;
  const int want = (sm_count() * 8 * 8 + a.n - 1) / a.n;
  const int bands = std::max(1, std::min(std::max(1, a.Th / min_rows), want));
  dim3 grid((unsigned)bands, (unsigned)a.n);
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
