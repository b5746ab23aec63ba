// Tile crop / TTA augment / TTA de-augment / weighted merge kernels for sm_100a.
//
// Everything here is HBM-bound index-permuting copy + a few flops per element, so the design rules are:
// 128-bit coalesced global accesses on both sides, transposing D4 views staged through padded shared
// memory, each byte of the big streams (augmented tiles out, model predictions in) touched exactly once,
// grids of many small CTAs (several resident per SM) so that the memory system always has loads in flight.
//
// Reference semantics restated here (file:line relative to /root/reference/pytorch_toolbelt):
//   inference/tiles.py:144-234  crop with border          -> crop_tiles_kernel, slice_augment_kernel
//   inference/tta.py:257-484    *_image_augment           -> slice_augment_kernel, tta_augment_kernel
//   inference/tta.py:287-524    *_image_deaugment         -> tta_deaugment_kernel, deaug_* helpers
//   inference/tiles.py:310-350  TileMerger                -> deaugment_merge_kernel, normalize_crop_kernel
//   (whole loop)                                          -> merge_all_kernel (output-stationary)
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <utility>

#include "common.cuh"
#include "merge_tma.cuh"
#include "tiles_args.cuh"

namespace tb {

constexpr int kThreads = 256;
constexpr int BS = 32;        // block edge (pixels) for 4- and 8-byte elements
constexpr int PITCH = BS + 1; // shared-memory row pitch (elements): conflict-free transposed reads

// ------------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_stream4(const float* p, float v[4]) {
  const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void ld_stream4(const double* p, double v[4]) {
  const double2 a = __ldcs(reinterpret_cast<const double2*>(p));
  const double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(double* p, const double v[4]) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

__device__ __forceinline__ void st4(unsigned int* p, const unsigned int v[4]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(unsigned long long* p, const unsigned long long v[4]) {
  reinterpret_cast<ulonglong2*>(p)[0] = make_ulonglong2(v[0], v[1]);
  reinterpret_cast<ulonglong2*>(p)[1] = make_ulonglong2(v[2], v[3]);
}
__device__ __forceinline__ void st4(unsigned short* p, const unsigned short v[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) p[e] = v[e];
}
__device__ __forceinline__ void st4(unsigned char* p, const unsigned char v[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) p[e] = v[e];
}

// Vector float reduction into global memory (sm_90+): one RED.128 instead of four RED.32.
__device__ __forceinline__ void red_add4(float* p, const float v[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
               : "memory");
}
__device__ __forceinline__ void red_add1(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add1(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add4(double* p, const double v[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) atomicAdd(p + e, v[e]);
}

// ------------------------------------------------------------------------------------------------
// 1. native-dtype HWC crop (ImageSlicer.split / iter_split / cut_patch on host arrays)
// ------------------------------------------------------------------------------------------------
struct CropArgs {
  const void* img;
  void* out;
  const int32_t* xy;
  int H, W, C, n, Th, Tw, border, clip_to_tile;
  unsigned long long bval[4];
};

template <typename E>
__global__ void __launch_bounds__(kThreads) crop_tiles_kernel(CropArgs a) {
  const int t = blockIdx.y;
  const int x0 = __ldg(a.xy + 2 * t), y0 = __ldg(a.xy + 2 * t + 1);
  // valid ranges for the border remap
  int ylo = 0, yhi = a.H, xlo = 0, xhi = a.W;
  if (a.clip_to_tile) {
    ylo = max(y0, 0); yhi = min(a.H, y0 + a.Th);
    xlo = max(x0, 0); xhi = min(a.W, x0 + a.Tw);
  }
  const E* img = static_cast<const E*>(a.img);
  E* out = static_cast<E*>(a.out) + (size_t)t * a.Th * a.Tw * a.C;
  const int rowlen = a.Tw * a.C;
  const long long total = (long long)a.Th * rowlen;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int i = (int)(idx / rowlen);
    const int rem = (int)(idx - (long long)i * rowlen);
    const int j = rem / a.C, c = rem - j * a.C;
    const int sy = border_remap(y0 + i, ylo, yhi, a.border);
    const int sx = border_remap(x0 + j, xlo, xhi, a.border);
    E v;
    if (sy < 0 || sx < 0) {
      const unsigned long long bits = c < 4 ? a.bval[c] : 0ull;
      v = (E)bits;
    } else {
      v = img[((size_t)sy * a.W + sx) * a.C + c];
    }
    out[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Shared "emit all views of one source block" routine (source-block-stationary augment).
// sm holds the source block [BS][PITCH]; valid extents (vr, vc); block origin (R0, C0) inside a source
// plane of (Hs, Ws). For every view a: out_a[i,j] = src[view_map(code_a, i, j, Hs, Ws)].
// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__device__ __forceinline__ void emit_views(const T* sm, int R0, int C0, int vr, int vc, int Hs, int Ws,
                                           ViewSet views, T* out_plane0, size_t view_stride) {
  for (int a = 0; a < views.count; ++a) {
    const unsigned code = view_code(views.packed, a);
    const unsigned inv = view_inverse(code);
    const int Ho = (code & V_T) ? Ws : Hs, Wo = (code & V_T) ? Hs : Ws;
    // image of the source block under the inverse map: map both corners, take the min corner
    int ia, ja, ib, jb;
    view_map(inv, R0, C0, Ho, Wo, ia, ja);
    view_map(inv, R0 + vr - 1, C0 + vc - 1, Ho, Wo, ib, jb);
    const int oi0 = min(ia, ib), oj0 = min(ja, jb);
    const int orows = (code & V_T) ? vc : vr, ocols = (code & V_T) ? vr : vc;
    T* out = out_plane0 + (size_t)a * view_stride;
    if (VEC) {
      const int ngroups = ocols >> 2;
      for (int idx = threadIdx.x; idx < orows * ngroups; idx += kThreads) {
        const int li = idx / ngroups, lq = idx - li * ngroups;
        const int i = oi0 + li, j = oj0 + 4 * lq;
        T v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int si, sj;
          view_map(code, i, j + e, Hs, Ws, si, sj);
          v[e] = sm[(si - R0) * PITCH + (sj - C0)];
        }
        st4(out + (size_t)i * Wo + j, v);
      }
    } else {
      for (int idx = threadIdx.x; idx < orows * ocols; idx += kThreads) {
        const int li = idx / ocols, lj = idx - li * ocols;
        const int i = oi0 + li, j = oj0 + lj;
        int si, sj;
        view_map(code, i, j, Hs, Ws, si, sj);
        out[(size_t)i * Wo + j] = sm[(si - R0) * PITCH + (sj - C0)];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2. slice + HWC->CHW + cast + TTA augment
// ------------------------------------------------------------------------------------------------
// SliceArgs: tiles_args.cuh
__device__ __forceinline__ float to_f32(unsigned char v) { return (float)v; }
__device__ __forceinline__ float to_f32(unsigned short v) { return (float)v; }
__device__ __forceinline__ float to_f32(short v) { return (float)v; }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

template <typename SrcT, bool VEC>
__global__ void __launch_bounds__(kThreads) slice_augment_kernel(SliceArgs a) {
  __shared__ float sm[BS * PITCH];
  const int per_tile = a.nbx * a.nby;
  const int t = blockIdx.x / per_tile;
  const int blk = blockIdx.x - t * per_tile;
  const int by = blk / a.nbx, bx = blk - by * a.nbx;
  const int R0 = by * BS, C0 = bx * BS;
  const int vr = min(BS, a.Th - R0), vc = min(BS, a.Tw - C0);
  const int x0 = __ldg(a.xy + 2 * t), y0 = __ldg(a.xy + 2 * t + 1);
  const SrcT* img = static_cast<const SrcT*>(a.img);
  const size_t plane = (size_t)a.Th * a.Tw;
  const size_t view_stride = (size_t)a.n * a.C * plane;

  for (int c = 0; c < a.C; ++c) {
    // ---- load the source block of channel c (coalesced along x; border synthesised) ----
    for (int idx = threadIdx.x; idx < vr * vc; idx += kThreads) {
      const int lr = idx / vc, lc = idx - lr * vc;
      const int sy = border_remap(y0 + R0 + lr, 0, a.H, a.border);
      const int sx = border_remap(x0 + C0 + lc, 0, a.W, a.border);
      float v;
      if (sy < 0 || sx < 0) {
        v = c < 4 ? a.bval[c] : 0.f;
      } else if (a.layout == TB_HWC) {
        v = to_f32(img[((size_t)sy * a.W + sx) * a.C + c]);
      } else {
        v = to_f32(img[((size_t)c * a.H + sy) * a.W + sx]);
      }
      if (a.has_affine && c < 8) v = __fadd_rn(__fmul_rn(v, a.scale[c]), a.bias[c]);
      sm[lr * PITCH + lc] = v;
    }
    __syncthreads();
    emit_views<float, VEC>(sm, R0, C0, vr, vc, a.Th, a.Tw, a.views, a.out + ((size_t)t * a.C + c) * plane,
                           view_stride);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// 3. stand-alone TTA augment on [B,C,H,W] (any element width)
// ------------------------------------------------------------------------------------------------
struct AugArgs {
  const void* x;
  void* out;
  int planes;  // B*C
  int H, W, nbx, nby;
  ViewSet views;
  size_t view_stride;  // B*C*H*W elements
};

template <typename E, bool VEC>
__global__ void __launch_bounds__(kThreads) tta_augment_kernel(AugArgs a) {
  __shared__ E sm[BS * PITCH];
  const int per_plane = a.nbx * a.nby;
  const int p = blockIdx.x / per_plane;
  const int blk = blockIdx.x - p * per_plane;
  const int by = blk / a.nbx, bx = blk - by * a.nbx;
  const int R0 = by * BS, C0 = bx * BS;
  const int vr = min(BS, a.H - R0), vc = min(BS, a.W - C0);
  const E* src = static_cast<const E*>(a.x) + (size_t)p * a.H * a.W;
  for (int idx = threadIdx.x; idx < vr * vc; idx += kThreads) {
    const int lr = idx / vc, lc = idx - lr * vc;
    sm[lr * PITCH + lc] = src[(size_t)(R0 + lr) * a.W + C0 + lc];
  }
  __syncthreads();
  emit_views<E, VEC>(sm, R0, C0, vr, vc, a.H, a.W, a.views, static_cast<E*>(a.out) + (size_t)p * a.H * a.W,
                     a.view_stride);
}

// ------------------------------------------------------------------------------------------------
// De-augment gather machinery (output-block-stationary).
//
// The calling CTA owns the output block rows [i0,i0+BS) x cols [j0,j0+BS) of a de-augmented plane of
// (Ho, Wo); thread (li, lq) = (tid/8, tid%8) owns pixels (i0+li, j0+4*lq+e), e = 0..3.
// For view a the value is  src_a[view_map(g_a, i, j, Hs, Ws)]  with g_a = inverse of the augment code.
// Non-transposing views are read straight from global memory (row-contiguous, possibly reversed);
// transposing views are staged through shared memory (coalesced row reads, conflict-free column reads).
// ------------------------------------------------------------------------------------------------
template <typename T, bool ALIGNED>
__device__ __forceinline__ void gather_direct4(const T* __restrict__ plane, unsigned g, int i, int j, int Hs,
                                               int Ws, T v[4]) {
  // caller guarantees 0 <= i < Ho (= Hs for non-transposing views); pixels with j+e outside [0,Ws) give 0
  const int si = (g & V_FI) ? Hs - 1 - i : i;
  const T* row = plane + (size_t)si * Ws;
  const bool rev = (g & V_FJ) != 0;
  const int sj0 = rev ? Ws - 4 - j : j;  // first (lowest) source column of the 4-group
  if (ALIGNED && sj0 >= 0 && sj0 + 3 < Ws) {
    T q[4];
    ld_stream4(row + sj0, q);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = rev ? q[3 - e] : q[e];
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int sj = rev ? Ws - 1 - (j + e) : j + e;
      v[e] = (j + e >= 0 && sj >= 0 && sj < Ws) ? __ldcs(row + sj) : T(0);
    }
  }
}

// Stage the source block a transposing view needs for output block (i0, j0) into sm[BS][PITCH].
template <typename T, bool ALIGNED>
__device__ __forceinline__ void stage_transposed(const T* __restrict__ plane, unsigned g, int i0, int j0, int Hs,
                                                 int Ws, T* sm, int& sr0, int& sc0) {
  // (a,b) = (j,i): source row depends on j, source column on i
  sr0 = (g & V_FI) ? Hs - j0 - BS : j0;
  sc0 = (g & V_FJ) ? Ws - i0 - BS : i0;
  const int lr = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int sr = sr0 + lr, sc = sc0 + 4 * lq;
  T q[4] = {T(0), T(0), T(0), T(0)};
  if (sr >= 0 && sr < Hs) {
    const T* row = plane + (size_t)sr * Ws;
    if (ALIGNED && sc >= 0 && sc + 3 < Ws) {
      ld_stream4(row + sc, q);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (sc + e >= 0 && sc + e < Ws) q[e] = __ldcs(row + sc + e);
    }
  }
  T* dst = sm + lr * PITCH + 4 * lq;
#pragma unroll
  for (int e = 0; e < 4; ++e) dst[e] = q[e];
}

template <typename T>
__device__ __forceinline__ void read_transposed4(const T* sm, unsigned g, int i, int j, int Hs, int Ws, int sr0,
                                                 int sc0, T v[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int si, sj;
    view_map(g, i, j + e, Hs, Ws, si, sj);
    const int rr = si - sr0, cc = sj - sc0;
    v[e] = (rr >= 0 && rr < BS && cc >= 0 && cc < BS) ? sm[rr * PITCH + cc] : T(0);
  }
}

__host__ __device__ constexpr bool packed_has_transpose(unsigned packed, int count) {
  for (int a = 0; a < count; ++a)
    if ((packed >> (3 * a)) & V_T) return true;
  return false;
}
__host__ __device__ inline bool views_have_transpose(ViewSet v) { return packed_has_transpose(v.packed, v.count); }

// Compile-time view sets (same tables as make_views) so that the per-view loops unroll with static
// register indexing.
template <int TTA>
struct Views;
#define TB_VIEWS(TTA, COUNT, ...)                                     \
  template <>                                                         \
  struct Views<TTA> {                                                 \
    static constexpr int count = COUNT;                               \
    static constexpr unsigned packed = pack_views({__VA_ARGS__});     \
  };
constexpr unsigned pack_views(std::initializer_list<unsigned> codes) {
  unsigned p = 0;
  int a = 0;
  for (unsigned c : codes) p |= c << (3 * a++);
  return p;
}
TB_VIEWS(TB_TTA_NONE, 1, V_ID)
TB_VIEWS(TB_TTA_FLIPLR, 2, V_ID, V_FJ)
TB_VIEWS(TB_TTA_FLIPUD, 2, V_ID, V_FI)
TB_VIEWS(TB_TTA_FLIPS, 3, V_ID, V_FJ, V_FI)
TB_VIEWS(TB_TTA_D2, 4, V_ID, V_FJ, V_FI, V_FI | V_FJ)
TB_VIEWS(TB_TTA_D4, 8, V_ID, V_T | V_FI, V_FI | V_FJ, V_T | V_FJ, V_T, V_FJ, V_T | V_FI | V_FJ, V_FI)
#undef TB_VIEWS

// De-augment + reduce the 4 pixels of the calling thread for one (tile/batch item, channel) plane set.
//   plane0       pointer to view 0's plane; view a is plane0 + a*view_stride
//   (Hs, Ws)     plane dims (identical for all views: transposing views require square planes)
//   sm           CTA-shared staging: [nT][BS][PITCH], nT = number of transposing views (<= 4)
// Contains two __syncthreads() when the view set has transposing views: call it from uniform control flow.
template <typename T>
__device__ __forceinline__ T pre_act_t(int pre, float temperature, T v) {
  if (pre == TB_PRE_SIGMOID) return T(1) / (T(1) + exp(-(v * T(temperature))));
  return v;
}

template <typename T, bool ALIGNED, unsigned PACKED, int COUNT>
__device__ __forceinline__ void deaug_reduce4(const T* __restrict__ plane0, size_t view_stride, int red, int i0,
                                              int j0, int Hs, int Ws, T* sm, T out[4], int pre = TB_PRE_NONE,
                                              float temperature = 1.f) {
  constexpr bool HAS_T = packed_has_transpose(PACKED, COUNT);
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int i = i0 + li, j = j0 + 4 * lq;
  const bool row_ok = (i >= 0 && i < Hs);
  int sr0[4], sc0[4];
  if (HAS_T) {
    int slot = 0;
#pragma unroll
    for (int a = 0; a < COUNT; ++a) {
      const unsigned g = view_inverse((PACKED >> (3 * a)) & 7u);
      if (g & V_T) {
        stage_transposed<T, ALIGNED>(plane0 + (size_t)a * view_stride, g, i0, j0, Hs, Ws, sm + slot * BS * PITCH,
                                     sr0[slot], sc0[slot]);
        ++slot;
      }
    }
  }
  // issue the direct loads before the barrier so they overlap the staging traffic
  T direct[4][4];
  {
    int d = 0;
#pragma unroll
    for (int a = 0; a < COUNT; ++a) {
      const unsigned g = view_inverse((PACKED >> (3 * a)) & 7u);
      if (!(g & V_T)) {
        if (row_ok)
          gather_direct4<T, ALIGNED>(plane0 + (size_t)a * view_stride, g, i, j, Hs, Ws, direct[d]);
        else
          direct[d][0] = direct[d][1] = direct[d][2] = direct[d][3] = T(0);
        ++d;
      }
    }
  }
  if (HAS_T) __syncthreads();
  T acc[4] = {T(0), T(0), T(0), T(0)};
  {
    int d = 0, slot = 0;
#pragma unroll
    for (int a = 0; a < COUNT; ++a) {
      const unsigned g = view_inverse((PACKED >> (3 * a)) & 7u);
      T v[4];
      if (g & V_T) {
        read_transposed4<T>(sm + slot * BS * PITCH, g, i, j, Hs, Ws, sr0[slot], sc0[slot], v);
        ++slot;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = direct[d][e];
        ++d;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const T pv = red_pre<T>(red, pre_act_t<T>(pre, temperature, v[e]));
        acc[e] = (a == 0) ? pv : add_rn(acc[e], pv);  // sequential view-order sum (torch CPU mean/sum)
      }
    }
  }
  if (HAS_T) __syncthreads();  // staging buffer may be overwritten by the caller's next plane
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = red_post<T>(red, acc[e], COUNT);
}


// ------------------------------------------------------------------------------------------------
// 2b. fast slice + augment for tiles whose sides are multiples of 32 (every block is full):
// compile-time view codes, per-thread offsets computed once, up to 4 channels resident in shared memory
// per pass. ~10 instructions per 16-byte store instead of ~200 in the generic kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kChunkC = 4;
constexpr int CH_STRIDE = BS * PITCH + 1;  // +1: channel planes start in different banks

template <typename SrcT>
__device__ __forceinline__ void load_block_fast(const SliceArgs& a, const SrcT* __restrict__ img, int x0, int y0,
                                                int R0, int C0, int c0, int cc, float* sm) {
  const int lr = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int gy = y0 + R0 + lr, gx = x0 + C0 + 4 * lq;
  const bool interior = (y0 + R0 >= 0) && (y0 + R0 + BS <= a.H) && (x0 + C0 >= 0) && (x0 + C0 + BS <= a.W);
  const int sy = interior ? gy : border_remap(gy, 0, a.H, a.border);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int sx = interior ? gx + e : border_remap(gx + e, 0, a.W, a.border);
    const bool fill = (sy < 0 || sx < 0);
    for (int ch = 0; ch < cc; ++ch) {
      const int c = c0 + ch;
      float v;
      if (fill) {
        v = c < 4 ? a.bval[c] : 0.f;
      } else if (a.layout == TB_HWC) {
        v = to_f32(img[((size_t)sy * a.W + sx) * a.C + c]);
      } else {
        v = to_f32(img[((size_t)c * a.H + sy) * a.W + sx]);
      }
      if (a.has_affine && c < 8) v = __fadd_rn(__fmul_rn(v, a.scale[c]), a.bias[c]);
      sm[ch * CH_STRIDE + lr * PITCH + 4 * lq + e] = v;
    }
  }
}

// VAR bit0: ask for 6 resident CTAs per SM (register cap 40); bit1: streaming (evict-first) stores
template <typename SrcT, int TTA, int VAR>
__global__ void __launch_bounds__(kThreads, (VAR & 1) ? 6 : 1) slice_augment_fast_kernel(SliceArgs a) {
  using V = Views<TTA>;
  __shared__ float sm[kChunkC * CH_STRIDE];
  const int per_tile = a.nbx * a.nby;
  const int t = blockIdx.x / per_tile;
  const int blk = blockIdx.x - t * per_tile;
  const int by = blk / a.nbx, bx = blk - by * a.nbx;
  const int R0 = by * BS, C0 = bx * BS;
  const int x0 = __ldg(a.xy + 2 * t), y0 = __ldg(a.xy + 2 * t + 1);
  const SrcT* img = static_cast<const SrcT*>(a.img);
  const size_t plane = (size_t)a.Th * a.Tw;
  const size_t view_stride = (size_t)a.n * a.C * plane;
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;

  for (int c0 = 0; c0 < a.C; c0 += kChunkC) {
    const int cc = min(kChunkC, a.C - c0);
    if (c0) __syncthreads();
    load_block_fast<SrcT>(a, img, x0, y0, R0, C0, c0, cc, sm);
    __syncthreads();
#pragma unroll
    for (int va = 0; va < V::count; ++va) {
      constexpr unsigned packed = V::packed;
      const unsigned code = (packed >> (3 * va)) & 7u;
      // shared-memory index of element e of this thread's 4-group: base + e*step
      int base, step, oi0, oj0;
      if (code & V_T) {
        // out[i,j] = src[(FI ? M-j : j), (FJ ? M-i : i)]: row varies with j, column with i
        const int col = (code & V_FJ) ? (BS - 1 - li) : li;
        const int row0 = (code & V_FI) ? (BS - 1 - 4 * lq) : 4 * lq;
        base = row0 * PITCH + col;
        step = (code & V_FI) ? -PITCH : PITCH;
        oi0 = (code & V_FJ) ? a.Tw - C0 - BS : C0;
        oj0 = (code & V_FI) ? a.Th - R0 - BS : R0;
      } else {
        const int row = (code & V_FI) ? (BS - 1 - li) : li;
        const int col0 = (code & V_FJ) ? (BS - 1 - 4 * lq) : 4 * lq;
        base = row * PITCH + col0;
        step = (code & V_FJ) ? -1 : 1;
        oi0 = (code & V_FI) ? a.Th - R0 - BS : R0;
        oj0 = (code & V_FJ) ? a.Tw - C0 - BS : C0;
      }
      float* o = a.out + (size_t)va * view_stride + ((size_t)t * a.C + c0) * plane + (size_t)(oi0 + li) * a.Tw + oj0 +
                 4 * lq;
      for (int ch = 0; ch < cc; ++ch) {
        const float* s = sm + ch * CH_STRIDE + base;
        const float4 q = make_float4(s[0], s[step], s[2 * step], s[3 * step]);
        if (VAR & 2)
          __stcs(reinterpret_cast<float4*>(o + (size_t)ch * plane), q);
        else
          *reinterpret_cast<float4*>(o + (size_t)ch * plane) = q;
      }
    }
  }
}



// ------------------------------------------------------------------------------------------------
// 2c. slice without TTA (one view): nothing is reused, so no shared memory: every thread converts 4 consecutive
//     pixels of one tile row (all channels) and writes one float4 per channel plane.
// ------------------------------------------------------------------------------------------------
template <typename SrcT>
__global__ void __launch_bounds__(kThreads) slice_plain_kernel(SliceArgs a) {
  // blockIdx.y = tile, blockIdx.x = band of rows; a thread keeps its 4 columns and walks down the rows of the band,
  // so the column border remap is computed once per thread and a row costs one remap + C*4 loads + C stores.
  const int groups = a.Tw >> 2;                 // host guarantees Tw % 4 == 0
  const int gpt = min(groups, kThreads);        // column groups handled side by side
  const int rsub = threadIdx.x / gpt, rstep = kThreads / gpt;
  const int t = blockIdx.y;
  const SrcT* img = static_cast<const SrcT*>(a.img);
  const size_t plane = (size_t)a.Th * a.Tw;
  const int x0 = __ldg(a.xy + 2 * t), y0 = __ldg(a.xy + 2 * t + 1);
  const int rows_per_block = (a.Th + gridDim.x - 1) / gridDim.x;
  const int i_begin = blockIdx.x * rows_per_block, i_end = min(a.Th, i_begin + rows_per_block);
  float* out_tile = a.out + (size_t)t * a.C * plane;
  for (int cg = threadIdx.x % gpt; cg < groups; cg += gpt) {
    const int j = cg << 2;
    int sx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sx[e] = border_remap(x0 + j + e, 0, a.W, a.border);
    if (rsub >= rstep) continue;
    for (int i = i_begin + rsub; i < i_end; i += rstep) {
      const int sy = border_remap(y0 + i, 0, a.H, a.border);
      float* o = out_tile + (size_t)i * a.Tw + j;
      if (sizeof(SrcT) == 1 && a.C == 1 && sy >= 0 && sx[0] >= 0 && sx[3] == sx[0] + 3) {
        // 4 consecutive bytes of one row: a single 32-bit load when the address allows it
        const size_t off = (size_t)sy * a.W + sx[0];
        if ((off & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) & 3) == 0)) {
          const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(img) + off));
          float v[4] = {(float)q.x, (float)q.y, (float)q.z, (float)q.w};
          if (a.has_affine) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __fadd_rn(__fmul_rn(v[e], a.scale[0]), a.bias[0]);
          }
          __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
          continue;
        }
      }
      for (int c = 0; c < a.C; ++c) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (sy < 0 || sx[e] < 0)
            v[e] = c < 4 ? a.bval[c] : 0.f;
          else if (a.layout == TB_HWC)
            v[e] = to_f32(img[((size_t)sy * a.W + sx[e]) * a.C + c]);
          else
            v[e] = to_f32(img[((size_t)c * a.H + sy) * a.W + sx[e]]);
          if (a.has_affine && c < 8) v[e] = __fadd_rn(__fmul_rn(v[e], a.scale[c]), a.bias[c]);
        }
        __stcs(reinterpret_cast<float4*>(o + (size_t)c * plane), make_float4(v[0], v[1], v[2], v[3]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3b. fast stand-alone TTA augment for 4-byte elements with H, W multiples of 32 (same scheme as 2b)
// ------------------------------------------------------------------------------------------------
template <int TTA>
__global__ void __launch_bounds__(kThreads) tta_augment_fast_kernel(AugArgs a, int B, int C) {
  using V = Views<TTA>;
  __shared__ unsigned int sm[kChunkC * CH_STRIDE];
  const int per_img = a.nbx * a.nby;
  const int b = blockIdx.x / per_img;
  const int blk = blockIdx.x - b * per_img;
  const int by = blk / a.nbx, bx = blk - by * a.nbx;
  const int R0 = by * BS, C0 = bx * BS;
  const size_t plane = (size_t)a.H * a.W;
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const unsigned int* src = static_cast<const unsigned int*>(a.x) + (size_t)b * C * plane;
  unsigned int* dst = static_cast<unsigned int*>(a.out) + (size_t)b * C * plane;
  for (int c0 = 0; c0 < C; c0 += kChunkC) {
    const int cc = min(kChunkC, C - c0);
    if (c0) __syncthreads();
    for (int ch = 0; ch < cc; ++ch) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)(c0 + ch) * plane + (size_t)(R0 + li) * a.W + C0 + 4 * lq));
      unsigned int* s = sm + ch * CH_STRIDE + li * PITCH + 4 * lq;
      s[0] = q.x; s[1] = q.y; s[2] = q.z; s[3] = q.w;
    }
    __syncthreads();
#pragma unroll
    for (int va = 0; va < V::count; ++va) {
      const unsigned code = (V::packed >> (3 * va)) & 7u;
      int base, step, oi0, oj0;
      if (code & V_T) {
        const int col = (code & V_FJ) ? (BS - 1 - li) : li;
        const int row0 = (code & V_FI) ? (BS - 1 - 4 * lq) : 4 * lq;
        base = row0 * PITCH + col;
        step = (code & V_FI) ? -PITCH : PITCH;
        oi0 = (code & V_FJ) ? a.W - C0 - BS : C0;
        oj0 = (code & V_FI) ? a.H - R0 - BS : R0;
      } else {
        const int row = (code & V_FI) ? (BS - 1 - li) : li;
        const int col0 = (code & V_FJ) ? (BS - 1 - 4 * lq) : 4 * lq;
        base = row * PITCH + col0;
        step = (code & V_FJ) ? -1 : 1;
        oi0 = (code & V_FI) ? a.H - R0 - BS : R0;
        oj0 = (code & V_FJ) ? a.W - C0 - BS : C0;
      }
      unsigned int* o = dst + (size_t)va * a.view_stride + (size_t)c0 * plane + (size_t)(oi0 + li) * a.W + oj0 + 4 * lq;
      for (int ch = 0; ch < cc; ++ch) {
        const unsigned int* s = sm + ch * CH_STRIDE + base;
        *reinterpret_cast<uint4*>(o + (size_t)ch * plane) = make_uint4(s[0], s[step], s[2 * step], s[3 * step]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 4. stand-alone TTA de-augment (+ reduction)
// ------------------------------------------------------------------------------------------------
struct DeaugArgs {
  const void* x;
  void* out;
  int planes;  // B*C
  int H, W, nbx, nby, red;
  ViewSet views;
  size_t view_stride;
};

template <typename T, bool ALIGNED>
__device__ __forceinline__ void store_out4(T* o, int i, int j, int H, int W, const T v[4]) {
  if (i < H) {
    if (ALIGNED && j + 3 < W) {
      st4(o, v);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (j + e < W) o[e] = v[e];
    }
  }
}

template <typename T, bool ALIGNED, unsigned PACKED, int... Is>
__device__ __forceinline__ void deaug_keep_views(const DeaugArgs& a, const T* plane0, int p, int i0, int j0, T* sm,
                                                 std::integer_sequence<int, Is...>) {
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int i = i0 + li, j = j0 + 4 * lq;
  // one view at a time through the same machinery (a 1-view "sum" is the inverse-permuted copy)
  (
      [&] {
        T v[4];
        deaug_reduce4<T, ALIGNED, (PACKED >> (3 * Is)) & 7u, 1>(plane0 + (size_t)Is * a.view_stride, 0, TB_RED_SUM, i0,
                                                               j0, a.H, a.W, sm, v);
        T* o = static_cast<T*>(a.out) + (size_t)Is * a.view_stride + (size_t)p * a.H * a.W + (size_t)i * a.W + j;
        store_out4<T, ALIGNED>(o, i, j, a.H, a.W, v);
      }(),
      ...);
}

template <typename T, bool ALIGNED, int TTA>
__global__ void __launch_bounds__(kThreads) tta_deaugment_kernel(DeaugArgs a) {
  using V = Views<TTA>;
  __shared__ T sm[(packed_has_transpose(V::packed, V::count) ? 4 : 1) * BS * PITCH];
  const int per_plane = a.nbx * a.nby;
  const int p = blockIdx.x / per_plane;
  const int blk = blockIdx.x - p * per_plane;
  const int by = blk / a.nbx, bx = blk - by * a.nbx;
  const int i0 = by * BS, j0 = bx * BS;
  const T* plane0 = static_cast<const T*>(a.x) + (size_t)p * a.H * a.W;
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int i = i0 + li, j = j0 + 4 * lq;
  if (a.red != TB_RED_NONE) {
    T v[4];
    deaug_reduce4<T, ALIGNED, V::packed, V::count>(plane0, a.view_stride, a.red, i0, j0, a.H, a.W, sm, v);
    T* o = static_cast<T*>(a.out) + (size_t)p * a.H * a.W + (size_t)i * a.W + j;
    store_out4<T, ALIGNED>(o, i, j, a.H, a.W, v);
  } else {
    deaug_keep_views<T, ALIGNED, V::packed>(a, plane0, p, i0, j0, sm, std::make_integer_sequence<int, V::count>{});
  }
}

// ------------------------------------------------------------------------------------------------
// 5. streaming scatter merge with fused de-augment (TileMerger.integrate_batch [+ *_image_deaugment])
// ------------------------------------------------------------------------------------------------
template <typename T>
struct MergeArgs {
  const T* pred;  // [A*n,K,Th,Tw]
  T* image;       // [K,Hp,Wp]
  T* norm;        // [Hp,Wp] or null
  TileXY xy;
  int n, K, Th, Tw, Hp, Wp, nbx, nby, red;
  ViewSet views;
  WeightDev<T> w;
};

template <typename T, bool ALIGNED, int TTA>
__global__ void __launch_bounds__(kThreads) deaugment_merge_kernel(MergeArgs<T> a) {
  using V = Views<TTA>;
  __shared__ T sm[(packed_has_transpose(V::packed, V::count) ? 4 : 1) * BS * PITCH];
  const int per_tile = a.nbx * a.nby;
  const int t = blockIdx.x / per_tile;
  const int blk = blockIdx.x - t * per_tile;
  const int by = blk / a.nbx, bx = blk - by * a.nbx;
  const int i0 = by * BS, j0 = bx * BS;
  int x0, y0;
  a.xy.get(t, x0, y0);
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int i = i0 + li, j = j0 + 4 * lq;
  const size_t plane = (size_t)a.Th * a.Tw;
  const size_t view_stride = (size_t)a.n * a.K * plane;
  const bool row_ok = i < a.Th;
  // weights of this thread's 4 pixels
  T w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = (row_ok && j + e < a.Tw) ? weight_at<T>(a.w, i, j + e) : T(0);
  const size_t cpix = (size_t)(y0 + i) * a.Wp + (x0 + j);
  // vector reductions need a 16-byte aligned canvas address
  const bool vec_ok = ALIGNED && ((x0 & 3) == 0) && ((a.Wp & 3) == 0) && (j + 3 < a.Tw);
  for (int k = 0; k < a.K; ++k) {
    T v[4];
    deaug_reduce4<T, ALIGNED, V::packed, V::count>(a.pred + ((size_t)t * a.K + k) * plane, view_stride, a.red, i0,
                                                   j0, a.Th, a.Tw, sm, v);
    if (row_ok) {
      T* dst = a.image + (size_t)k * a.Hp * a.Wp + cpix;
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = mul_rn(v[e], w[e]);
      if (vec_ok) {
        red_add4(dst, v);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e < a.Tw) red_add1(dst + e, v[e]);
      }
    }
  }
  if (a.norm && row_ok) {
    T* dst = a.norm + cpix;
    if (vec_ok) {
      red_add4(dst, w);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (j + e < a.Tw) red_add1(dst + e, w[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 6. normalise + crop
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) normalize_crop_kernel(const T* __restrict__ image,
                                                                  const T* __restrict__ norm, int K, int Hp, int Wp,
                                                                  int top, int left, int H, int W, T eps, int use_eps,
                                                                  T* out) {
  const long long total = (long long)H * W;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int i = (int)(idx / W), j = (int)(idx - (long long)i * W);
    const size_t src = (size_t)(top + i) * Wp + (left + j);
    T d = norm[src];
    if (use_eps) d = max(d, eps);
    for (int k = 0; k < K; ++k) out[(size_t)k * total + idx] = image[(size_t)k * Hp * Wp + src] / d;
  }
}

// ------------------------------------------------------------------------------------------------
// 6b. class map / binary masks from merged logits (output side: 1 byte per pixel leaves the GPU)
// ------------------------------------------------------------------------------------------------
// mode 0: out[i] = argmax_k logits[k][i] (first maximum wins, NaN counts as maximal: torch.argmax)
// mode 1: out[k][i] = sigmoid(logits[k][i]) > thr
template <int VW>
__global__ void __launch_bounds__(kThreads) logits_to_labels_kernel(const float* __restrict__ logits, int K,
                                                                     long long plane_stride, long long n, int mode,
                                                                     float thr, unsigned char* __restrict__ out,
                                                                     long long out_plane_stride) {
  const long long nv = n / VW;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < nv; v += (long long)gridDim.x * kThreads) {
    const long long i0 = v * VW;
    if (mode == 0) {
      float best[VW];
      int arg[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        best[e] = 0.f;
        arg[e] = 0;
      }
      for (int k = 0; k < K; ++k) {
        float x[VW];
        if constexpr (VW == 4) {
          const float4 q = __ldcs(reinterpret_cast<const float4*>(logits + (size_t)k * plane_stride + i0));
          x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
        } else {
          x[0] = __ldcs(logits + (size_t)k * plane_stride + i0);
        }
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          const bool take = (k == 0) || (x[e] > best[e]) || (x[e] != x[e] && best[e] == best[e]);
          if (take) {
            best[e] = x[e];
            arg[e] = k;
          }
        }
      }
      if constexpr (VW == 4)
        *reinterpret_cast<uchar4*>(out + i0) = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1],
                                                           (unsigned char)arg[2], (unsigned char)arg[3]);
      else
        out[i0] = (unsigned char)arg[0];
    } else {
      for (int k = 0; k < K; ++k) {
        float x[VW];
        if constexpr (VW == 4) {
          const float4 q = __ldcs(reinterpret_cast<const float4*>(logits + (size_t)k * plane_stride + i0));
          x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
        } else {
          x[0] = __ldcs(logits + (size_t)k * plane_stride + i0);
        }
        unsigned char m[VW];
#pragma unroll
        for (int e = 0; e < VW; ++e) m[e] = (1.f / (1.f + expf(-x[e]))) > thr ? 1 : 0;
        unsigned char* o = out + (size_t)k * out_plane_stride + i0;
        if constexpr (VW == 4)
          *reinterpret_cast<uchar4*>(o) = make_uchar4(m[0], m[1], m[2], m[3]);
        else
          o[0] = m[0];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 7. output-stationary fused de-augment + merge + normalise + crop
// ------------------------------------------------------------------------------------------------
// MergeAllArgs: tiles_args.cuh

template <bool ALIGNED, int TTA>
__global__ void __launch_bounds__(kThreads) merge_all_kernel(MergeAllArgs a) {
  using V = Views<TTA>;
  __shared__ float sm[(packed_has_transpose(V::packed, V::count) ? 4 : 1) * BS * PITCH];
  const tb_grid& g = a.grid;
  const int by = blockIdx.x / a.nbx, bxi = blockIdx.x - by * a.nbx;
  const int Y0 = a.y0 - (a.y0 % BS) + by * BS;  // blocks are aligned to the canvas origin
  const int X0 = (a.bx0 + bxi) * BS;
  const int li = threadIdx.x >> 3, lq = threadIdx.x & 7;
  const int y = Y0 + li, x = X0 + 4 * lq;
  const size_t plane = (size_t)g.Th * g.Tw;
  const size_t view_stride = (size_t)g.n_slots * a.K * plane;
  const int rows = a.y1 - a.y0;

  // covering tile rows / columns of this block (any pixel of the block)
  int ra = (Y0 - g.Th + g.Sy) / g.Sy;  // ceil((Y0 - Th + 1)/Sy) for non-negative numerators
  if (Y0 - g.Th + 1 <= 0) ra = 0;
  int rb = (Y0 + BS - 1) / g.Sy;
  ra = max(ra, a.r0);
  rb = min(rb, min(g.ny, a.r1) - 1);
  int ca = (X0 - g.Tw + g.Sx) / g.Sx;
  if (X0 - g.Tw + 1 <= 0) ca = 0;
  int cb = min((X0 + BS - 1) / g.Sx, g.nx - 1);

  const bool y_ok = (y >= a.y0 && y < a.y1);
  LabelState ls;
  for (int k = 0; k < a.K; ++k) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.init_num && y_ok) {
      const size_t o = (size_t)(y - a.y0) * a.Wp + x;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (x + e < a.Wp) {
          acc[e] = a.init_num[(size_t)k * rows * a.Wp + o + e];
          den[e] = a.init_den[o + e];
        }
    }
    if (V::count == 1 && ALIGNED) {
      // Single view (no TTA): nothing to stage or reduce, so the only enemy is latency. Walk the covering tiles in
      // chunks of four and issue all prediction and weight loads of a chunk before consuming any of them.
      int r = ra, c = ca;
      while (r <= rb) {
        float pv[4][4], wv[4][4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ok[u] = false;
          if (r <= rb) {
            const int t = r * g.nx + c;
            const float* base = nullptr;
            if (g.tile_ptr) {
              base = g.tile_ptr[t];
            } else {
              const int slot = g.tile_slot ? __ldg(g.tile_slot + t) : t;
              if (slot >= 0) base = a.pred + (size_t)slot * a.K * plane;
            }
            const int i = Y0 - r * g.Sy + li, j = X0 - c * g.Sx + 4 * lq;
            if (base && i >= 0 && i < g.Th && j >= 0 && j + 3 < g.Tw) {
              ok[u] = true;
              ld_stream4(base + (size_t)k * plane + (size_t)i * g.Tw + j, pv[u]);
              if (a.w.kind == TB_W_TABLE) {
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
              const float val = red_post<float>(a.red, red_pre<float>(a.red, pre_act(a.pre, a.temperature, pv[u][e])), 1);
              acc[e] = add_rn(acc[e], mul_rn(val, wv[u][e]));
              den[e] = add_rn(den[e], wv[u][e]);
            }
          }
        }
      }
    } else
    for (int r = ra; r <= rb; ++r) {
      for (int c = ca; c <= cb; ++c) {
        const int t = r * g.nx + c;
        const float* base;
        size_t vstride = view_stride;
        if (g.tile_ptr) {
          base = g.tile_ptr[t];
          if (!base) continue;  // uniform across the CTA
          vstride = (size_t)__ldg(g.tile_vstride + t);
        } else {
          const int slot = g.tile_slot ? __ldg(g.tile_slot + t) : t;
          if (slot < 0) continue;  // uniform across the CTA
          base = a.pred + (size_t)slot * a.K * plane;
        }
        const int i0 = Y0 - r * g.Sy, j0 = X0 - c * g.Sx;
        float v[4];
        deaug_reduce4<float, ALIGNED, V::packed, V::count>(base + (size_t)k * plane, vstride, a.red, i0, j0, g.Th,
                                                           g.Tw, sm, v, a.pre, a.temperature);
        const int i = i0 + li, j = j0 + 4 * lq;
        if (i >= 0 && i < g.Th) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (j + e >= 0 && j + e < g.Tw) {
              const float w = weight_at<float>(a.w, i, j + e);
              acc[e] = add_rn(acc[e], mul_rn(v[e], w));
              den[e] = add_rn(den[e], w);
            }
          }
        }
      }
    }
    if (!y_ok) continue;
    if (a.raw) {
      const size_t o = (size_t)(y - a.y0) * a.Wp + x;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (x + e < a.Wp) {
          a.out_num[(size_t)k * rows * a.Wp + o + e] = acc[e];
          if (k == 0) a.out_den[o + e] = den[e];
        }
    } else {
      float q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) q[e] = acc[e] / den[e];
      merge_epilogue(a, k, y - a.top, x - a.left, q, ls);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Shared pieces of the TMA-pipelined kernels (7b, 7c)
// ------------------------------------------------------------------------------------------------
// Consumer thread -> pixel mapping: row li = 8*(warp&3) + (lane&7), 16-byte chunk cq = 4*(warp>>2) + (lane>>3)
// (columns 4*cq .. 4*cq+3). A quarter-warp touches 8 rows with distinct (row & 7), so with the 128-byte swizzle the
// LDS.128 of non-transposing views are bank-conflict free; transposing views read 4 scalars down a column.
template <unsigned PACKED, int NV>
struct TmaOffsets {
  int vec[NV];     // float index of the float4 (non-transposing views)
  int col[NV][4];  // float indices (transposing views)
  __device__ __forceinline__ void init(int li, int cq) {
    const int lj = 4 * cq;
#pragma unroll
    for (int va = 0; va < NV; ++va) {
      const unsigned gc = view_inverse((PACKED >> (3 * va)) & 7u);
      if (gc & V_T) {
        const int cc = (gc & V_FJ) ? BS - 1 - li : li;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int rr = (gc & V_FI) ? BS - 1 - (lj + e) : lj + e;
          col[va][e] = va * kBoxFloats + swz(rr, cc);
        }
        vec[va] = 0;
      } else {
        const int rr = (gc & V_FI) ? BS - 1 - li : li;
        const int chunk = (gc & V_FJ) ? 7 - cq : cq;
        vec[va] = va * kBoxFloats + rr * 32 + ((chunk ^ (rr & 7)) << 2);
#pragma unroll
        for (int e = 0; e < 4; ++e) col[va][e] = 0;
      }
    }
  }
  // sequential view-order sum of the (transformed) view values of this thread's 4 pixels
  template <bool LINEAR>
  __device__ __forceinline__ void reduce(const float* st, int red, float v[4], int pre = TB_PRE_NONE, float temperature = 1.f) const {
#pragma unroll
    for (int va = 0; va < NV; ++va) {
      const unsigned gc = view_inverse((PACKED >> (3 * va)) & 7u);
      float q[4];
      if (gc & V_T) {
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = st[col[va][e]];
      } else {
        const float4 f = *reinterpret_cast<const float4*>(st + vec[va]);
        if (gc & V_FJ) {
          q[0] = f.w; q[1] = f.z; q[2] = f.y; q[3] = f.x;
        } else {
          q[0] = f.x; q[1] = f.y; q[2] = f.z; q[3] = f.w;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = LINEAR ? q[e] : red_pre<float>(red, pre_act(pre, temperature, q[e]));
        v[e] = (va == 0) ? pv : add_rn(v[e], pv);
      }
    }
  }
};

template <bool LINEAR, int NV>
__device__ __forceinline__ float tma_post(int red, float v) {
  if (LINEAR) return (red == TB_RED_MEAN) ? v / (float)NV : v;
  return red_post<float>(red, v, NV);
}

// Producer side: issue the NV box loads of one (tile position `pos` in the map, channel k) for block origin (i0, j0).
template <unsigned PACKED, int NV>
__device__ __forceinline__ void tma_issue_views(float* stage, const CUtensorMap* map, uint64_t* bar, int pos, int K, int k,
                                                int Th, int Tw, int i0, int j0, bool hint = false, uint64_t policy = 0) {
#pragma unroll
  for (int va = 0; va < NV; ++va) {
    const unsigned gc = view_inverse((PACKED >> (3 * va)) & 7u);
    int sr0, sc0;
    if (gc & V_T) {
      sr0 = (gc & V_FI) ? Th - j0 - BS : j0;
      sc0 = (gc & V_FJ) ? Tw - i0 - BS : i0;
    } else {
      sr0 = (gc & V_FI) ? Th - i0 - BS : i0;
      sc0 = (gc & V_FJ) ? Tw - j0 - BS : j0;
    }
    if (hint)
      tma_load_box_hint(stage + (size_t)va * kBoxFloats, map, bar, sc0, (pos * K + k) * Th + sr0, va, policy);
    else
      tma_load_box(stage + (size_t)va * kBoxFloats, map, bar, sc0, (pos * K + k) * Th + sr0, va);
  }
}

// ------------------------------------------------------------------------------------------------
// 7b. TMA-pipelined output-stationary merge (requires Th, Tw, Sx, Sy multiples of 32 and a single pred tensor)
//
// Persistent CTAs: each CTA walks canvas blocks blockIdx.x, blockIdx.x + gridDim.x, ... and the stage ring
// keeps running across block boundaries, so the producer lane is always STAGES stages ahead of the
// consumers (no pipeline fill/drain per block).
// ------------------------------------------------------------------------------------------------
struct BlockWork {
  int Y0, X0, ra, rb, ca, cb;
};

__device__ __forceinline__ BlockWork block_work(const MergeAllArgs& a, int blk) {
  const tb_grid& g = a.grid;
  BlockWork w;
  const int by = blk / a.nbx, bxi = blk - by * a.nbx;
  w.Y0 = a.y0 - (a.y0 % BS) + by * BS;
  w.X0 = (a.bx0 + bxi) * BS;
  int ra = (w.Y0 - g.Th + g.Sy) / g.Sy;
  if (w.Y0 - g.Th + 1 <= 0) ra = 0;
  int rb = (w.Y0 + BS - 1) / g.Sy;
  w.ra = max(ra, a.r0);
  w.rb = min(rb, min(g.ny, a.r1) - 1);
  int ca = (w.X0 - g.Tw + g.Sx) / g.Sx;
  if (w.X0 - g.Tw + 1 <= 0) ca = 0;
  w.ca = ca;
  w.cb = min((w.X0 + BS - 1) / g.Sx, g.nx - 1);
  return w;
}

// fused ragged-strip launch: block classes handled by pass 0 / 1 / 2 (spill first: the neighbour waits for it; seeded last:
// we wait for the neighbour)
__device__ __forceinline__ int peer_pass_mask(int pass) {
  return pass == 0 ? TB_PH_SPILL : (pass == 1 ? TB_PH_FINAL : (TB_PH_SEEDED_FINAL | TB_PH_SEEDED_SPILL));
}

// Iterator over the resident covering tiles of a block, row-major (the reference's accumulation order).
struct TileIter {
  int r, c;
};
__device__ __forceinline__ bool tile_resident(const MergeAllArgs& a, int r, int c) {
  const int rel = r * a.grid.nx + c - a.first_tile;
  return rel >= 0 && rel < a.covered_tiles;
}
__device__ __forceinline__ void tile_iter_settle(const MergeAllArgs& a, const BlockWork& w, TileIter& p) {
  while (p.r <= w.rb && !tile_resident(a, p.r, p.c)) {
    if (++p.c > w.cb) {
      p.c = w.ca;
      ++p.r;
    }
  }
}
__device__ __forceinline__ void tile_iter_next(const MergeAllArgs& a, const BlockWork& w, TileIter& p) {
  if (++p.c > w.cb) {
    p.c = w.ca;
    ++p.r;
  }
  tile_iter_settle(a, w, p);
}

// Number of resident covering tiles of a block (identical on the producer and the consumer side).
__device__ __forceinline__ int block_tile_count(const MergeAllArgs& a, const BlockWork& w) {
  int nt = 0;
  for (int r = w.ra; r <= w.rb; ++r)
    for (int c = w.ca; c <= w.cb; ++c) {
      const int rel = r * a.grid.nx + c - a.first_tile;
      nt += (rel >= 0 && rel < a.covered_tiles) ? 1 : 0;
    }
  return nt;
}

// LINEAR: the reduction is MEAN or SUM (no per-element transform): keeps the switch out of the inner loop
// NG = independent pipelines ("virtual CTAs") inside the CTA: each has 8 consumer warps, one producer warp, its own
// share of the stage ring and walks its own canvas blocks. NG = 2 doubles the warps available to hide the consumers'
// fixed-latency stalls, which is what few-view stages (few bytes per barrier round trip) need; D4 keeps NG = 1 and the
// deepest ring.
template <int TTA, bool LINEAR, int MAXB, int NG>
__global__ void __launch_bounds__(kTmaThreads * NG, 1) merge_all_tma_kernel(MergeAllArgs a, const __grid_constant__ TmaMaps<MAXB> maps,
                                                                           int nblocks) {
  using V = Views<TTA>;
  constexpr int NV = V::count;
  constexpr int G = tma_group(NV);         // covering tiles per stage
  constexpr int SB = tma_stage_boxes(NV);  // boxes per stage
  constexpr int STAGES = tma_stages(NV, true) / NG;  // per pipeline
  extern __shared__ __align__(16) float smem_boxes[];
  // 1024-byte alignment for the 128-byte swizzle atom, computed on the SHARED address (the window of a CTA is not
  // always 1024-aligned when other kernels' CTAs are resident) and applied as an element offset, so the pointer
  // provably stays in the shared address space (LDS, not generic LD).
  float* boxes_all = smem_boxes + (((1024u - (smem_u32(smem_boxes) & 1023u)) & 1023u) >> 2);
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(boxes_all + NG * STAGES * SB * kBoxFloats);

  const tb_grid& g = a.grid;
  const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warps [0, 8*NG) consume (pipeline = warp / 8), warps [8*NG, 9*NG) produce
  const bool is_producer = warp_all >= kTmaConsumerWarps * NG;
  const int pipe = is_producer ? warp_all - kTmaConsumerWarps * NG : warp_all / kTmaConsumerWarps;
  const int warp = is_producer ? kTmaConsumerWarps : warp_all % kTmaConsumerWarps;
  float* boxes = boxes_all + (size_t)pipe * STAGES * SB * kBoxFloats;
  uint64_t* full = bars_all + (size_t)pipe * 2 * STAGES;
  uint64_t* empty = full + STAGES;
  const int blk0 = blockIdx.x * NG + pipe, blk_step = gridDim.x * NG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * STAGES * NG; ++s) mbar_init(&bars_all[s], ((s / STAGES) & 1) ? kTmaConsumerWarps : 1);
    mbar_fence_init();
  }
  __syncthreads();

  if (is_producer) {
    // ---------------- producer of this pipeline: one elected lane feeds its ring ----------------
    const int prod = 0;
    if (lane == 0) {
      int it = 0;
      const bool hint = (a.debug & 8) != 0;
      const uint64_t policy = hint ? l2_evict_first_policy() : 0ull;
      const int npass = a.peer.fused ? 3 : 1;
      for (int pass = 0; pass < npass; ++pass)
      for (int blk = blk0; blk < nblocks; blk += blk_step) {
        const BlockWork w = block_work(a, blk);
        if (a.peer.phase_mask) {  // (same test, same order as the consumers)
          const int cls = peer_block_class(a.peer, g.nx, w.ra, w.rb, w.ca, w.cb);
          if (!cls || (a.peer.fused && !(cls & peer_pass_mask(pass)))) continue;
        }
        const int nt = block_tile_count(a, w);
        for (int k = 0; k < a.K; ++k) {
          int idx = 0, s = 0;
          bool mine = false;
          for (int r = w.ra; r <= w.rb; ++r) {
            for (int c = w.ca; c <= w.cb; ++c) {
              const int rel = r * g.nx + c - a.first_tile;
              if (rel < 0 || rel >= a.covered_tiles) continue;  // tile not resident (band mode)
              const int gi = idx % G;
              if (gi == 0) {  // open a stage
                mine = (it % kTmaProducerWarps) == prod;
                if (mine) {
                  s = it % STAGES;
                  const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                  mbar_wait(&empty[s], ph ^ 1u);
                  if (a.debug & 2) fence_proxy_async();  // (A/B switch: producer-side fence, see mbar_release_after)
                  mbar_arrive_expect_tx(&full[s], (uint32_t)min(G, nt - idx) * NV * kBoxBytes);
                }
              }
              if (mine) {
                const int bi = rel / a.batch_tiles, pos = rel - bi * a.batch_tiles;
                const int i0 = w.Y0 - r * g.Sy, j0 = w.X0 - c * g.Sx;  // block origin inside the tile
                tma_issue_views<V::packed, NV>(boxes + ((size_t)s * SB + gi * NV) * kBoxFloats, &maps.m[MAXB == 1 ? 0 : bi],
                                               &full[s], pos, a.K, k, g.Th, g.Tw, i0, j0, hint, policy);
              }
              ++idx;
              if (gi == G - 1 || idx == nt) ++it;  // stage closed
            }
          }
        }
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  // Each thread owns 4 consecutive columns (one 16-byte chunk) of one row of the 32x32 block (TmaOffsets).
  const int li = 8 * (warp & 3) + (lane & 7);
  const int cq = 4 * (warp >> 2) + (lane >> 3);
  const int lj = 4 * cq;
  const int rows = a.y1 - a.y0;

  TmaOffsets<V::packed, NV> offs;  // loop-invariant shared-memory offsets of this thread's 4 pixels
  offs.init(li, cq);

  int it = 0;
  LabelState ls;
  const int npass = a.peer.fused ? 3 : 1;
  const unsigned long long arrivals = a.peer.epoch * (unsigned long long)(gridDim.x * NG);  // grid-wide target of the counters
  const bool leader = (warp == 0 && lane == 0);  // one thread per pipeline talks to the flags
  for (int pass = 0; pass < npass; ++pass) {
  if (a.peer.fused) {
    // ---- before the pass: the waits of the ragged-strip protocol, by one thread, then the pipeline's consumers meet ----
    if (pass == 0 && a.peer.has_next && a.peer.epoch > 2 && leader)
      peer_spin_until(a.peer.ack_flag, a.peer.epoch - 2, a.peer.timeout_ns, a.peer.status);   // double buffer free again?
    if (pass == 2 && a.peer.has_prev && leader)
      peer_spin_until(a.peer.ready_flag, a.peer.epoch, a.peer.timeout_ns, a.peer.status);     // seeds of this epoch complete?
    if ((pass == 0 && a.peer.has_next && a.peer.epoch > 2) || (pass == 2 && a.peer.has_prev))
      asm volatile("bar.sync %0, %1;" ::"r"(1 + pipe), "r"(kTmaConsumerWarps * 32) : "memory");
  }
  for (int blk = blk0; blk < nblocks; blk += blk_step) {
    const BlockWork w = block_work(a, blk);
    int cls = 0;
    if (a.peer.phase_mask) {
      cls = peer_block_class(a.peer, g.nx, w.ra, w.rb, w.ca, w.cb);
      if (!cls || (a.peer.fused && !(cls & peer_pass_mask(pass)))) continue;
    }
    const bool seeded = cls & (TB_PH_SEEDED_FINAL | TB_PH_SEEDED_SPILL), spill = cls & (TB_PH_SPILL | TB_PH_SEEDED_SPILL);
    const int nt = block_tile_count(a, w);
    const int y = w.Y0 + li;
    const bool y_ok = (y >= a.y0 && y < a.y1);
    for (int k = 0; k < a.K; ++k) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.init_num && y_ok) {
        const size_t o = (size_t)(y - a.y0) * a.Wp + w.X0 + lj;
        const float4 n4 = *reinterpret_cast<const float4*>(a.init_num + (size_t)k * rows * a.Wp + o);
        const float4 d4 = *reinterpret_cast<const float4*>(a.init_den + o);
        acc[0] = n4.x; acc[1] = n4.y; acc[2] = n4.z; acc[3] = n4.w;
        den[0] = d4.x; den[1] = d4.y; den[2] = d4.z; den[3] = d4.w;
      }
      if (seeded && y_ok) {  // partial sums of the lower ranks' tiles (written by the previous rank through NVLink)
        const size_t o = (size_t)(y - a.peer.seed_y0) * a.Wp + w.X0 + lj;
        const float4 n4 = __ldcg(reinterpret_cast<const float4*>(a.peer.seed_num + (size_t)k * a.peer.seed_rows * a.Wp + o));
        const float4 d4 = __ldcg(reinterpret_cast<const float4*>(a.peer.seed_den + o));
        acc[0] = n4.x; acc[1] = n4.y; acc[2] = n4.z; acc[3] = n4.w;
        den[0] = d4.x; den[1] = d4.y; den[2] = d4.z; den[3] = d4.w;
      }
      TileIter cur = {w.ra, w.ca};
      tile_iter_settle(a, w, cur);
      for (int g0 = 0; g0 < nt; g0 += G, ++it) {
        const int cnt = min(G, nt - g0);
        // the weights of ALL tiles of the stage first: their (L2) latencies overlap each other and the wait below
        float wt[G][4];
        {
          TileIter p = cur;
#pragma unroll
          for (int gi = 0; gi < G; ++gi) {
            if (gi < cnt) {
              const int ti = w.Y0 - p.r * g.Sy + li, tj = w.X0 - p.c * g.Sx + lj;
              if (a.w.kind == TB_W_TABLE) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w.table + (size_t)ti * g.Tw + tj));
                wt[gi][0] = w4.x; wt[gi][1] = w4.y; wt[gi][2] = w4.z; wt[gi][3] = w4.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) wt[gi][e] = weight_at<float>(a.w, ti, tj + e);
              }
              tile_iter_next(a, w, p);
            }
          }
          cur = p;
        }
        const int s = it % STAGES;
        mbar_wait(&full[s], (uint32_t)(it / STAGES) & 1u);
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          if (gi < cnt) {
            float v[4];
            offs.template reduce<LINEAR>(boxes + ((size_t)s * SB + gi * NV) * kBoxFloats, a.red, v, a.pre, a.temperature);
            if (gi == cnt - 1)  // slot back to the producer (debug bit 16: timing-only switch that drops the proxy fence)
              mbar_release_after(&empty[s], v[0], v[1], v[2], v[3], lane == 0, (a.debug & 16) != 0);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float r4 = tma_post<LINEAR, NV>(a.red, v[e]);
              acc[e] = add_rn(acc[e], mul_rn(r4, wt[gi][e]));
              den[e] = add_rn(den[e], wt[gi][e]);
            }
          }
        }
      }
      if (!y_ok) continue;
      if (spill) {  // later covering tiles live on higher ranks: hand the partial sums to the next rank (peer stores)
        const size_t o = (size_t)(y - a.peer.spill_y0) * a.Wp + w.X0 + lj;
        *reinterpret_cast<float4*>(a.peer.spill_num + (size_t)k * a.peer.spill_rows * a.Wp + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (k == 0) *reinterpret_cast<float4*>(a.peer.spill_den + o) = make_float4(den[0], den[1], den[2], den[3]);
      } else if (a.raw) {
        const size_t o = (size_t)(y - a.y0) * a.Wp + w.X0 + lj;
        *reinterpret_cast<float4*>(a.out_num + (size_t)k * rows * a.Wp + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (k == 0) *reinterpret_cast<float4*>(a.out_den + o) = make_float4(den[0], den[1], den[2], den[3]);
      } else {
        const float q[4] = {acc[0] / den[0], acc[1] / den[1], acc[2] / den[2], acc[3] / den[3]};
        merge_epilogue(a, k, y - a.top, w.X0 + lj - a.left, q, ls);
      }
    }
  }
  if (a.peer.fused) {
    // ---- after the pass: every consumer makes its peer stores visible, the pipeline meets, one thread counts the arrival;
    //      the last arrival of the whole grid publishes the flag ----
    const bool publish_ready = a.peer.has_next && ((pass == 0 && !a.peer.chained) || (pass == 2 && a.peer.chained));
    const bool publish_ack = a.peer.has_prev && pass == 2;
    if (publish_ready || publish_ack) {
      __threadfence_system();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + pipe), "r"(kTmaConsumerWarps * 32) : "memory");
      if (leader) {
        if (publish_ready) peer_arrive_and_publish(a.peer.counters + (pass == 0 ? 0 : 2), arrivals, a.peer.next_ready, a.peer.epoch);
        if (publish_ack) peer_arrive_and_publish(a.peer.counters + 1, arrivals, a.peer.prev_ack, a.peer.epoch);
      }
    }
  }
  }  // pass
}

// ------------------------------------------------------------------------------------------------
// 7c. TMA-pipelined per-tile de-augment (stand-alone tta.*_image_deaugment) and streaming scatter merge
//     (TileMerger.integrate_batch with the de-augment fused). Same ring as 7b; a block is a 32x32 patch of one tile.
// ------------------------------------------------------------------------------------------------
struct TileTmaArgs {
  float* out;    // DEAUG: [n,K,Th,Tw]
  float* image;  // SCATTER: [K,Hp,Wp]
  float* norm;   // SCATTER: [Hp,Wp] or null
  TileXY xy;     // SCATTER: canvas coordinates of the tiles
  int n, K, Th, Tw, Hp, Wp, nbx, nby, red;
  WeightDev<float> w;
};

template <int TTA, bool LINEAR, bool SCATTER>
__global__ void __launch_bounds__(kTmaThreads, 2) tile_deaug_tma_kernel(TileTmaArgs a, const __grid_constant__ TmaMaps<1> maps,
                                                                       int nblocks) {
  using V = Views<TTA>;
  constexpr int NV = V::count;
  constexpr int G = tma_group(NV);  // channels per stage
  constexpr int SB = tma_stage_boxes(NV);
  constexpr int STAGES = tma_stages(NV);
  extern __shared__ __align__(16) float smem_boxes[];
  // 1024-byte alignment for the 128-byte swizzle atom, computed on the SHARED address (the window of a CTA is not
  // always 1024-aligned when other kernels' CTAs are resident) and applied as an element offset, so the pointer
  // provably stays in the shared address space (LDS, not generic LD).
  float* boxes = smem_boxes + (((1024u - (smem_u32(smem_boxes) & 1023u)) & 1023u) >> 2);
  uint64_t* full = reinterpret_cast<uint64_t*>(boxes + STAGES * SB * kBoxFloats);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int per_tile = a.nbx * a.nby;

  if (warp >= kTmaConsumerWarps) {
    const int prod = warp - kTmaConsumerWarps;
    if (lane == 0) {
      int it = 0;
      for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int t = blk / per_tile, rem = blk - t * per_tile;
        const int i0 = (rem / a.nbx) * BS, j0 = (rem % a.nbx) * BS;
        for (int k0 = 0; k0 < a.K; k0 += G, ++it) {
          if ((it % kTmaProducerWarps) != prod) continue;
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((uint32_t)(it / STAGES) & 1u) ^ 1u);
          const int cnt = min(G, a.K - k0);
          mbar_arrive_expect_tx(&full[s], (uint32_t)cnt * NV * kBoxBytes);
          for (int gi = 0; gi < cnt; ++gi)
            tma_issue_views<V::packed, NV>(boxes + ((size_t)s * SB + gi * NV) * kBoxFloats, &maps.m[0], &full[s], t, a.K, k0 + gi,
                                           a.Th, a.Tw, i0, j0);
        }
      }
    }
    return;
  }

  const int li = 8 * (warp & 3) + (lane & 7);
  const int cq = 4 * (warp >> 2) + (lane >> 3);
  const int lj = 4 * cq;
  TmaOffsets<V::packed, NV> offs;
  offs.init(li, cq);
  const size_t plane = (size_t)a.Th * a.Tw;
  int it = 0;
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int t = blk / per_tile, rem = blk - t * per_tile;
    const int i = (rem / a.nbx) * BS + li, j = (rem % a.nbx) * BS + lj;
    float wt[4] = {1.f, 1.f, 1.f, 1.f};
    size_t cpix = 0;
    bool vec_ok = true;
    if (SCATTER) {
      int x0, y0;
      a.xy.get(t, x0, y0);
      if (a.w.kind == TB_W_TABLE) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w.table + (size_t)i * a.Tw + j));
        wt[0] = w4.x; wt[1] = w4.y; wt[2] = w4.z; wt[3] = w4.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) wt[e] = weight_at<float>(a.w, i, j + e);
      }
      cpix = (size_t)(y0 + i) * a.Wp + (x0 + j);
      vec_ok = ((x0 & 3) == 0) && ((a.Wp & 3) == 0);
    }
    for (int k0 = 0; k0 < a.K; k0 += G, ++it) {
      const int s = it % STAGES;
      const int cnt = min(G, a.K - k0);
      mbar_wait(&full[s], (uint32_t)(it / STAGES) & 1u);
      for (int gi = 0; gi < cnt; ++gi) {
        const int k = k0 + gi;
        float v[4];
        offs.template reduce<LINEAR>(boxes + ((size_t)s * SB + gi * NV) * kBoxFloats, a.red, v);
        if (gi == cnt - 1) mbar_release_after(&empty[s], v[0], v[1], v[2], v[3], lane == 0);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = tma_post<LINEAR, NV>(a.red, v[e]);
        if (SCATTER) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = mul_rn(v[e], wt[e]);
          float* dst = a.image + (size_t)k * a.Hp * a.Wp + cpix;
          if (vec_ok) {
            red_add4(dst, v);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) red_add1(dst + e, v[e]);
          }
        } else {
          __stcs(reinterpret_cast<float4*>(a.out + ((size_t)t * a.K + k) * plane + (size_t)i * a.Tw + j),
                 make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    }
    if (SCATTER && a.norm) {
      float* dst = a.norm + cpix;
      if (vec_ok) {
        red_add4(dst, wt);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) red_add1(dst + e, wt[e]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 8. stand-in model (bench / tests only): y[b,k,i,j] = sum_c Wm[k,c] x[b,c,i,j] + 0.01 (i - 2j + 7k)/T
// ------------------------------------------------------------------------------------------------
struct StandinArgs {
  const float* x;
  float* y;
  int B, Cin, K, Th, Tw;
  float Wm[64];
};

__global__ void __launch_bounds__(kThreads) standin_model_kernel(StandinArgs a) {
  const size_t plane = (size_t)a.Th * a.Tw;
  const size_t quads = plane / 4;
  const size_t total = (size_t)a.B * quads;
  for (size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kThreads) {
    const size_t b = idx / quads, q = idx - b * quads;
    const int i = (int)((q * 4) / a.Tw), j = (int)((q * 4) - (size_t)i * a.Tw);
    float4 xin[8];
    for (int c = 0; c < a.Cin; ++c) xin[c] = __ldcs(reinterpret_cast<const float4*>(a.x + (b * a.Cin + c) * plane) + q);
    for (int k = 0; k < a.K; ++k) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = 0.01f * (float)(i - 2 * (j + e) + 7 * k) / (float)a.Th;
      for (int c = 0; c < a.Cin; ++c) {
        const float wkc = a.Wm[k * a.Cin + c];
        o[0] = fmaf(wkc, xin[c].x, o[0]);
        o[1] = fmaf(wkc, xin[c].y, o[1]);
        o[2] = fmaf(wkc, xin[c].z, o[2]);
        o[3] = fmaf(wkc, xin[c].w, o[3]);
      }
      __stcs(reinterpret_cast<float4*>(a.y + (b * a.K + k) * plane) + q, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

// TB200_FORCE_GENERIC_MERGE=1 routes tb_merge_all through the non-TMA kernel (used by tests to cross-check both).
static inline bool tb_force_generic_merge() {
  const char* v = getenv("TB200_FORCE_GENERIC_MERGE");
  return v && v[0] == '1';
}

static inline int slice_variant() {
  const char* v = getenv("TB200_SLICE_VARIANT");
  return v ? atoi(v) : 3;  // 6 CTAs/SM + streaming stores (measured best on cfg2: 1.60 ms vs 1.67-1.75 ms)
}
static inline int merge_variant() {
  const char* v = getenv("TB200_MERGE_VARIANT");
  return v ? atoi(v) : 0;  // bit1: L2 evict-first hint on the TMA loads (A/B switch; no measurable gain)
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// TB200_SINGLE_VIEW=0 keeps TTA_NONE on the general kernels (A/B switch; tests cross-check the two bit for bit)
static inline bool single_view_enabled() {
  const char* v = getenv("TB200_SINGLE_VIEW");
  return !(v && v[0] == '0');
}

}  // namespace tb

using namespace tb;

// Run `BODY` with TTA_ bound to the compile-time TTA group.
#define TB_DISPATCH_TTA(tta, ...)                                        \
  switch (tta) {                                                         \
    case TB_TTA_NONE: { constexpr int TTA_ = TB_TTA_NONE; __VA_ARGS__; break; }     \
    case TB_TTA_FLIPLR: { constexpr int TTA_ = TB_TTA_FLIPLR; __VA_ARGS__; break; } \
    case TB_TTA_FLIPUD: { constexpr int TTA_ = TB_TTA_FLIPUD; __VA_ARGS__; break; } \
    case TB_TTA_FLIPS: { constexpr int TTA_ = TB_TTA_FLIPS; __VA_ARGS__; break; }   \
    case TB_TTA_D2: { constexpr int TTA_ = TB_TTA_D2; __VA_ARGS__; break; }         \
    case TB_TTA_D4: { constexpr int TTA_ = TB_TTA_D4; __VA_ARGS__; break; }         \
    default: break;                                                      \
  }

// ================================================================================================
// C-ABI entry points
// ================================================================================================

static PFN_cuTensorMapEncodeTiled get_tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

// One batch tensor [A*b, K, Th, Tw] as the rank-3 fp32 tensor {Tw, b*K*Th, A}; box = 32 x 32 x 1, 128-byte swizzle
static bool make_pred_tensor_map(const float* base, int b, int K, int Th, int Tw, int A, CUtensorMap* map) {
  PFN_cuTensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return false;
  const unsigned long long rows = (unsigned long long)b * K * Th;
  if (rows >= (1ull << 31)) return false;
  cuuint64_t dims[3] = {(cuuint64_t)Tw, (cuuint64_t)rows, (cuuint64_t)A};
  cuuint64_t strides[2] = {(cuuint64_t)Tw * sizeof(float), (cuuint64_t)rows * Tw * sizeof(float)};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


template <int TTA, bool LINEAR, int MAXB, int NG>
static cudaError_t launch_merge_all_tma_impl2(const MergeAllArgs& a, const TmaMaps<MAXB>& maps, int nblocks, cudaStream_t s) {
  constexpr int NV = Views<TTA>::count;
  const size_t smem = tma_smem_bytes(NV, true);
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(merge_all_tma_kernel<TTA, LINEAR, MAXB, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  const int grid = std::min((nblocks + NG - 1) / NG, sm_count());  // persistent: one CTA per SM
  merge_all_tma_kernel<TTA, LINEAR, MAXB, NG><<<grid, kTmaThreads * NG, smem, s>>>(a, maps, nblocks);
  return cudaGetLastError();
}

template <int TTA, bool LINEAR, int MAXB>
static cudaError_t launch_merge_all_tma_impl(const MergeAllArgs& a, const TmaMaps<MAXB>& maps, int nblocks, cudaStream_t s) {
  // D4 (8 views, 32 KB per tile and channel): one pipeline with the deepest ring. Fewer views: two pipelines per CTA.
  // TB200_MERGE_VARIANT bit0 flips the choice (A/B switch).
  constexpr int NV = Views<TTA>::count;
  const bool two = (NV < 8) != ((merge_variant() & 1) != 0);
  if (two) return launch_merge_all_tma_impl2<TTA, LINEAR, MAXB, 2>(a, maps, nblocks, s);
  return launch_merge_all_tma_impl2<TTA, LINEAR, MAXB, 1>(a, maps, nblocks, s);
}

template <int TTA, int MAXB>
static cudaError_t launch_merge_all_tma(const MergeAllArgs& a, const TmaMaps<MAXB>& maps, int nblocks, cudaStream_t s) {
  if ((a.red == TB_RED_MEAN || a.red == TB_RED_SUM) && a.pre == TB_PRE_NONE) return launch_merge_all_tma_impl<TTA, true, MAXB>(a, maps, nblocks, s);
  return launch_merge_all_tma_impl<TTA, false, MAXB>(a, maps, nblocks, s);
}

template <int TTA, bool LINEAR, bool SCATTER>
static cudaError_t launch_tile_deaug_tma_impl(const TileTmaArgs& a, const TmaMaps<1>& maps, int nblocks, cudaStream_t s) {
  constexpr int NV = Views<TTA>::count;
  const size_t smem = tma_smem_bytes(NV);
  static int ctas_per_sm_dev[kMaxDevices] = {};
  const int dev = current_device();
  if (!ctas_per_sm_dev[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tile_deaug_tma_kernel<TTA, LINEAR, SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_deaug_tma_kernel<TTA, LINEAR, SCATTER>, kTmaThreads, smem);
    if (e != cudaSuccess) return e;
    ctas_per_sm_dev[dev] = n > 0 ? n : 1;
  }
  const int grid = std::min(nblocks, sm_count() * ctas_per_sm_dev[dev]);
  tile_deaug_tma_kernel<TTA, LINEAR, SCATTER><<<grid, kTmaThreads, smem, s>>>(a, maps, nblocks);
  return cudaGetLastError();
}

template <int TTA, bool SCATTER>
static cudaError_t launch_tile_deaug_tma(const TileTmaArgs& a, const TmaMaps<1>& maps, int nblocks, cudaStream_t s) {
  if (a.red == TB_RED_MEAN || a.red == TB_RED_SUM) return launch_tile_deaug_tma_impl<TTA, true, SCATTER>(a, maps, nblocks, s);
  return launch_tile_deaug_tma_impl<TTA, false, SCATTER>(a, maps, nblocks, s);
}

// Geometry test for the TMA path (everything on the 32-pixel lattice, 16-byte aligned side buffers).
static bool tma_geometry_ok(const tb_grid* grid, int Wp, const float* init_num, const float* init_den, int raw,
                            const float* out_num, const float* out_den, const tb_weight* weight) {
  return (grid->Th % BS == 0) && (grid->Tw % BS == 0) && (grid->Sx % BS == 0) && (grid->Sy % BS == 0) && (Wp % 4 == 0) &&
         aligned16(out_num) && (weight->kind != TB_W_TABLE || aligned16(weight->table)) &&
         (!init_num || (aligned16(init_num) && aligned16(init_den))) && (!raw || aligned16(out_den)) &&
         !tb_force_generic_merge();
}

extern "C" {

int tb_tta_views(int tta) {
  ViewSet v;
  if (!make_views(tta, &v)) return fail(TB_ERR_ARG, "tb_tta_views: unknown tta %d", tta);
  return v.count;
}

int tb_crop_tiles(const void* img, int elem_bytes, int H, int W, int C, const int32_t* xy, int n, int Th, int Tw,
                  int border_type, const uint64_t* border_value_host, int clip_to_tile, void* out, void* stream) {
  TB_REQUIRE(img && xy && out, TB_ERR_ARG, "tb_crop_tiles: null pointer");
  TB_REQUIRE(H > 0 && W > 0 && C > 0 && Th > 0 && Tw > 0 && n >= 0, TB_ERR_ARG, "tb_crop_tiles: bad shape");
  TB_REQUIRE(border_type >= 0 && border_type <= 4, TB_ERR_ARG, "tb_crop_tiles: bad border_type %d", border_type);
  TB_REQUIRE(n <= 65535, TB_ERR_ARG, "tb_crop_tiles: n > 65535");
  if (n == 0) return TB_OK;
  CropArgs a;
  a.img = img; a.out = out; a.xy = xy;
  a.H = H; a.W = W; a.C = C; a.n = n; a.Th = Th; a.Tw = Tw; a.border = border_type; a.clip_to_tile = clip_to_tile;
  for (int c = 0; c < 4; ++c) a.bval[c] = border_value_host ? border_value_host[c] : 0ull;
  const long long total = (long long)Th * Tw * C;
  const int gx = (int)std::min<long long>((total + kThreads - 1) / kThreads, 4096);
  dim3 grid(gx, n);
  cudaStream_t s = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: crop_tiles_kernel<unsigned char><<<grid, kThreads, 0, s>>>(a); break;
    case 2: crop_tiles_kernel<unsigned short><<<grid, kThreads, 0, s>>>(a); break;
    case 4: crop_tiles_kernel<unsigned int><<<grid, kThreads, 0, s>>>(a); break;
    case 8: crop_tiles_kernel<unsigned long long><<<grid, kThreads, 0, s>>>(a); break;
    default: return fail(TB_ERR_ARG, "tb_crop_tiles: elem_bytes must be 1, 2, 4 or 8 (got %d)", elem_bytes);
  }
  TB_LAUNCH_CHECK("crop_tiles_kernel");
  return TB_OK;
}

int tb_slice_augment(const void* img, int img_dtype, int layout, int H, int W, int C, const int32_t* xy, int n,
                     int Th, int Tw, int tta, int border_type, const float* border_value_host,
                     const float* scale_host, const float* bias_host, float* out, void* stream) {
  TB_REQUIRE(img && xy && out, TB_ERR_ARG, "tb_slice_augment: null pointer");
  TB_REQUIRE(H > 0 && W > 0 && C > 0 && Th > 0 && Tw > 0 && n >= 0, TB_ERR_ARG, "tb_slice_augment: bad shape");
  TB_REQUIRE(layout == TB_HWC || layout == TB_CHW, TB_ERR_ARG, "tb_slice_augment: bad layout %d", layout);
  TB_REQUIRE(border_type >= 0 && border_type <= 4, TB_ERR_ARG, "tb_slice_augment: bad border_type %d", border_type);
  SliceArgs a;
  TB_REQUIRE(make_views(tta, &a.views), TB_ERR_ARG, "tb_slice_augment: unknown tta %d", tta);
  TB_REQUIRE(!views_have_transpose(a.views) || Th == Tw, TB_ERR_ARG,
             "tb_slice_augment: D4 needs square tiles (got %dx%d)", Th, Tw);
  TB_REQUIRE(!(scale_host || bias_host) || C <= 8, TB_ERR_UNSUPPORTED,
             "tb_slice_augment: per-channel affine supports C <= 8");
  if (n == 0) return TB_OK;
  a.img = img; a.out = out; a.xy = xy; a.layout = layout;
  a.H = H; a.W = W; a.C = C; a.n = n; a.Th = Th; a.Tw = Tw; a.border = border_type;
  a.nbx = (Tw + BS - 1) / BS; a.nby = (Th + BS - 1) / BS;
  for (int c = 0; c < 4; ++c) a.bval[c] = border_value_host ? border_value_host[c] : 0.f;
  a.has_affine = (scale_host || bias_host) ? 1 : 0;
  for (int c = 0; c < 8; ++c) {
    a.scale[c] = (scale_host && c < C) ? scale_host[c] : 1.f;
    a.bias[c] = (bias_host && c < C) ? bias_host[c] : 0.f;
  }
  const long long blocks = (long long)n * a.nbx * a.nby;
  TB_REQUIRE(blocks < (1ll << 31), TB_ERR_ARG, "tb_slice_augment: grid too large");
  const bool vec = (Th % 4 == 0) && (Tw % 4 == 0) && aligned16(out);
  const bool fast = (Th % BS == 0) && (Tw % BS == 0) && aligned16(out);
  cudaStream_t s = (cudaStream_t)stream;
#define TB_SLICE(SrcT)                                                                   \
  do {                                                                                   \
    if (tta == TB_TTA_NONE && single_view_enabled() &&                                   \
        launch_slice_single_view(a, img_dtype, s) != cudaErrorNotSupported) {            \
    } else if (tta == TB_TTA_NONE && (Tw % 4 == 0) && aligned16(out) && n <= 65535) {    \
      const int bands = std::max(1, std::min(Th / 8, std::max(1, (sm_count() * 8 + n - 1) / n))); \
      dim3 g((unsigned)bands, (unsigned)n);                                              \
      slice_plain_kernel<SrcT><<<g, kThreads, 0, s>>>(a);                                \
    } else if (fast) {                                                                   \
      switch (slice_variant()) {                                                         \
        case 1: TB_DISPATCH_TTA(tta, slice_augment_fast_kernel<SrcT, TTA_, 1><<<(unsigned)blocks, kThreads, 0, s>>>(a)); break; \
        case 2: TB_DISPATCH_TTA(tta, slice_augment_fast_kernel<SrcT, TTA_, 2><<<(unsigned)blocks, kThreads, 0, s>>>(a)); break; \
        case 3: TB_DISPATCH_TTA(tta, slice_augment_fast_kernel<SrcT, TTA_, 3><<<(unsigned)blocks, kThreads, 0, s>>>(a)); break; \
        default: TB_DISPATCH_TTA(tta, slice_augment_fast_kernel<SrcT, TTA_, 0><<<(unsigned)blocks, kThreads, 0, s>>>(a)); break; \
      }                                                                                  \
    } else if (vec)                                                                      \
      slice_augment_kernel<SrcT, true><<<(unsigned)blocks, kThreads, 0, s>>>(a);         \
    else                                                                                 \
      slice_augment_kernel<SrcT, false><<<(unsigned)blocks, kThreads, 0, s>>>(a);        \
  } while (0)
  switch (img_dtype) {
    case TB_U8: TB_SLICE(unsigned char); break;
    case TB_U16: TB_SLICE(unsigned short); break;
    case TB_I16: TB_SLICE(short); break;
    case TB_F16: TB_SLICE(__half); break;
    case TB_F32: TB_SLICE(float); break;
    default: return fail(TB_ERR_UNSUPPORTED, "tb_slice_augment: unsupported image dtype %d", img_dtype);
  }
#undef TB_SLICE
  TB_LAUNCH_CHECK("slice_augment_kernel");
  return TB_OK;
}

int tb_tta_augment(const void* x, int elem_bytes, int B, int C, int H, int W, int tta, void* out, void* stream) {
  TB_REQUIRE(x && out, TB_ERR_ARG, "tb_tta_augment: null pointer");
  TB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, TB_ERR_ARG, "tb_tta_augment: bad shape");
  AugArgs a;
  TB_REQUIRE(make_views(tta, &a.views), TB_ERR_ARG, "tb_tta_augment: unknown tta %d", tta);
  TB_REQUIRE(!views_have_transpose(a.views) || H == W, TB_ERR_ARG, "tb_tta_augment: D4 needs square input (%dx%d)",
             H, W);
  if (B == 0) return TB_OK;
  a.x = x; a.out = out; a.planes = B * C; a.H = H; a.W = W;
  a.nbx = (W + BS - 1) / BS; a.nby = (H + BS - 1) / BS;
  a.view_stride = (size_t)B * C * H * W;
  const long long blocks = (long long)a.planes * a.nbx * a.nby;
  TB_REQUIRE(blocks < (1ll << 31), TB_ERR_ARG, "tb_tta_augment: grid too large");
  const bool vec = (H % 4 == 0) && (W % 4 == 0) && aligned16(out);
  cudaStream_t s = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: tta_augment_kernel<unsigned char, false><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 2: tta_augment_kernel<unsigned short, false><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 4:
      if ((H % BS == 0) && (W % BS == 0) && aligned16(x) && aligned16(out)) {
        const long long fblocks = (long long)B * a.nbx * a.nby;
        TB_DISPATCH_TTA(tta, tta_augment_fast_kernel<TTA_><<<(unsigned)fblocks, kThreads, 0, s>>>(a, B, C));
      } else if (vec)
        tta_augment_kernel<unsigned int, true><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      else
        tta_augment_kernel<unsigned int, false><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      break;
    case 8:
      if (vec)
        tta_augment_kernel<unsigned long long, true><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      else
        tta_augment_kernel<unsigned long long, false><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      break;
    default: return fail(TB_ERR_ARG, "tb_tta_augment: elem_bytes must be 1, 2, 4 or 8 (got %d)", elem_bytes);
  }
  TB_LAUNCH_CHECK("tta_augment_kernel");
  return TB_OK;
}

int tb_view_permute(const void* x, int elem_bytes, int planes, int H, int W, int view_code, void* out, void* stream) {
  TB_REQUIRE(x && out, TB_ERR_ARG, "tb_view_permute: null pointer");
  TB_REQUIRE(planes >= 0 && H > 0 && W > 0, TB_ERR_ARG, "tb_view_permute: bad shape");
  TB_REQUIRE(view_code >= 0 && view_code < 8, TB_ERR_ARG, "tb_view_permute: view_code must be in [0,8)");
  if (planes == 0) return TB_OK;
  AugArgs a;
  a.views.count = 1;
  a.views.packed = (unsigned)view_code;
  a.x = x; a.out = out; a.planes = planes; a.H = H; a.W = W;
  a.nbx = (W + BS - 1) / BS; a.nby = (H + BS - 1) / BS;
  a.view_stride = 0;
  const long long blocks = (long long)planes * a.nbx * a.nby;
  TB_REQUIRE(blocks < (1ll << 31), TB_ERR_ARG, "tb_view_permute: grid too large");
  const bool vec = (H % 4 == 0) && (W % 4 == 0) && aligned16(out);
  cudaStream_t s = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: tta_augment_kernel<unsigned char, false><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 2: tta_augment_kernel<unsigned short, false><<<(unsigned)blocks, kThreads, 0, s>>>(a); break;
    case 4:
      if (vec)
        tta_augment_kernel<unsigned int, true><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      else
        tta_augment_kernel<unsigned int, false><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      break;
    case 8:
      if (vec)
        tta_augment_kernel<unsigned long long, true><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      else
        tta_augment_kernel<unsigned long long, false><<<(unsigned)blocks, kThreads, 0, s>>>(a);
      break;
    default: return fail(TB_ERR_ARG, "tb_view_permute: elem_bytes must be 1, 2, 4 or 8 (got %d)", elem_bytes);
  }
  TB_LAUNCH_CHECK("tta_augment_kernel(view_permute)");
  return TB_OK;
}

int tb_tta_deaugment(const void* x, int dtype, int B, int C, int H, int W, int tta, int reduction, void* out,
                     void* stream) {
  TB_REQUIRE(x && out, TB_ERR_ARG, "tb_tta_deaugment: null pointer");
  TB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, TB_ERR_ARG, "tb_tta_deaugment: bad shape");
  TB_REQUIRE(reduction >= TB_RED_NONE && reduction <= TB_RED_LOG1P, TB_ERR_ARG, "tb_tta_deaugment: bad reduction %d",
             reduction);
  DeaugArgs a;
  TB_REQUIRE(make_views(tta, &a.views), TB_ERR_ARG, "tb_tta_deaugment: unknown tta %d", tta);
  TB_REQUIRE(!views_have_transpose(a.views) || H == W, TB_ERR_ARG, "tb_tta_deaugment: D4 needs square input (%dx%d)",
             H, W);
  if (B == 0) return TB_OK;
  a.x = x; a.out = out; a.planes = B * C; a.H = H; a.W = W; a.red = reduction;
  a.nbx = (W + BS - 1) / BS; a.nby = (H + BS - 1) / BS;
  a.view_stride = (size_t)B * C * H * W;
  const long long blocks = (long long)a.planes * a.nbx * a.nby;
  TB_REQUIRE(blocks < (1ll << 31), TB_ERR_ARG, "tb_tta_deaugment: grid too large");
  const bool al = (H % 4 == 0) && (W % 4 == 0) && aligned16(x) && aligned16(out);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == TB_F32 && reduction != TB_RED_NONE && al && (H % BS == 0) && (W % BS == 0) && !tb_force_generic_merge()) {
    // TMA ring: the [A*B, C, H, W] stack as the rank-3 tensor {W, B*C*H, A}
    TmaMaps<1> maps;
    if (make_pred_tensor_map((const float*)x, B, C, H, W, a.views.count, &maps.m[0])) {
      TileTmaArgs ta{};
      ta.out = (float*)out; ta.n = B; ta.K = C; ta.Th = H; ta.Tw = W; ta.nbx = W / BS; ta.nby = H / BS; ta.red = reduction;
      const long long nb = (long long)B * ta.nbx * ta.nby;
      if (nb < (1ll << 31)) {
        cudaError_t e = cudaErrorInvalidValue;
        TB_DISPATCH_TTA(tta, e = (launch_tile_deaug_tma<TTA_, false>(ta, maps, (int)nb, s)));
        if (e != cudaSuccess) return cuda_fail(e, "tile_deaug_tma_kernel");
        return TB_OK;
      }
    }
  }
  if (dtype == TB_F32) {
    if (al) {
      TB_DISPATCH_TTA(tta, tta_deaugment_kernel<float, true, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
    } else {
      TB_DISPATCH_TTA(tta, tta_deaugment_kernel<float, false, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
    }
  } else if (dtype == TB_F64) {
    if (al) {
      TB_DISPATCH_TTA(tta, tta_deaugment_kernel<double, true, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
    } else {
      TB_DISPATCH_TTA(tta, tta_deaugment_kernel<double, false, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
    }
  } else {
    return fail(TB_ERR_UNSUPPORTED, "tb_tta_deaugment: dtype must be TB_F32 or TB_F64 (got %d)", dtype);
  }
  TB_LAUNCH_CHECK("tta_deaugment_kernel");
  return TB_OK;
}

}  // extern "C"

template <typename T>
static int launch_deaugment_merge(const void* pred, const int32_t* xy, int xy_is_host, int n, int K, int Th, int Tw,
                                  int tta, ViewSet views, int reduction, const tb_weight* weight, void* image, void* norm,
                                  int Hp, int Wp, cudaStream_t s) {
  MergeArgs<T> a;
  a.pred = (const T*)pred; a.image = (T*)image; a.norm = (T*)norm;
  a.n = n; a.K = K; a.Th = Th; a.Tw = Tw; a.Hp = Hp; a.Wp = Wp; a.red = reduction; a.views = views;
  a.nbx = (Tw + BS - 1) / BS; a.nby = (Th + BS - 1) / BS;
  if (!make_weight<T>(weight, Th, Tw, &a.w)) return fail(TB_ERR_ARG, "tb_deaugment_merge: bad weight spec");
  if (xy_is_host) {
    if (n > kMaxInlineTiles)
      return fail(TB_ERR_ARG, "tb_deaugment_merge: host coordinates allow n <= %d (got %d)", kMaxInlineTiles, n);
    a.xy.dev = nullptr;
    for (int t = 0; t < n; ++t) {
      a.xy.inl[2 * t] = xy[2 * t];
      a.xy.inl[2 * t + 1] = xy[2 * t + 1];
      if (xy[2 * t] < 0 || xy[2 * t + 1] < 0 || xy[2 * t] + Tw > Wp || xy[2 * t + 1] + Th > Hp)
        return fail(TB_ERR_ARG, "tb_deaugment_merge: tile %d at (%d,%d) leaves the %dx%d canvas", t, xy[2 * t],
                    xy[2 * t + 1], Wp, Hp);
    }
  } else {
    a.xy.dev = xy;
  }
  const long long blocks = (long long)n * a.nbx * a.nby;
  if (blocks >= (1ll << 31)) return fail(TB_ERR_ARG, "tb_deaugment_merge: grid too large");
  const bool al = (Th % 4 == 0) && (Tw % 4 == 0) && aligned16(pred) && aligned16(image) &&
                  (!norm || aligned16(norm)) && (weight->kind != TB_W_TABLE || aligned16(weight->table));
  if (std::is_same<T, float>::value && al && (Th % BS == 0) && (Tw % BS == 0) && !tb_force_generic_merge()) {
    TmaMaps<1> maps;
    if (make_pred_tensor_map((const float*)pred, n, K, Th, Tw, views.count, &maps.m[0])) {
      TileTmaArgs ta{};
      ta.image = (float*)image; ta.norm = (float*)norm; ta.xy = a.xy; ta.n = n; ta.K = K; ta.Th = Th; ta.Tw = Tw;
      ta.Hp = Hp; ta.Wp = Wp; ta.nbx = Tw / BS; ta.nby = Th / BS; ta.red = reduction;
      make_weight<float>(weight, Th, Tw, &ta.w);
      cudaError_t e = cudaErrorInvalidValue;
      TB_DISPATCH_TTA(tta, e = (launch_tile_deaug_tma<TTA_, true>(ta, maps, (int)blocks, s)));
      if (e != cudaSuccess) return cuda_fail(e, "tile_deaug_tma_kernel(scatter)");
      return TB_OK;
    }
  }
  if (al) {
    TB_DISPATCH_TTA(tta, deaugment_merge_kernel<T, true, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
  } else {
    TB_DISPATCH_TTA(tta, deaugment_merge_kernel<T, false, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
  }
  TB_LAUNCH_CHECK("deaugment_merge_kernel");
  return TB_OK;
}

extern "C" {

int tb_deaugment_merge(const void* pred, int dtype, const int32_t* xy, int xy_is_host, int n, int K, int Th, int Tw,
                       int tta, int reduction, const tb_weight* weight, void* image, void* norm, int Hp, int Wp,
                       void* stream) {
  TB_REQUIRE(pred && xy && weight && image, TB_ERR_ARG, "tb_deaugment_merge: null pointer");
  TB_REQUIRE(n >= 0 && K > 0 && Th > 0 && Tw > 0 && Hp >= Th && Wp >= Tw, TB_ERR_ARG,
             "tb_deaugment_merge: bad shape");
  TB_REQUIRE(reduction >= TB_RED_MEAN && reduction <= TB_RED_LOG1P, TB_ERR_ARG,
             "tb_deaugment_merge: reduction must collapse the views (got %d)", reduction);
  ViewSet views;
  TB_REQUIRE(make_views(tta, &views), TB_ERR_ARG, "tb_deaugment_merge: unknown tta %d", tta);
  TB_REQUIRE(!views_have_transpose(views) || Th == Tw, TB_ERR_ARG, "tb_deaugment_merge: D4 needs square tiles");
  if (n == 0) return TB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == TB_F32)
    return launch_deaugment_merge<float>(pred, xy, xy_is_host, n, K, Th, Tw, tta, views, reduction, weight, image,
                                         norm, Hp, Wp, s);
  if (dtype == TB_F64)
    return launch_deaugment_merge<double>(pred, xy, xy_is_host, n, K, Th, Tw, tta, views, reduction, weight, image,
                                          norm, Hp, Wp, s);
  return fail(TB_ERR_UNSUPPORTED, "tb_deaugment_merge: dtype must be TB_F32 or TB_F64 (got %d)", dtype);
}

int tb_normalize_crop(const void* image, const void* norm, int dtype, int K, int Hp, int Wp, int top, int left,
                      int H, int W, double eps, void* out, void* stream) {
  TB_REQUIRE(image && norm && out, TB_ERR_ARG, "tb_normalize_crop: null pointer");
  TB_REQUIRE(K > 0 && H > 0 && W > 0 && top >= 0 && left >= 0 && top + H <= Hp && left + W <= Wp, TB_ERR_ARG,
             "tb_normalize_crop: crop window outside the canvas");
  const long long total = (long long)H * W;
  const int gx = (int)std::min<long long>((total + kThreads - 1) / kThreads, (long long)sm_count() * 16);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == TB_F32)
    normalize_crop_kernel<float><<<gx, kThreads, 0, s>>>((const float*)image, (const float*)norm, K, Hp, Wp, top,
                                                          left, H, W, (float)eps, eps >= 0, (float*)out);
  else if (dtype == TB_F64)
    normalize_crop_kernel<double><<<gx, kThreads, 0, s>>>((const double*)image, (const double*)norm, K, Hp, Wp, top,
                                                           left, H, W, eps, eps >= 0, (double*)out);
  else
    return fail(TB_ERR_UNSUPPORTED, "tb_normalize_crop: dtype must be TB_F32 or TB_F64 (got %d)", dtype);
  TB_LAUNCH_CHECK("normalize_crop_kernel");
  return TB_OK;
}

int tb_logits_to_labels(const float* logits, int K, int64_t plane_stride, int64_t n, int mode, float threshold,
                        uint8_t* out, int64_t out_plane_stride, void* stream) {
  TB_REQUIRE(logits && out, TB_ERR_ARG, "tb_logits_to_labels: null pointer");
  TB_REQUIRE(K > 0 && n >= 0 && plane_stride >= n, TB_ERR_ARG, "tb_logits_to_labels: bad shape");
  TB_REQUIRE(mode == 0 || mode == 1, TB_ERR_ARG, "tb_logits_to_labels: mode must be 0 (argmax) or 1 (sigmoid > threshold)");
  TB_REQUIRE(mode == 1 || K <= 256, TB_ERR_UNSUPPORTED, "tb_logits_to_labels: argmax into uint8 needs K <= 256");
  TB_REQUIRE(mode == 0 || out_plane_stride >= n, TB_ERR_ARG, "tb_logits_to_labels: bad out_plane_stride");
  if (n == 0) return TB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec4 = (n % 4 == 0) && (plane_stride % 4 == 0) && aligned16(logits) && (reinterpret_cast<uintptr_t>(out) & 3) == 0 &&
                    (mode == 0 || out_plane_stride % 4 == 0);
  const long long items = vec4 ? n / 4 : n;
  const int gx = (int)std::min<long long>((items + kThreads - 1) / kThreads, (long long)sm_count() * 16);
  if (vec4)
    logits_to_labels_kernel<4><<<gx, kThreads, 0, s>>>(logits, K, plane_stride, n, mode, threshold, out, out_plane_stride);
  else
    logits_to_labels_kernel<1><<<gx, kThreads, 0, s>>>(logits, K, plane_stride, n, mode, threshold, out, out_plane_stride);
  TB_LAUNCH_CHECK("logits_to_labels_kernel");
  return TB_OK;
}

}  // extern "C"

// pre-activation / output-mode options shared by tb_merge_all_ex and tb_merge_all_batched_ex
static int apply_merge_extra(const tb_merge_extra* ex, int raw, int K, float*& out_num, MergeAllArgs* a) {
  a->pre = TB_PRE_NONE; a->temperature = 1.f; a->out_mode = TB_OUT_LOGITS; a->thr = 0.5f; a->out_u8 = nullptr;
  if (ex) {
    TB_REQUIRE(ex->pre == TB_PRE_NONE || ex->pre == TB_PRE_SIGMOID, TB_ERR_ARG, "tb_merge_all: bad pre-activation %d", ex->pre);
    TB_REQUIRE(ex->out_mode >= TB_OUT_LOGITS && ex->out_mode <= TB_OUT_MASK, TB_ERR_ARG, "tb_merge_all: bad out_mode %d", ex->out_mode);
    a->pre = ex->pre; a->temperature = ex->temperature; a->out_mode = ex->out_mode; a->thr = ex->threshold; a->out_u8 = ex->out_u8;
    if (ex->out_mode != TB_OUT_LOGITS) {
      TB_REQUIRE(!raw, TB_ERR_ARG, "tb_merge_all: label / mask output needs the normalised (raw = 0) mode");
      TB_REQUIRE(ex->out_u8, TB_ERR_ARG, "tb_merge_all: out_mode %d needs out_u8", ex->out_mode);
      TB_REQUIRE(ex->out_mode != TB_OUT_ARGMAX || K <= 256, TB_ERR_UNSUPPORTED, "tb_merge_all: argmax into uint8 needs K <= 256");
    }
  }
  a->out_num = out_num;
  a->vec_ok = aligned16(out_num) && (!a->out_u8 || (reinterpret_cast<uintptr_t>(a->out_u8) & 3) == 0);
  a->peer = PeerPlan{};
  if (ex && ex->peer && ex->peer->phase_mask) {
    const tb_peer_plan& p = *ex->peer;
    TB_REQUIRE(!raw, TB_ERR_ARG, "tb_merge_all: the peer plan replaces the raw band mode");
    TB_REQUIRE(p.phase_mask > 0 && p.phase_mask < 16 && p.t0 >= 0 && p.t0 < p.t1, TB_ERR_ARG, "tb_merge_all: bad peer plan");
    const bool reads = p.phase_mask & (TB_PH_SEEDED_FINAL | TB_PH_SEEDED_SPILL), writes = p.phase_mask & (TB_PH_SPILL | TB_PH_SEEDED_SPILL);
    TB_REQUIRE(!reads || (p.seed_num && p.seed_den && p.seed_rows > 0 && aligned16(p.seed_num) && aligned16(p.seed_den)),
               TB_ERR_ARG, "tb_merge_all: seeded phases need 16-byte aligned seed buffers");
    TB_REQUIRE(!writes || (p.spill_num && p.spill_den && p.spill_rows > 0 && aligned16(p.spill_num) && aligned16(p.spill_den)),
               TB_ERR_ARG, "tb_merge_all: spill phases need 16-byte aligned spill buffers");
    a->peer.phase_mask = p.phase_mask; a->peer.t0 = p.t0; a->peer.t1 = p.t1;
    a->peer.seed_num = p.seed_num; a->peer.seed_den = p.seed_den; a->peer.seed_y0 = p.seed_y0; a->peer.seed_rows = p.seed_rows;
    a->peer.spill_num = p.spill_num; a->peer.spill_den = p.spill_den; a->peer.spill_y0 = p.spill_y0; a->peer.spill_rows = p.spill_rows;
    a->peer.fused = p.fused ? 1 : 0;
    if (p.fused) {
      TB_REQUIRE(p.epoch > 0 && p.counters && p.timeout_ms > 0, TB_ERR_ARG, "tb_merge_all: fused peer plan needs epoch, counters, timeout");
      TB_REQUIRE(!p.has_next || (p.next_ready && p.ack_flag && writes), TB_ERR_ARG, "tb_merge_all: fused plan with a next rank needs its flags and spill buffers");
      TB_REQUIRE(!p.has_prev || (p.prev_ack && p.ready_flag && reads), TB_ERR_ARG, "tb_merge_all: fused plan with a previous rank needs its flags and seed buffers");
      a->peer.chained = p.chained ? 1 : 0; a->peer.has_next = p.has_next ? 1 : 0; a->peer.has_prev = p.has_prev ? 1 : 0;
      a->peer.epoch = p.epoch;
      a->peer.counters = reinterpret_cast<unsigned long long*>(p.counters);
      a->peer.ready_flag = reinterpret_cast<const unsigned long long*>(p.ready_flag);
      a->peer.ack_flag = reinterpret_cast<const unsigned long long*>(p.ack_flag);
      a->peer.next_ready = reinterpret_cast<unsigned long long*>(p.next_ready);
      a->peer.prev_ack = reinterpret_cast<unsigned long long*>(p.prev_ack);
      a->peer.status = reinterpret_cast<unsigned long long*>(p.status);
      a->peer.timeout_ns = (long long)p.timeout_ms * 1000000ll;
    }
  }
  return TB_OK;
}

extern "C" {

int tb_merge_all(const float* pred, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight,
                 int Hp, int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den, int raw,
                 int top, int left, int H, int W, float* out_num, float* out_den, void* stream) {
  return tb_merge_all_ex(pred, grid, K, tta, reduction, weight, Hp, Wp, y0, y1, r0, r1, init_num, init_den, raw, top, left, H, W,
                         out_num, out_den, nullptr, stream);
}

int tb_merge_all_batched(const float* const* batch_ptrs_host, int n_batches, int batch_tiles, int last_batch_tiles,
                         int first_tile, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight, int Hp,
                         int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den, int raw, int top,
                         int left, int H, int W, float* out_num, float* out_den, void* stream) {
  return tb_merge_all_batched_ex(batch_ptrs_host, n_batches, batch_tiles, last_batch_tiles, first_tile, grid, K, tta, reduction,
                                 weight, Hp, Wp, y0, y1, r0, r1, init_num, init_den, raw, top, left, H, W, out_num, out_den, nullptr,
                                 stream);
}

int tb_merge_all_ex(const float* pred, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight,
                    int Hp, int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den, int raw,
                    int top, int left, int H, int W, float* out_num, float* out_den, const tb_merge_extra* ex, void* stream) {
  TB_REQUIRE(grid && weight && (out_num || (ex && ex->out_mode != TB_OUT_LOGITS)), TB_ERR_ARG, "tb_merge_all: null pointer");
  TB_REQUIRE(pred || grid->tile_ptr, TB_ERR_ARG, "tb_merge_all: need pred or grid->tile_ptr");
  TB_REQUIRE(!grid->tile_ptr || grid->tile_vstride, TB_ERR_ARG, "tb_merge_all: tile_ptr needs tile_vstride");
  TB_REQUIRE(K > 0 && grid->nx > 0 && grid->ny > 0 && grid->Sx > 0 && grid->Sy > 0 && grid->Th > 0 && grid->Tw > 0 &&
                 grid->n_slots > 0,
             TB_ERR_ARG, "tb_merge_all: bad grid");
  TB_REQUIRE(0 <= y0 && y0 < y1 && y1 <= Hp, TB_ERR_ARG, "tb_merge_all: bad row window [%d,%d) of %d", y0, y1, Hp);
  TB_REQUIRE(0 <= r0 && r0 <= r1 && r1 <= grid->ny, TB_ERR_ARG, "tb_merge_all: bad tile-row window");
  TB_REQUIRE((grid->nx - 1) * grid->Sx + grid->Tw <= Wp && (grid->ny - 1) * grid->Sy + grid->Th <= Hp, TB_ERR_ARG,
             "tb_merge_all: grid exceeds the canvas");
  TB_REQUIRE(reduction >= TB_RED_MEAN && reduction <= TB_RED_LOG1P, TB_ERR_ARG, "tb_merge_all: bad reduction %d",
             reduction);
  TB_REQUIRE(!raw || out_den, TB_ERR_ARG, "tb_merge_all: raw output needs out_den");
  TB_REQUIRE((init_num == nullptr) == (init_den == nullptr), TB_ERR_ARG, "tb_merge_all: init_num/init_den mismatch");
  TB_REQUIRE(raw || (top >= 0 && left >= 0 && H > 0 && W > 0 && top + H <= Hp && left + W <= Wp), TB_ERR_ARG,
             "tb_merge_all: crop window outside the canvas");
  MergeAllArgs a;
  TB_REQUIRE(make_views(tta, &a.views), TB_ERR_ARG, "tb_merge_all: unknown tta %d", tta);
  TB_REQUIRE(!views_have_transpose(a.views) || grid->Th == grid->Tw, TB_ERR_ARG, "tb_merge_all: D4 needs square tiles");
  TB_REQUIRE(make_weight<float>(weight, grid->Th, grid->Tw, &a.w), TB_ERR_ARG, "tb_merge_all: bad weight spec");
  a.pred = pred; a.grid = *grid; a.K = K; a.red = reduction; a.Hp = Hp; a.Wp = Wp;
  a.y0 = y0; a.y1 = y1; a.r0 = r0; a.r1 = r1; a.raw = raw; a.top = top; a.left = left; a.H = H; a.W = W;
  a.init_num = init_num; a.init_den = init_den; a.out_den = out_den;
  {
    const int rc = apply_merge_extra(ex, raw, K, out_num, &a);
    if (rc != TB_OK) return rc;
  }
  {
    const char* dbg = getenv("TB200_TMA_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
    if (merge_variant() & 2) a.debug |= 8;
  }
  // block window: rows from floor(y0/BS) to ceil(y1/BS); columns cover the crop (or the whole canvas when raw)
  const int xa = raw ? 0 : left, xb = raw ? Wp : left + W;
  a.bx0 = xa / BS;
  a.nbx = (xb + BS - 1) / BS - a.bx0;
  a.nby = (y1 + BS - 1) / BS - y0 / BS;
  const long long blocks = (long long)a.nbx * a.nby;
  TB_REQUIRE(blocks > 0 && blocks < (1ll << 31), TB_ERR_ARG, "tb_merge_all: grid too large");
  const bool al = (grid->Th % 4 == 0) && (grid->Tw % 4 == 0) && (grid->Sx % 4 == 0) && (grid->Sy % 4 == 0) &&
                  (!pred || aligned16(pred)) && !grid->tile_ptr && a.vec_ok &&
                  (weight->kind != TB_W_TABLE || aligned16(weight->table));
  // with a pointer table the caller vouches for 16-byte aligned planes via tile_vstride % 4 == 0 (checked by the shim)
  const bool al_tbl = grid->tile_ptr && (grid->Th % 4 == 0) && (grid->Tw % 4 == 0) && (grid->Sx % 4 == 0) &&
                      (grid->Sy % 4 == 0) && a.vec_ok &&
                      (weight->kind != TB_W_TABLE || aligned16(weight->table));
  cudaStream_t s = (cudaStream_t)stream;
  // single view (no TTA): plain-load kernel, no staging (single_view.cu)
  if (a.views.count == 1 && !a.peer.fused && al && pred && !grid->tile_slot && single_view_enabled() && !tb_force_generic_merge()) {
    a.first_tile = 0;
    a.batch_tiles = grid->n_slots;
    a.covered_tiles = grid->n_slots;
    BatchPtrs ptrs;
    ptrs.p[0] = pred;
    const cudaError_t e = launch_merge_single_view(a, ptrs, s);
    if (e == cudaSuccess) return TB_OK;
    if (e != cudaErrorNotSupported) return cuda_fail(e, "merge_single_view_kernel");
    (void)cudaGetLastError();
  }
  // TMA-pipelined path: one pred tensor, tile geometry on the 32-pixel block lattice
  if (al && pred && !grid->tile_slot && tma_geometry_ok(grid, Wp, init_num, init_den, raw, out_num, out_den, weight)) {
    TmaMaps<1> maps;
    if (make_pred_tensor_map(pred, grid->n_slots, K, grid->Th, grid->Tw, a.views.count, &maps.m[0])) {
      a.first_tile = 0;
      a.batch_tiles = grid->n_slots;
      a.covered_tiles = grid->n_slots;
      cudaError_t e = cudaErrorInvalidValue;
      TB_DISPATCH_TTA(tta, e = (launch_merge_all_tma<TTA_, 1>(a, maps, (int)blocks, s)));
      if (e != cudaSuccess) return cuda_fail(e, "merge_all_tma_kernel");
      return TB_OK;
    }
  }
  TB_REQUIRE(!a.peer.phase_mask, TB_ERR_UNSUPPORTED, "tb_merge_all: the peer plan needs the lattice kernels (Th, Tw, Sx, Sy % 32 == 0)");
  if (al || al_tbl) {
    TB_DISPATCH_TTA(tta, merge_all_kernel<true, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
  } else {
    TB_DISPATCH_TTA(tta, merge_all_kernel<false, TTA_><<<(unsigned)blocks, kThreads, 0, s>>>(a));
  }
  TB_LAUNCH_CHECK("merge_all_kernel");
  return TB_OK;
}

int tb_merge_all_batched_ex(const float* const* batch_ptrs_host, int n_batches, int batch_tiles, int last_batch_tiles,
                            int first_tile, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight, int Hp,
                            int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den, int raw, int top,
                            int left, int H, int W, float* out_num, float* out_den, const tb_merge_extra* ex, void* stream) {
  TB_REQUIRE(batch_ptrs_host && grid && weight && (out_num || (ex && ex->out_mode != TB_OUT_LOGITS)), TB_ERR_ARG,
             "tb_merge_all_batched: null pointer");
  TB_REQUIRE(n_batches > 0 && batch_tiles > 0 && last_batch_tiles > 0 && last_batch_tiles <= batch_tiles && first_tile >= 0,
             TB_ERR_ARG, "tb_merge_all_batched: bad batch description");
  TB_REQUIRE(K > 0 && grid->nx > 0 && grid->ny > 0 && grid->Sx > 0 && grid->Sy > 0 && grid->Th > 0 && grid->Tw > 0,
             TB_ERR_ARG, "tb_merge_all_batched: bad grid");
  const long long covered = (long long)(n_batches - 1) * batch_tiles + last_batch_tiles;
  TB_REQUIRE(first_tile + covered <= (long long)grid->nx * grid->ny, TB_ERR_ARG, "tb_merge_all_batched: batches exceed the grid");
  TB_REQUIRE(0 <= y0 && y0 < y1 && y1 <= Hp, TB_ERR_ARG, "tb_merge_all_batched: bad row window [%d,%d) of %d", y0, y1, Hp);
  TB_REQUIRE(0 <= r0 && r0 <= r1 && r1 <= grid->ny, TB_ERR_ARG, "tb_merge_all_batched: bad tile-row window");
  TB_REQUIRE((grid->nx - 1) * grid->Sx + grid->Tw <= Wp && (grid->ny - 1) * grid->Sy + grid->Th <= Hp, TB_ERR_ARG,
             "tb_merge_all_batched: grid exceeds the canvas");
  TB_REQUIRE(reduction >= TB_RED_MEAN && reduction <= TB_RED_LOG1P, TB_ERR_ARG, "tb_merge_all_batched: bad reduction %d", reduction);
  TB_REQUIRE(!raw || out_den, TB_ERR_ARG, "tb_merge_all_batched: raw output needs out_den");
  TB_REQUIRE((init_num == nullptr) == (init_den == nullptr), TB_ERR_ARG, "tb_merge_all_batched: init_num/init_den mismatch");
  TB_REQUIRE(raw || (top >= 0 && left >= 0 && H > 0 && W > 0 && top + H <= Hp && left + W <= Wp), TB_ERR_ARG,
             "tb_merge_all_batched: crop window outside the canvas");
  MergeAllArgs a;
  TB_REQUIRE(make_views(tta, &a.views), TB_ERR_ARG, "tb_merge_all_batched: unknown tta %d", tta);
  TB_REQUIRE(!views_have_transpose(a.views) || grid->Th == grid->Tw, TB_ERR_ARG, "tb_merge_all_batched: D4 needs square tiles");
  TB_REQUIRE(make_weight<float>(weight, grid->Th, grid->Tw, &a.w), TB_ERR_ARG, "tb_merge_all_batched: bad weight spec");
  if (n_batches > kMaxTmaBatches || !tma_geometry_ok(grid, Wp, init_num, init_den, raw, out_num, out_den, weight))
    return fail(TB_ERR_UNSUPPORTED, "tb_merge_all_batched: geometry off the 32-pixel lattice or > %d batches; use tb_merge_all",
                kMaxTmaBatches);
  for (int i = 0; i < n_batches; ++i)
    if (!batch_ptrs_host[i] || !aligned16(batch_ptrs_host[i]))
      return fail(TB_ERR_ALIGN, "tb_merge_all_batched: batch %d pointer is null or not 16-byte aligned", i);
  a.pred = nullptr; a.grid = *grid; a.K = K; a.red = reduction; a.Hp = Hp; a.Wp = Wp;
  a.grid.tile_slot = nullptr; a.grid.tile_ptr = nullptr; a.grid.tile_vstride = nullptr;
  a.y0 = y0; a.y1 = y1; a.r0 = r0; a.r1 = r1; a.raw = raw; a.top = top; a.left = left; a.H = H; a.W = W;
  a.init_num = init_num; a.init_den = init_den; a.out_den = out_den;
  {
    const int rc = apply_merge_extra(ex, raw, K, out_num, &a);
    if (rc != TB_OK) return rc;
  }
  {
    const char* dbg = getenv("TB200_TMA_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
    if (merge_variant() & 2) a.debug |= 8;
  }
  a.first_tile = first_tile; a.batch_tiles = batch_tiles; a.covered_tiles = (int)covered;
  const int xa = raw ? 0 : left, xb = raw ? Wp : left + W;
  a.bx0 = xa / BS;
  a.nbx = (xb + BS - 1) / BS - a.bx0;
  a.nby = (y1 + BS - 1) / BS - y0 / BS;
  const long long blocks = (long long)a.nbx * a.nby;
  TB_REQUIRE(blocks > 0 && blocks < (1ll << 31), TB_ERR_ARG, "tb_merge_all_batched: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaErrorInvalidValue;
  if (a.views.count == 1 && !a.peer.fused && single_view_enabled()) {
    BatchPtrs ptrs;
    for (int i = 0; i < n_batches; ++i) ptrs.p[i] = batch_ptrs_host[i];
    const int keep = a.batch_tiles;
    if (n_batches == 1) a.batch_tiles = last_batch_tiles;
    e = launch_merge_single_view(a, ptrs, s);
    if (e == cudaSuccess) return TB_OK;
    if (e != cudaErrorNotSupported) return cuda_fail(e, "merge_single_view_kernel(batched)");
    (void)cudaGetLastError();
    a.batch_tiles = keep;
    e = cudaErrorInvalidValue;
  }
  if (n_batches == 1) {
    TmaMaps<1> maps;
    if (!make_pred_tensor_map(batch_ptrs_host[0], last_batch_tiles, K, grid->Th, grid->Tw, a.views.count, &maps.m[0]))
      return fail(TB_ERR_UNSUPPORTED, "tb_merge_all_batched: cuTensorMapEncodeTiled failed");
    a.batch_tiles = last_batch_tiles;
    TB_DISPATCH_TTA(tta, e = (launch_merge_all_tma<TTA_, 1>(a, maps, (int)blocks, s)));
  } else {
    static thread_local TmaMaps<kMaxTmaBatches> maps;  // 24 KB: keep it off the stack
    for (int i = 0; i < n_batches; ++i) {
      const int b = (i == n_batches - 1) ? last_batch_tiles : batch_tiles;
      if (!make_pred_tensor_map(batch_ptrs_host[i], b, K, grid->Th, grid->Tw, a.views.count, &maps.m[i]))
        return fail(TB_ERR_UNSUPPORTED, "tb_merge_all_batched: cuTensorMapEncodeTiled failed for batch %d", i);
    }
    TB_DISPATCH_TTA(tta, e = (launch_merge_all_tma<TTA_, kMaxTmaBatches>(a, maps, (int)blocks, s)));
  }
  if (e != cudaSuccess) return cuda_fail(e, "merge_all_tma_kernel(batched)");
  return TB_OK;
}

int tb_standin_model(const float* x, int B, int Cin, int K, int Th, int Tw, const float* Wm_host, float* y,
                     void* stream) {
  TB_REQUIRE(x && y && Wm_host, TB_ERR_ARG, "tb_standin_model: null pointer");
  TB_REQUIRE(B >= 0 && Cin > 0 && Cin <= 8 && K > 0 && K * Cin <= 64 && Th > 0 && Tw > 0 && (Th * Tw) % 4 == 0 &&
                 Tw % 4 == 0,
             TB_ERR_ARG, "tb_standin_model: bad shape");
  if (B == 0) return TB_OK;
  StandinArgs a;
  a.x = x; a.y = y; a.B = B; a.Cin = Cin; a.K = K; a.Th = Th; a.Tw = Tw;
  for (int i = 0; i < K * Cin; ++i) a.Wm[i] = Wm_host[i];
  const size_t total = (size_t)B * Th * Tw / 4;
  const int gx = (int)std::min<size_t>((total + kThreads - 1) / kThreads, (size_t)sm_count() * 32);
  standin_model_kernel<<<gx, kThreads, 0, (cudaStream_t)stream>>>(a);
  TB_LAUNCH_CHECK("standin_model_kernel");
  return TB_OK;
}

}  // extern "C"
