// 3-D tiles: VolumeSlicer.split / VolumeMerger.integrate_batch / merge of the reference
// (inference/tiles_3d.py:106-135,170-211) for sm_100a.
//
//   volume_crop_kernel        split / iter_split: constant-padded sub-volume gather, native element type (:106-133)
//   volume_scatter_kernel     VolumeMerger.integrate_batch: volume[:, roi] += tile * w; norm[:, roi] += w   (:195-208)
//   volume_normalize_kernel   VolumeMerger.merge: volume / norm_mask                                        (:210-211)
//   volume_merge_all_kernel   the whole loop, output-stationary: every output voxel gathers its covering tiles in the
//                             reference's tile order (i, j, k nested, :80-96), accumulates in registers with separately
//                             rounded mul / add, divides once, writes once -- no atomics, no canvas, deterministic and
//                             bit-identical to integrate_batch fed in crops order
// All are HBM-bound gathers / scatters with W (the fastest dimension) mapped to consecutive lanes; a thread owns 4
// consecutive voxels when the geometry allows 128-bit accesses and 1 otherwise.
#include <algorithm>

#include "common.cuh"

namespace tb {

constexpr int kVT = 256;

struct VolCropArgs {
  const void* vol;
  void* out;
  const int32_t* zyx;  // [n,3] origin of every tile in the UNPADDED volume frame (may be negative / overhang)
  int D, H, W, n, d, h, w;
  unsigned long long fill;
};

template <typename E>
__global__ void __launch_bounds__(kVT) volume_crop_kernel(VolCropArgs a) {
  const int t = blockIdx.y;
  const int z0 = __ldg(a.zyx + 3 * t), y0 = __ldg(a.zyx + 3 * t + 1), x0 = __ldg(a.zyx + 3 * t + 2);
  const E* vol = static_cast<const E*>(a.vol);
  E* out = static_cast<E*>(a.out) + (size_t)t * a.d * a.h * a.w;
  const long long total = (long long)a.d * a.h * a.w;
  for (long long idx = (long long)blockIdx.x * kVT + threadIdx.x; idx < total; idx += (long long)gridDim.x * kVT) {
    const int x = (int)(idx % a.w);
    const long long r = idx / a.w;
    const int y = (int)(r % a.h), z = (int)(r / a.h);
    const int sz = z0 + z, sy = y0 + y, sx = x0 + x;
    E v = (E)a.fill;
    if (sz >= 0 && sz < a.D && sy >= 0 && sy < a.H && sx >= 0 && sx < a.W) v = vol[((size_t)sz * a.H + sy) * a.W + sx];
    out[idx] = v;
  }
}

struct VolScatterArgs {
  const float* batch;  // [n,K,d,h,w]
  const float* weight; // [d,h,w]
  float* volume;       // [K,Dp,Hp,Wp]
  float* norm;         // [Dp,Hp,Wp]
  const int32_t* zyx;  // [n,3] in the padded frame
  int n, K, d, h, w, Dp, Hp, Wp;
};

template <int VW>
__global__ void __launch_bounds__(kVT) volume_scatter_kernel(VolScatterArgs a) {
  const int t = blockIdx.y;
  const int z0 = __ldg(a.zyx + 3 * t), y0 = __ldg(a.zyx + 3 * t + 1), x0 = __ldg(a.zyx + 3 * t + 2);
  const long long tile = (long long)a.d * a.h * a.w, totalv = tile / VW;
  const size_t canvas = (size_t)a.Dp * a.Hp * a.Wp;
  const int wv = a.w / VW;
  for (long long v = (long long)blockIdx.x * kVT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kVT) {
    const int xq = (int)(v % wv);
    const long long r = v / wv;
    const int y = (int)(r % a.h), z = (int)(r / a.h);
    const size_t dst = ((size_t)(z0 + z) * a.Hp + (y0 + y)) * a.Wp + x0 + xq * VW;
    const long long src = (long long)v * VW;
    float wt[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) wt[e] = __ldg(a.weight + src + e);
    for (int k = 0; k < a.K; ++k) {
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const float val = __ldcs(a.batch + ((size_t)t * a.K + k) * tile + src + e);
        atomicAdd(a.volume + (size_t)k * canvas + dst + e, mul_rn(val, wt[e]));
      }
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) atomicAdd(a.norm + dst + e, wt[e]);
  }
}

__global__ void __launch_bounds__(kVT) volume_normalize_kernel(const float* __restrict__ volume, const float* __restrict__ norm,
                                                               int K, long long nvox, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < nvox; i += (long long)gridDim.x * kVT) {
    const float d = norm[i];
    for (int k = 0; k < K; ++k) out[(size_t)k * nvox + i] = volume[(size_t)k * nvox + i] / d;
  }
}

struct VolMergeArgs {
  const float* pred;    // [n,K,d,h,w], tiles in crops order: index (i*nh + j)*nw + k
  const float* weight;  // [d,h,w]
  float* out;           // [K,D,H,W] (cropped)
  int K, d, h, w;       // tile size
  int sd, sh, sw;       // steps
  int nd, nh, nw;       // tiles per axis
  int pd, ph, pw;       // pad_before (crop origin in the padded frame)
  int D, H, W;          // output (original volume) size
};

// covering tile index range of padded coordinate p along one axis: tiles a with a*s <= p < a*s + t, 0 <= a < n
__device__ __forceinline__ void cover_range(int p, int t, int s, int n, int& lo, int& hi) {
  lo = p - t + 1 <= 0 ? 0 : (p - t + s) / s;  // ceil((p - t + 1) / s)
  hi = min(p / s, n - 1);
}

template <int VW>
__global__ void __launch_bounds__(kVT) volume_merge_all_kernel(VolMergeArgs a) {
  const long long nvox = (long long)a.D * a.H * a.W, totalv = nvox / VW;
  const long long tile = (long long)a.d * a.h * a.w;
  const int Wv = a.W / VW;
  for (long long v = (long long)blockIdx.x * kVT + threadIdx.x; v < totalv; v += (long long)gridDim.x * kVT) {
    const int xq = (int)(v % Wv);
    const long long r = v / Wv;
    const int y = (int)(r % a.H), z = (int)(r / a.H);
    const int pz = z + a.pd, py = y + a.ph, px = xq * VW + a.pw;  // padded-frame coordinates of the first voxel
    int i0, i1, j0, j1, k0, k1;
    cover_range(pz, a.d, a.sd, a.nd, i0, i1);
    cover_range(py, a.h, a.sh, a.nh, j0, j1);
    cover_range(px, a.w, a.sw, a.nw, k0, k1);  // (VW == 4: the host guarantees all 4 voxels share the covering columns)
    for (int c = 0; c < a.K; ++c) {
      float acc[VW], den[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) acc[e] = den[e] = 0.f;
      for (int i = i0; i <= i1; ++i)
        for (int j = j0; j <= j1; ++j)
          for (int k = k0; k <= k1; ++k) {
            const long long t = ((long long)i * a.nh + j) * a.nw + k;
            const long long off = ((long long)(pz - i * a.sd) * a.h + (py - j * a.sh)) * a.w + (px - k * a.sw);
            const float* src = a.pred + (t * a.K + c) * tile + off;
            float pv[VW], wv[VW];
            if (VW == 4) {
              const float4 q = __ldcs(reinterpret_cast<const float4*>(src));
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.weight + off));
              pv[0] = q.x; pv[1 % VW] = q.y; pv[2 % VW] = q.z; pv[3 % VW] = q.w;
              wv[0] = w4.x; wv[1 % VW] = w4.y; wv[2 % VW] = w4.z; wv[3 % VW] = w4.w;
            } else {
              pv[0] = __ldcs(src);
              wv[0] = __ldg(a.weight + off);
            }
#pragma unroll
            for (int e = 0; e < VW; ++e) {
              acc[e] = add_rn(acc[e], mul_rn(pv[e], wv[e]));
              den[e] = add_rn(den[e], wv[e]);
            }
          }
      float* o = a.out + (size_t)c * nvox + (size_t)v * VW;
      if (VW == 4) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(acc[0] / den[0], acc[1 % VW] / den[1 % VW], acc[2 % VW] / den[2 % VW],
                                                          acc[3 % VW] / den[3 % VW]));
      } else {
        o[0] = acc[0] / den[0];
      }
    }
  }
}

static inline bool val16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int vol_grid(long long items) {
  return (int)std::max<long long>(1, std::min<long long>((items + kVT - 1) / kVT, (long long)sm_count() * 16));
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_volume_crop_tiles(const void* vol, int elem_bytes, int D, int H, int W, const int32_t* zyx, int n, int d, int h, int w,
                         uint64_t fill_bits, void* out, void* stream) {
  TB_REQUIRE(vol && zyx && out, TB_ERR_ARG, "tb_volume_crop_tiles: null pointer");
  TB_REQUIRE(D > 0 && H > 0 && W > 0 && d > 0 && h > 0 && w > 0 && n >= 0 && n <= 65535, TB_ERR_ARG, "tb_volume_crop_tiles: bad shape");
  if (n == 0) return TB_OK;
  VolCropArgs a{vol, out, zyx, D, H, W, n, d, h, w, (unsigned long long)fill_bits};
  const long long total = (long long)d * h * w;
  dim3 grid((unsigned)std::min<long long>((total + kVT - 1) / kVT, 2048), (unsigned)n);
  cudaStream_t s = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: volume_crop_kernel<unsigned char><<<grid, kVT, 0, s>>>(a); break;
    case 2: volume_crop_kernel<unsigned short><<<grid, kVT, 0, s>>>(a); break;
    case 4: volume_crop_kernel<unsigned int><<<grid, kVT, 0, s>>>(a); break;
    case 8: volume_crop_kernel<unsigned long long><<<grid, kVT, 0, s>>>(a); break;
    default: return fail(TB_ERR_ARG, "tb_volume_crop_tiles: elem_bytes must be 1, 2, 4 or 8 (got %d)", elem_bytes);
  }
  TB_LAUNCH_CHECK("volume_crop_kernel");
  return TB_OK;
}

int tb_volume_integrate(const float* batch, const int32_t* zyx, int n, int K, int d, int h, int w, const float* weight,
                        float* volume, float* norm, int Dp, int Hp, int Wp, void* stream) {
  TB_REQUIRE(batch && zyx && weight && volume && norm, TB_ERR_ARG, "tb_volume_integrate: null pointer");
  TB_REQUIRE(n >= 0 && n <= 65535 && K > 0 && d > 0 && h > 0 && w > 0 && Dp >= d && Hp >= h && Wp >= w, TB_ERR_ARG,
             "tb_volume_integrate: bad shape");
  if (n == 0) return TB_OK;
  VolScatterArgs a{batch, weight, volume, norm, zyx, n, K, d, h, w, Dp, Hp, Wp};
  const long long total = (long long)d * h * w;
  dim3 grid((unsigned)std::min<long long>((total + kVT - 1) / kVT, 2048), (unsigned)n);
  volume_scatter_kernel<1><<<grid, kVT, 0, (cudaStream_t)stream>>>(a);
  TB_LAUNCH_CHECK("volume_scatter_kernel");
  return TB_OK;
}

int tb_volume_normalize(const float* volume, const float* norm, int K, int64_t nvox, float* out, void* stream) {
  TB_REQUIRE(volume && norm && out && K > 0 && nvox > 0, TB_ERR_ARG, "tb_volume_normalize: bad argument");
  volume_normalize_kernel<<<vol_grid(nvox), kVT, 0, (cudaStream_t)stream>>>(volume, norm, K, nvox, out);
  TB_LAUNCH_CHECK("volume_normalize_kernel");
  return TB_OK;
}

int tb_volume_merge_all(const float* pred, const float* weight, int K, int d, int h, int w, int sd, int sh, int sw, int nd, int nh,
                        int nw, int pd, int ph, int pw, int D, int H, int W, float* out, void* stream) {
  TB_REQUIRE(pred && weight && out, TB_ERR_ARG, "tb_volume_merge_all: null pointer");
  TB_REQUIRE(K > 0 && d > 0 && h > 0 && w > 0 && sd > 0 && sh > 0 && sw > 0 && sd <= d && sh <= h && sw <= w && nd > 0 && nh > 0 &&
                 nw > 0 && pd >= 0 && ph >= 0 && pw >= 0 && D > 0 && H > 0 && W > 0,
             TB_ERR_ARG, "tb_volume_merge_all: bad geometry");
  TB_REQUIRE(pd + D <= (nd - 1) * sd + d && ph + H <= (nh - 1) * sh + h && pw + W <= (nw - 1) * sw + w, TB_ERR_ARG,
             "tb_volume_merge_all: the crop window leaves the tiled volume");
  VolMergeArgs a{pred, weight, out, K, d, h, w, sd, sh, sw, nd, nh, nw, pd, ph, pw, D, H, W};
  const long long nvox = (long long)D * H * W;
  // 128-bit path: groups of 4 voxels along W never straddle a tile edge and every address is 16-byte aligned
  const bool vec4 = (W % 4 == 0) && (w % 4 == 0) && (sw % 4 == 0) && (pw % 4 == 0) && val16(pred) && val16(weight) && val16(out);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec4) volume_merge_all_kernel<4><<<vol_grid(nvox / 4), kVT, 0, s>>>(a);
  else volume_merge_all_kernel<1><<<vol_grid(nvox), kVT, 0, s>>>(a);
  TB_LAUNCH_CHECK("volume_merge_all_kernel");
  return TB_OK;
}

}  // extern "C"
