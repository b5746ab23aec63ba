// Shared host/device helpers for the tb200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tb200.h"

namespace tb {

// ---- error plumbing ---------------------------------------------------------------------------
char* err_buf();  // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TB_REQUIRE(cond, code, ...)                  \
  do {                                               \
    if (!(cond)) return ::tb::fail(code, __VA_ARGS__); \
  } while (0)

#define TB_CUDA(call)                                         \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return ::tb::cuda_fail(_e, #call); \
  } while (0)

#define TB_LAUNCH_CHECK(name)                                    \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::tb::cuda_fail(_e, name);     \
  } while (0)

// ---- D4 views ---------------------------------------------------------------------------------
// A view is an element of the dihedral group D4 encoded in 3 bits:
//   bit0 = transpose, bit1 = flip along rows (i), bit2 = flip along columns (j)
// Augment semantics: out[i,j] = src[map(i,j)], with
//   (a,b) = transpose ? (j,i) : (i,j);   si = flip_i ? Hs-1-a : a;   sj = flip_j ? Ws-1-b : b.
// The codes below restate the forward maps of the reference views:
//   d4 (tta.py:410-420): id, rot90_cw, rot180, rot90_ccw, transpose, rot90_cw(T), rot180(T), rot90_ccw(T)
enum : unsigned {
  V_ID = 0u,
  V_T = 1u,
  V_FI = 2u,
  V_FJ = 4u,
};

struct ViewSet {
  int count;
  unsigned packed;  // 3 bits per view, view a at bits [3a, 3a+3)
};

__host__ __device__ __forceinline__ unsigned view_code(unsigned packed, int a) { return (packed >> (3 * a)) & 7u; }

// Inverse group element: non-transposing views are involutions; (T,fi,fj)^-1 = (T,fj,fi).
__host__ __device__ __forceinline__ unsigned view_inverse(unsigned code) {
  if (code & V_T) return V_T | ((code & V_FJ) ? V_FI : 0u) | ((code & V_FI) ? V_FJ : 0u);
  return code;
}

__host__ __device__ __forceinline__ void view_map(unsigned code, int i, int j, int Hs, int Ws, int& si, int& sj) {
  const int a = (code & V_T) ? j : i;
  const int b = (code & V_T) ? i : j;
  si = (code & V_FI) ? Hs - 1 - a : a;
  sj = (code & V_FJ) ? Ws - 1 - b : b;
}

inline bool make_views(int tta, ViewSet* vs) {
  auto pack = [](std::initializer_list<unsigned> codes) {
    unsigned p = 0;
    int a = 0;
    for (unsigned c : codes) p |= c << (3 * a++);
    return p;
  };
  switch (tta) {
    case TB_TTA_NONE: *vs = {1, pack({V_ID})}; return true;
    case TB_TTA_FLIPLR: *vs = {2, pack({V_ID, V_FJ})}; return true;
    case TB_TTA_FLIPUD: *vs = {2, pack({V_ID, V_FI})}; return true;
    case TB_TTA_FLIPS: *vs = {3, pack({V_ID, V_FJ, V_FI})}; return true;
    case TB_TTA_D2: *vs = {4, pack({V_ID, V_FJ, V_FI, V_FI | V_FJ})}; return true;
    case TB_TTA_D4:
      *vs = {8, pack({V_ID, V_T | V_FI, V_FI | V_FJ, V_T | V_FJ, V_T, V_FJ, V_T | V_FI | V_FJ, V_FI})};
      return true;
    default: return false;
  }
}

// ---- border index remap (OpenCV borderInterpolate semantics; the reference delegates to
//      cv2.copyMakeBorder, tiles.py:161-169,182-190). Valid range is [lo,hi). Returns -1 for CONSTANT. ----
__device__ __forceinline__ int border_remap(int p, int lo, int hi, int mode) {
  if (p >= lo && p < hi) return p;
  const int len = hi - lo;
  int q = p - lo;
  if (mode == TB_BORDER_CONSTANT || len <= 0) return -1;
  if (mode == TB_BORDER_REPLICATE) return q < 0 ? lo : hi - 1;
  if (mode == TB_BORDER_WRAP) {
    q %= len;
    if (q < 0) q += len;
    return lo + q;
  }
  // REFLECT (fedcba|abcdefgh|hgfedcb) and REFLECT_101 (gfedcb|abcdefgh|gfedcba)
  if (len == 1) return lo;
  const int delta = (mode == TB_BORDER_REFLECT_101) ? 1 : 0;
  do {
    if (q < 0)
      q = -q - 1 + delta;
    else
      q = len - 1 - (q - len) - delta;
  } while (q < 0 || q >= len);
  return lo + q;
}

// ---- tile coordinates: device array or by-value parameter pack --------------------------------
constexpr int kMaxInlineTiles = 64;
struct TileXY {
  const int32_t* dev;                  // [n,2] or nullptr
  int32_t inl[2 * kMaxInlineTiles];    // used when dev == nullptr
  __device__ __forceinline__ void get(int t, int& x, int& y) const {
    if (dev) {
      x = __ldg(dev + 2 * t);
      y = __ldg(dev + 2 * t + 1);
    } else {
      x = inl[2 * t];
      y = inl[2 * t + 1];
    }
  }
};

// ---- tile weight ------------------------------------------------------------------------------
template <typename AccT>
struct WeightDev {
  int kind;
  const AccT* table;
  AccT param;    // alpha or 1/(2 sigma^2)
  AccT hc, wc;   // pyramid: Th/2, Tw/2;  gaussian: (Th-1)/2, (Tw-1)/2
  int Th, Tw;
};

template <typename AccT>
__device__ __forceinline__ AccT weight_at(const WeightDev<AccT>& w, int i, int j) {
  switch (w.kind) {
    case TB_W_TABLE: return __ldg(w.table + (size_t)i * w.Tw + j);
    case TB_W_ONES: return AccT(1);
    case TB_W_PYRAMID: {
      // tiles.py:26-49 with axis 0 = i (extent Th), axis 1 = j (extent Tw)
      const AccT di = AccT(i) - w.hc + AccT(0.5), dj = AccT(j) - w.wc + AccT(0.5);
      const AccT dc = sqrt(di * di + dj * dj);
      const AccT il = AccT(i) + AccT(0.5), ir = AccT(i) - AccT(w.Th) + AccT(0.5);
      const AccT jl = AccT(j) + AccT(0.5), jr = AccT(j) - AccT(w.Tw) + AccT(0.5);
      const AccT dei = sqrt(min(il * il, ir * ir) + AccT(0.25));
      const AccT dej = sqrt(min(jl * jl, jr * jr) + AccT(0.25));
      const AccT de = min(dei, dej);
      return w.param * (de / (dc + de));
    }
    default: {  // TB_W_GAUSSIAN
      const AccT di = AccT(i) - w.hc, dj = AccT(j) - w.wc;
      return exp(-(di * di + dj * dj) * w.param);
    }
  }
}

template <typename AccT>
inline bool make_weight(const tb_weight* w, int Th, int Tw, WeightDev<AccT>* out) {
  if (!w) return false;
  out->kind = w->kind;
  out->table = (const AccT*)w->table;
  out->Th = Th;
  out->Tw = Tw;
  out->param = AccT(0);
  out->hc = out->wc = AccT(0);
  switch (w->kind) {
    case TB_W_TABLE: return w->table != nullptr;
    case TB_W_ONES: return true;
    case TB_W_PYRAMID:
      out->param = AccT(w->param);
      out->hc = AccT(Th * 0.5);
      out->wc = AccT(Tw * 0.5);
      return true;
    case TB_W_GAUSSIAN:
      if (!(w->param > 0)) return false;
      out->param = AccT(1.0 / (2.0 * w->param * w->param));
      out->hc = AccT((Th - 1) * 0.5);
      out->wc = AccT((Tw - 1) * 0.5);
      return true;
    default: return false;
  }
}

// ---- TTA reductions (functional.py:250-333, tta.py:63-96) -------------------------------------
// pre(): per-view transform before the sequential sum; post(): applied to sum (A = view count).
template <typename T>
__device__ __forceinline__ T red_pre(int red, T v) {
  switch (red) {
    case TB_RED_GMEAN: return log(v);
    case TB_RED_HMEAN: return T(1) / max(v, T(1e-6));
    case TB_RED_HARMONIC1P: return T(1) / (v + T(1));
    case TB_RED_LOGODD: {
      const T c = min(max(v, T(1e-6)), T(1) - T(1e-6));
      return log(c / (T(1) - c));
    }
    case TB_RED_LOG1P: return log1p(v);
    default: return v;
  }
}
template <typename T>
__device__ __forceinline__ T red_post(int red, T s, int A) {
  switch (red) {
    case TB_RED_SUM: return s;
    case TB_RED_MEAN: return s / T(A);
    case TB_RED_GMEAN: return exp(s / T(A));
    case TB_RED_HMEAN: return T(1) / max(s / T(A), T(1e-6));
    case TB_RED_HARMONIC1P: return T(1) / (s / T(A)) - T(1);
    case TB_RED_LOGODD: {
      const T e = exp(s / T(A));
      return e / (T(1) + e);
    }
    case TB_RED_LOG1P: return exp(s / T(A)) - T(1);
    default: return s;
  }
}

// IEEE add / mul without FMA contraction: the reference computes `tile * weight` and `+=` as two
// separately rounded ATen ops (tiles.py:338), so the fused kernels must not contract them.
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// Per-device caches: function attributes (opt-in shared memory), occupancy and the SM count belong to a DEVICE, and one
// process may drive several (TileMerger(device="cuda:1"), DataParallel threads). Slots are written with the same value
// by whoever gets there first, so unsynchronised concurrent initialisation is benign.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

}  // namespace tb
