// TMA + mbarrier pipelined variant of the output-stationary merge (sm_100a).
//
// The generic merge_all_kernel is latency-bound: each CTA alternates "issue 8 loads / wait / reduce" and
// its register footprint caps occupancy at 2 CTAs per SM. Here a producer lane streams the 32x32 fp32 boxes
// of all A views of a (channel, tile) pair into a ring of shared-memory stages with cp.async.bulk.tensor
// (TMA); 8 consumer warps reduce stage s while stages s+1.. are in flight. Bytes in flight are set by the
// ring depth (3 x 32 KB per CTA for D4), not by occupancy, and the load side costs one instruction per 4 KB.
//
// Shared-memory boxes use the TMA 128-byte swizzle (row = 32 floats = 128 B): 16-byte chunk j of row r
// lives at chunk j ^ (r & 7). Consumer lanes are mapped (lane>>2 -> row&7, lane&3 -> column&3) so that
// row-wise (non-transposing views) reads are bank-conflict free and column-wise (transposing views) reads
// are at most 2-way conflicted.
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace tb {

constexpr int kTmaConsumerWarps = 8;
// One producer warp (one elected lane). More producers are NOT safe with a parity-based ring unless their count is
// <= the ring depth: a producer that skips more than one ring revolution between its own waits aliases the mbarrier
// phase parity (found by compute-sanitizer with 4 producers on the 3-stage ring) -- and they bought no speed.
constexpr int kTmaProducerWarps = 1;
constexpr int kTmaThreads = (kTmaConsumerWarps + kTmaProducerWarps) * 32;
constexpr int kBoxFloats = 32 * 32;
constexpr int kBoxBytes = kBoxFloats * 4;

// A stage holds up to tma_group(views) work items (covering tiles of one canvas block in the merge, channels of one
// tile patch in the de-augment), each with all its views: 6-8 boxes = 24-32 KB per stage whatever the TTA group, so
// that few-view configurations move as many bytes per barrier round trip as D4 does.
__host__ __device__ constexpr int tma_group(int views) { return views >= 8 ? 1 : 8 / views; }
__host__ __device__ constexpr int tma_stage_boxes(int views) { return tma_group(views) * views; }
// deep = one CTA per SM with a ring that fills the 227 KB of shared memory (7 x 32 KB); else 2 CTAs x 3 stages.
__host__ __device__ constexpr int tma_stages(int views, bool deep = false) {
  return (deep ? 56 : 24) / tma_stage_boxes(views);  // 7 x 32 KB + barriers + 1 KB alignment slack <= 227 KB
}
__host__ __device__ constexpr size_t tma_smem_bytes(int views, bool deep = false) {
  return (size_t)tma_stages(views, deep) * tma_stage_boxes(views) * kBoxBytes + 2 * (size_t)tma_stages(views, deep) * sizeof(uint64_t) +
         1024 /* the CTA's dynamic window is not always 1024-aligned: aligned in-kernel */ + 64;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Warp-level stage release: every lane's loaded values are consumed by the asm (so the loads have completed), the
// warp converges, then one lane arrives.
// The slot was read through the generic proxy and will be refilled through the async proxy (TMA): the reading threads
// themselves issue the cross-proxy fence, then the warp converges and one lane arrives. (Putting the fence in the
// producer instead also hides the race, but there it waits for the producer's own outstanding TMA loads and
// serialises the whole ring to one stage per DRAM round trip.)
__device__ __forceinline__ void mbar_release_after(uint64_t* bar, float a, float b, float c, float d, bool elected,
                                                   bool skip_fence = false) {
  asm volatile("" ::"f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  if (!skip_fence) fence_proxy_async();
  __syncwarp();
  if (elected) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 3-D tiled TMA load: box (32 cols x 32 rows x 1 view) at (col, row, view) -> shared memory, completion on `bar`.
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* map, uint64_t* bar, int col, int row, int view) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(row), "r"(view)
      : "memory");
}

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_box_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int col, int row, int view,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(row), "r"(view), "l"(policy)
      : "memory");
}

// Tensor maps of the prediction tensors, passed by value in the kernel parameters (__grid_constant__). One map per
// batch tensor [A*b, K, Th, Tw] viewed as the rank-3 tensor {Tw, b*K*Th, A} (view stride = b*K*Th*Tw elements).
template <int MAXB>
struct TmaMaps {
  CUtensorMap m[MAXB];
};

// float index of element (rr, cc) inside a 128B-swizzled 32x32 fp32 box
__device__ __forceinline__ int swz(int rr, int cc) { return rr * 32 + ((((cc >> 2) ^ (rr & 7)) << 2) | (cc & 3)); }

struct MergeAllArgs;  // defined in tiles.cu

}  // namespace tb
