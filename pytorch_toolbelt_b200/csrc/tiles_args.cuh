// Argument blocks shared by the tile kernels' translation units (tiles.cu, single_view.cu) and the launchers that
// one unit exposes to the other. Device code only; the C-ABI structs (tb_grid, tb_weight) come from include/tb200.h.
#pragma once

#include "common.cuh"

namespace tb {

constexpr int kMaxTmaBatches = 192;  // 192 x 128 B of tensor maps = 24 KB of kernel parameters (limit 32 KB)

struct SliceArgs {
  const void* img;
  float* out;
  const int32_t* xy;
  int layout, H, W, C, n, Th, Tw, border;
  int nbx, nby;  // blocks per tile
  ViewSet views;
  float bval[4];
  int has_affine;
  float scale[8], bias[8];  // first 8 channels by value; more channels -> identity (host rejects)
};


// Multi-GPU "ragged strips" (distributed/ragged.py): this rank holds the predictions of the row-major tile range [t0,t1).
// A canvas block is FINAL here iff its last covering tile (row-major order) is ours; if earlier covering tiles belong to
// lower ranks the accumulation starts from their partial sums (SEEDED: read from `seed_*`, which the previous rank filled
// through NVLink peer stores); if later covering tiles belong to higher ranks the block is a SPILL: the un-normalised
// partial sums go to `spill_*` = the next rank's seed buffers (peer memory). Tile order -- and therefore every rounding --
// is the single-GPU order, so the sharded result is bit-identical. phase_mask == 0 disables all of this.
enum : int { TB_PH_FINAL = 1, TB_PH_SPILL = 2, TB_PH_SEEDED_FINAL = 4, TB_PH_SEEDED_SPILL = 8 };
struct PeerPlan {
  int phase_mask;   // which block classes this launch processes (bits above)
  int t0, t1;
  const float* seed_num;  // [K, seed_rows, Wp]
  const float* seed_den;  // [seed_rows, Wp]
  int seed_y0, seed_rows;
  float* spill_num;       // [K, spill_rows, Wp] (peer)
  float* spill_den;       // [spill_rows, Wp]    (peer)
  int spill_y0, spill_rows;
  // FUSED mode (persistent TMA kernel only): ONE launch per step does all block classes in the order spill -> final -> seeded
  // and handles the flags itself: before the first spill store it waits for `ack_flag >= epoch - 2` (the double buffer is
  // free), after the last spill block of the whole grid it publishes `*next_ready = epoch`, before the first seeded block it
  // waits for `ready_flag >= epoch`, after the last seeded block it publishes `*prev_ack = epoch`. `counters` = 3 local words
  // that only ever grow (arrivals of the grid's pipelines; target = epoch * pipelines): no reset between launches.
  int fused, chained, has_next, has_prev;
  unsigned long long epoch;
  unsigned long long* counters;            // local [3]: spill arrivals, seeded arrivals, chained-ready arrivals
  const unsigned long long* ready_flag;    // local
  const unsigned long long* ack_flag;      // local
  unsigned long long* next_ready;          // peer (next rank's mailbox)
  unsigned long long* prev_ack;            // peer (previous rank's mailbox)
  unsigned long long* status;              // local: set when a wait timed out
  long long timeout_ns;
};

// acquire-spin on a flag with a time-out (a dead neighbour must not hang the GPU); returns false on time-out
__device__ __forceinline__ bool peer_spin_until(const unsigned long long* flag, unsigned long long value, long long timeout_ns,
                                                unsigned long long* status) {
  unsigned long long seen = 0;
  long long start, now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
    if (seen >= value) return true;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - start > timeout_ns) break;
    __nanosleep(100);
  }
  if (status) *status = seen + 1ull;
  return false;
}

// one thread per pipeline: count this pipeline's arrival; the LAST arrival of the grid publishes `value` to `flag`
__device__ __forceinline__ void peer_arrive_and_publish(unsigned long long* counter, unsigned long long target, unsigned long long* flag,
                                                        unsigned long long value) {
  __threadfence_system();
  const unsigned long long old = atomicAdd(counter, 1ull);
  if (old + 1ull == target) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
  }
}

// Block class (0 = not ours / filtered out). (ra..rb, ca..cb) = ALL covering tile rows / columns of the block.
__device__ __forceinline__ int peer_block_class(const PeerPlan& p, int nx, int ra, int rb, int ca, int cb) {
  bool own = false;
  for (int r = ra; r <= rb && !own; ++r) {
    // row r covers tile indices [r*nx+ca, r*nx+cb]: intersects [t0,t1)?
    own = (r * nx + ca < p.t1) && (r * nx + cb >= p.t0);
  }
  if (!own) return 0;
  const bool seeded = ra * nx + ca < p.t0;
  const bool spill = rb * nx + cb >= p.t1;
  const int cls = seeded ? (spill ? TB_PH_SEEDED_SPILL : TB_PH_SEEDED_FINAL) : (spill ? TB_PH_SPILL : TB_PH_FINAL);
  return (p.phase_mask & cls) ? cls : 0;
}

struct MergeAllArgs {
  const float* pred;
  tb_grid grid;
  int K, red, Hp, Wp, y0, y1, r0, r1, raw, top, left, H, W;
  int bx0, nbx, nby;  // block-column offset and block counts covering the requested window
  int debug;          // TB200_TMA_DEBUG bits (diagnostics only)
  // TMA path: tile t (grid order) lives in batch tensor (t - first_tile) / batch_tiles at position (t - first_tile) % batch_tiles
  int first_tile, batch_tiles, covered_tiles;
  ViewSet views;
  WeightDev<float> w;
  const float* init_num;
  const float* init_den;
  float* out_num;
  float* out_den;
  // activation applied to every prediction BEFORE the view reduction (ApplySigmoidTo, ensembling.py:46-68)
  int pre;            // TB_PRE_NONE | TB_PRE_SIGMOID
  float temperature;  // logits are multiplied by this first
  // what the epilogue writes instead of float logits (normalised, cropped output only)
  int out_mode;       // TB_OUT_LOGITS | TB_OUT_ARGMAX (u8 [H,W]) | TB_OUT_MASK (u8 [K,H,W], sigmoid(logit) > thr)
  float thr;
  unsigned char* out_u8;
  int vec_ok;         // out_num 16-byte / out_u8 4-byte aligned: vector stores allowed
  PeerPlan peer;
};

__device__ __forceinline__ float pre_act(int pre, float T, float v) {
  if (pre == TB_PRE_SIGMOID) return 1.f / (1.f + expf(-(v * T)));  // x.mul(T).sigmoid_()
  return v;
}

// Epilogue of the output-stationary merges for the 4 consecutive pixels (row oi, columns oj..oj+3 of the CROPPED output) a
// thread owns; called once per channel k, k ascending, by the same thread for the same pixels, so an argmax over channels
// is a running maximum in registers (first maximum wins, NaN is maximal: torch.argmax).
struct LabelState {
  float best[4];
  int arg[4];
};

__device__ __forceinline__ void merge_epilogue(const MergeAllArgs& a, int k, int oi, int oj, const float q[4], LabelState& ls) {
  if (oi < 0 || oi >= a.H) return;
  const bool vec = a.vec_ok && ((a.left & 3) == 0) && ((a.W & 3) == 0) && oj >= 0 && oj + 3 < a.W;
  if (a.out_mode == TB_OUT_LOGITS) {
    float* o = a.out_num + ((size_t)k * a.H + oi) * a.W;
    if (vec) {
      __stcs(reinterpret_cast<float4*>(o + oj), make_float4(q[0], q[1], q[2], q[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (oj + e >= 0 && oj + e < a.W) o[oj + e] = q[e];
    }
    return;
  }
  if (a.out_mode == TB_OUT_ARGMAX) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool take = (k == 0) || (q[e] > ls.best[e]) || (q[e] != q[e] && ls.best[e] == ls.best[e]);
      if (take) {
        ls.best[e] = q[e];
        ls.arg[e] = k;
      }
    }
    if (k != a.K - 1) return;
    unsigned char* o = a.out_u8 + (size_t)oi * a.W;
    if (vec) {
      *reinterpret_cast<uchar4*>(o + oj) = make_uchar4((unsigned char)ls.arg[0], (unsigned char)ls.arg[1], (unsigned char)ls.arg[2],
                                                        (unsigned char)ls.arg[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (oj + e >= 0 && oj + e < a.W) o[oj + e] = (unsigned char)ls.arg[e];
    }
    return;
  }
  // TB_OUT_MASK
  unsigned char m[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) m[e] = (1.f / (1.f + expf(-q[e]))) > a.thr ? 1 : 0;
  unsigned char* o = a.out_u8 + ((size_t)k * a.H + oi) * a.W;
  if (vec) {
    *reinterpret_cast<uchar4*>(o + oj) = make_uchar4(m[0], m[1], m[2], m[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (oj + e >= 0 && oj + e < a.W) o[oj + e] = m[e];
  }
}


// Base pointers of the per-batch prediction tensors (tb_merge_all_batched) for kernels that use plain loads.
struct BatchPtrs {
  const float* p[kMaxTmaBatches];
};

// single_view.cu: specialised kernels for TTA_NONE (one view: nothing to permute, nothing to stage).
// Return cudaErrorNotSupported when the geometry is off their lattice (the caller then takes the generic kernels).
cudaError_t launch_merge_single_view(const MergeAllArgs& a, const BatchPtrs& ptrs, cudaStream_t s);
cudaError_t launch_slice_single_view(const SliceArgs& a, int img_dtype, cudaStream_t s);

}  // namespace tb
