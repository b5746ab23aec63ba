/*
 * tb200.h -- C-ABI boundary of the B200 (sm_100a) tiled-inference / TTA / segmentation-loss engine.
 *
 * The reference (BloodAxe/pytorch-toolbelt @ 05e07ac, v0.8.0) is 100 % Python: it has no FFI. Its
 * boundary for this path is the Python API of
 *     pytorch_toolbelt/inference/tiles.py      (ImageSlicer :53, TileMerger :290)
 *     pytorch_toolbelt/inference/tta.py        (*_image_augment / *_image_deaugment :257-524)
 *     pytorch_toolbelt/inference/functional.py (flip / rot90 / transpose primitives :47-145, means :250-333)
 *     pytorch_toolbelt/losses/functional.py    (focal :20, softmax focal :110, jaccard :188, dice :221, wing :250)
 *     pytorch_toolbelt/losses/lovasz.py        (:23-160)
 * Every entry point below names the reference lines whose ATen / NumPy / OpenCV work it replaces. The
 * binding a maintainer would add is a ctypes stub; pytorch_toolbelt_b200/_native.py
 * is that stub.
 *
 * Conventions
 *   - every function returns int: 0 = OK, <0 = argument error (TB_ERR_*), >0 = cudaError_t;
 *     tb_last_error() returns a thread-local message for the last non-zero return on this thread;
 *   - all data pointers are DEVICE pointers borrowed for the duration of the call unless the parameter
 *     name ends in _host; the library never allocates, frees or retains caller-visible memory;
 *   - every launch is asynchronous on the cudaStream_t passed as `stream` (void* here so that the header
 *     needs no CUDA include) and does no implicit synchronisation;
 *   - the caller selects the device (cudaSetDevice) before the call;
 *   - no global mutable state besides one-time kernel-attribute setup: safe from several host threads.
 */
#ifndef TB200_H_
#define TB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define TB_API __attribute__((visibility("default")))
#else
#define TB_API
#endif

/* ---- error codes (negative); positive values are cudaError_t ---- */
#define TB_OK 0
#define TB_ERR_ARG (-1)         /* bad argument value */
#define TB_ERR_UNSUPPORTED (-2) /* combination not implemented by this build */
#define TB_ERR_ALIGN (-3)       /* pointer / pitch alignment requirement violated */
#define TB_ERR_WORKSPACE (-4)   /* caller workspace too small */

/* ---- element types ---- */
enum tb_dtype {
  TB_U8 = 0,
  TB_I8 = 1,
  TB_U16 = 2,
  TB_I16 = 3,
  TB_I32 = 4,
  TB_I64 = 5,
  TB_F16 = 6,
  TB_BF16 = 7,
  TB_F32 = 8,
  TB_F64 = 9
};

enum tb_layout { TB_HWC = 0, TB_CHW = 1 };

/* TTA groups, views in the order the reference concatenates them (tta.py:269,284,333-341,410-420,484). */
enum tb_tta {
  TB_TTA_NONE = 0,   /* 1 view  */
  TB_TTA_FLIPLR = 1, /* 2 views: id, lr                 tta.py:257-269 */
  TB_TTA_FLIPUD = 2, /* 2 views: id, ud                 tta.py:272-284 */
  TB_TTA_FLIPS = 3,  /* 3 views: id, lr, ud             tta.py:470-484 */
  TB_TTA_D2 = 4,     /* 4 views: id, lr, ud, rot180     tta.py:319-341 */
  TB_TTA_D4 = 5      /* 8 views (square tiles only)     tta.py:385-422 */
};

/* Same numeric values as OpenCV's cv2.BORDER_* (the reference forwards them to cv2.copyMakeBorder,
 * tiles.py:161-169,182-190). */
enum tb_border {
  TB_BORDER_CONSTANT = 0,
  TB_BORDER_REPLICATE = 1,
  TB_BORDER_REFLECT = 2,
  TB_BORDER_WRAP = 3,
  TB_BORDER_REFLECT_101 = 4
};

/* _deaugment_averaging modes, tta.py:63-96 and functional.py:250-333. */
enum tb_reduction {
  TB_RED_NONE = 0, /* keep the A de-augmented views: out is [A,B,C,H,W] */
  TB_RED_MEAN = 1,
  TB_RED_SUM = 2,
  TB_RED_GMEAN = 3,
  TB_RED_HMEAN = 4,
  TB_RED_HARMONIC1P = 5,
  TB_RED_LOGODD = 6,
  TB_RED_LOG1P = 7
};

/* Tile weight. TABLE reads the caller's [Th,Tw] array (bit-exact with TileMerger.weight, tiles.py:306);
 * the analytic kinds regenerate the weight in registers (no memory traffic for it). */
enum tb_weight_kind {
  TB_W_TABLE = 0,
  TB_W_ONES = 1,     /* "mean", tiles.py:282-283 */
  TB_W_PYRAMID = 2,  /* compute_pyramid_patch_weight_loss, tiles.py:16-50; param = alpha (host float64 sum) */
  TB_W_GAUSSIAN = 3  /* exp(-((i-(Th-1)/2)^2+(j-(Tw-1)/2)^2)/(2 sigma^2)); param = sigma (builder-defined) */
};

typedef struct tb_weight {
  int kind;           /* enum tb_weight_kind */
  const void* table;  /* device [Th,Tw], f32 (f64 for the f64 accumulator); TB_W_TABLE only */
  double param;       /* alpha (PYRAMID) or sigma (GAUSSIAN) */
} tb_weight;

TB_API int tb_version(void);
TB_API const char* tb_last_error(void);
/* Number of views of a TTA group, or TB_ERR_ARG. */
TB_API int tb_tta_views(int tta);

/* ------------------------------------------------------------------------------------------------
 * TILES
 * ---------------------------------------------------------------------------------------------- */

/* Crop n tiles (native dtype, HWC) out of an HWC image with border synthesis.
 * Replaces ImageSlicer.split / iter_split / cut_patch (tiles.py:144-234): NumPy slicing + cv2.copyMakeBorder.
 *   img        [H,W,C] elements of elem_bytes (1,2,4,8)
 *   xy         device int32 [n,2] = (x,y) of each tile in IMAGE frame (bbox_crops[:, :2]; may be negative)
 *   clip_to_tile 0: border indices are remapped against the whole image (split semantics, tiles.py:182-190)
 *                1: against the clipped tile extent (iter_split/cut_patch semantics, tiles.py:154-169)
 *   border_value_host  4 x 8 bytes: per-channel raw element bits for channels 0..3 (cv2 Scalar); channels >=4 get 0
 *   out        [n,Th,Tw,C] same element type
 */
TB_API int tb_crop_tiles(const void* img, int elem_bytes, int H, int W, int C, const int32_t* xy, int n, int Th,
                  int Tw, int border_type, const uint64_t* border_value_host, int clip_to_tile, void* out,
                  void* stream);

/* Fused crop + HWC->CHW + cast to f32 (+ optional per-channel affine) + TTA augment.
 * Replaces, per batch: ImageSlicer.split (tiles.py:177-204) -> image_to_tensor (utils/torch_utils.py:204-231)
 * -> .float() -> tta.*_image_augment (tta.py:257-284,319-341,385-422,470-484).
 *   img        [H,W,C] (TB_HWC) or [C,H,W] (TB_CHW); dtype TB_U8, TB_U16, TB_I16, TB_F16, TB_F32
 *   xy         device int32 [n,2] image-frame (x,y)
 *   scale_host / bias_host  NULL or C floats: out = float(src)*scale[c] + bias[c] (computed as mul then add)
 *   out        f32 [A*n, C, Th, Tw], aug-major: row a*n + t  (torch.cat(dim=0) order)
 */
TB_API int tb_slice_augment(const void* img, int img_dtype, int layout, int H, int W, int C, const int32_t* xy,
                     int n, int Th, int Tw, int tta, int border_type, const float* border_value_host /*[4]*/,
                     const float* scale_host, const float* bias_host, float* out, void* stream);

/* Stand-alone TTA augment on [B,C,H,W] of any 1/2/4/8-byte element type (pure index permutation).
 * Replaces tta.*_image_augment (tta.py:257-284,319-341,385-422,470-484) = flips/rot90s + torch.cat.
 * out [A*B,C,H,W]. D4 requires H == W (tta.py:403-407). */
TB_API int tb_tta_augment(const void* x, int elem_bytes, int B, int C, int H, int W, int tta, void* out, void* stream);

/* One D4 element applied to `planes` planes of [H,W]: out[p,i,j] = x[p, map(i,j)] with
 * view_code bit0 = transpose, bit1 = flip rows, bit2 = flip columns ((a,b) = transpose ? (j,i) : (i,j);
 * si = flip_rows ? H-1-a : a; sj = flip_cols ? W-1-b : b). Transposing codes produce [W,H] planes.
 * Replaces the primitives torch_fliplr/flipud/rot90_cw/rot90_ccw/rot180/transpose and their compositions
 * (inference/functional.py:47-145). */
TB_API int tb_view_permute(const void* x, int elem_bytes, int planes, int H, int W, int view_code, void* out,
                           void* stream);

/* Stand-alone TTA de-augment + reduction. Replaces tta.*_image_deaugment (tta.py:287-316,344-365,442-467,
 * 503-524): chunk -> inverse flips/rot90s -> torch.stack -> _deaugment_averaging (tta.py:63-96).
 *   x [A*B,C,H,W] (dtype TB_F32 or TB_F64); out [B,C,H,W], or [A,B,C,H,W] for TB_RED_NONE.
 * Views are summed sequentially in view order 0..A-1 and MEAN divides by A (torch CPU semantics). */
TB_API int tb_tta_deaugment(const void* x, int dtype, int B, int C, int H, int W, int tta, int reduction, void* out,
                     void* stream);

/* Streaming weighted scatter-accumulate of n tile predictions into the canvas, with the TTA de-augment
 * fused in (tta = TB_TTA_NONE for plain TileMerger.integrate_batch).
 * Replaces TileMerger.accumulate_single / integrate_batch (tiles.py:310-339) [+ tta.*_image_deaugment].
 *   pred   [A*n,K,Th,Tw] dtype TB_F32 (accumulator f32) or TB_F64 (accumulator f64)
 *   xy     int32 [n,2] canvas-frame (x,y) = crops[:, :2]; device pointer, or host pointer when
 *          xy_is_host != 0 (n <= 64: passed by value in the kernel parameters, no copy)
 *   image  [K,Hp,Wp], norm [Hp,Wp] (may be NULL): accumulated with float reductions (red.global.add);
 *          tiles must lie inside the canvas
 */
TB_API int tb_deaugment_merge(const void* pred, int dtype, const int32_t* xy, int xy_is_host, int n, int K, int Th,
                       int Tw, int tta, int reduction, const tb_weight* weight, void* image, void* norm,
                       int Hp, int Wp, void* stream);

/* merge(): out[k,i,j] = image[k,top+i,left+j] / norm[top+i,left+j]   (TileMerger.merge, tiles.py:345-350,
 * fused with crop_to_orignal_size, tiles.py:271-280). dtype TB_F32 / TB_F64. out may alias image when
 * top = left = 0, H = Hp, W = Wp (merge_). eps < 0: plain division (TileMerger); eps >= 0: the norm is
 * clipped to eps first (ImageSlicer.merge, tiles.py:266). */
TB_API int tb_normalize_crop(const void* image, const void* norm, int dtype, int K, int Hp, int Wp, int top, int left,
                      int H, int W, double eps, void* out, void* stream);

/* SURVEY 8f rank 1, output side: class map / binary masks straight from the merged logits, so that one byte per
 * pixel (instead of 4*K) crosses PCIe. Replaces `merged.argmax(dim=0).byte()` / `(merged.sigmoid() > thr).byte()`
 * of the README loop (README.md:225; utils/torch_utils.py:144-159 to_numpy).
 *   logits f32, K planes of n pixels, plane k at logits + k*plane_stride (so a band of rows of a [K,H,W] tensor works)
 *   mode 0: out u8 [n]    = argmax over k (first maximum wins, NaN is maximal: torch.argmax); K <= 256
 *   mode 1: out u8 [K][n] = sigmoid(logit) > threshold, plane k at out + k*out_plane_stride
 * Validated on a B200 in round 2 (tests/test_joint_labels_gpu.py). */
TB_API int tb_logits_to_labels(const float* logits, int K, int64_t plane_stride, int64_t n, int mode, float threshold,
                        uint8_t* out, int64_t out_plane_stride, void* stream);

/* Regular tile grid: tile (r,c), r < ny, c < nx sits at canvas (x,y) = (c*Sx, r*Sy). Where its A*K
 * prediction planes live:
 *   default            rows a*n_slots + slot of `pred`, slot = tile_slot ? tile_slot[r*nx+c] : r*nx+c
 *                      (slot < 0: tile absent);
 *   tile_ptr != NULL   view 0 of tile t = r*nx+c starts at tile_ptr[t] ([K,Th,Tw] planes, f32; NULL: absent) and
 *                      view a is tile_vstride[t] elements further per view: lets the caller merge the
 *                      model's per-batch outputs ([A*b,K,Th,Tw] each, separate allocations) with no concat copy. */
typedef struct tb_grid {
  int nx, ny, Sx, Sy, Th, Tw;
  int n_slots;                      /* tiles per view in pred (the aug-major stride) */
  const int32_t* tile_slot;         /* device int32 [ny*nx] or NULL */
  const float* const* tile_ptr;     /* device pointer table [ny*nx] or NULL */
  const int64_t* tile_vstride;      /* device int64 [ny*nx]; required with tile_ptr */
} tb_grid;

/* Output-stationary fused de-augment + weighted merge + normalise + crop over a whole (or a row band of
 * a) regular grid. Every CTA owns one canvas block, gathers the <= ceil(Th/Sy)*ceil(Tw/Sx) covering tile
 * predictions in the reference's tile order (row-major over crops, tiles.py:132-139,337-339), accumulates
 * in registers and writes each output pixel once: no atomics, no canvas read-modify-write, no norm_mask.
 * Replaces the whole loop tta.*_image_deaugment -> TileMerger.integrate_batch -> merge -> crop
 * (tta.py:442-467, tiles.py:321-350,271-280).
 *   pred       f32 [A*n_slots,K,Th,Tw] (may be NULL when grid->tile_ptr is given)
 *   rows       canvas rows [y0,y1) to produce (0,Hp for everything)
 *   tile_rows  only tiles with r in [r0,r1) contribute (0,ny for everything)
 *   init_num / init_den   NULL, or f32 [K,y1-y0,Wp] / [y1-y0,Wp] partial sums the accumulation starts from
 *                         (halo received from the rank that owns the earlier tile rows)
 *   raw != 0:  write un-normalised sums: out_num f32 [K,y1-y0,Wp], out_den f32 [y1-y0,Wp]; crop ignored
 *   raw == 0:  write out f32 [K,H,W] rows (y-top) for y in [y0,y1) ∩ [top,top+H): out_num = out, out_den unused
 */
TB_API int tb_merge_all(const float* pred, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight,
                 int Hp, int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den,
                 int raw, int top, int left, int H, int W, float* out_num, float* out_den, void* stream);

/* Same as tb_merge_all for predictions that arrive as the model's per-batch outputs (separate allocations):
 * batch i is f32 [A*b_i, K, Th, Tw] (aug-major inside the batch) holding grid tiles
 * [first_tile + i*batch_tiles, first_tile + i*batch_tiles + b_i), b_i = batch_tiles except b_last = last_batch_tiles.
 * n_batches = 1 with first_tile = r0*nx is the "only this strip's tiles are resident" case of the multi-GPU path.
 * Always takes the TMA-pipelined kernel (one rank-3 tensor map per batch in the kernel parameters, <= 192 batches);
 * returns TB_ERR_UNSUPPORTED when the geometry is off the 32-pixel lattice (Th, Tw, Sx, Sy multiples of 32) so that
 * the caller can fall back to tb_merge_all with a pointer table. batch_ptrs_host is a HOST array of device pointers. */
TB_API int tb_merge_all_batched(const float* const* batch_ptrs_host, int n_batches, int batch_tiles, int last_batch_tiles,
                                int first_tile, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight,
                                int Hp, int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den,
                                int raw, int top, int left, int H, int W, float* out_num, float* out_den, void* stream);

/* Options on top of tb_merge_all / tb_merge_all_batched; NULL = none of them.
 *   pre / temperature   activation applied to every prediction BEFORE the view reduction, as the reference's
 *                       ApplySigmoidTo(model, temperature) wrapper does in front of the TTA mean
 *                       (inference/ensembling.py:46-68: output.mul(temperature).sigmoid_())
 *   out_mode            what the epilogue writes (raw == 0 only):
 *                         TB_OUT_LOGITS  f32 [K,H,W] into out_num (the default)
 *                         TB_OUT_ARGMAX  u8 [H,W] into out_u8 = merged.argmax(dim=0) (first maximum wins, NaN maximal; K <= 256)
 *                         TB_OUT_MASK    u8 [K,H,W] into out_u8 = merged.sigmoid() > threshold
 *                       replaces `merged.argmax(0).byte()` / `(merged.sigmoid() > thr).byte()` after TileMerger.merge
 *                       (README.md:225, utils/torch_utils.py:144-159): no float canvas is written at all and one byte per
 *                       pixel leaves the GPU; out_num may be NULL in these modes. */
enum tb_pre_activation { TB_PRE_NONE = 0, TB_PRE_SIGMOID = 1 };
enum tb_out_mode { TB_OUT_LOGITS = 0, TB_OUT_ARGMAX = 1, TB_OUT_MASK = 2 };

/* Multi-GPU "ragged strips" (new relative to the reference, whose TileMerger is single-device, tiles.py:306-308):
 * the rank holds the predictions of the row-major tile range [t0,t1) (= the batched call's first_tile / covered tiles).
 * A 32x32 canvas block whose LAST covering tile is ours is FINAL here; if earlier covering tiles belong to lower ranks its
 * accumulation starts from `seed_*` (partial sums the previous rank stored there through NVLink); if later covering tiles
 * belong to higher ranks the block is a SPILL and its un-normalised sums are stored to `spill_*` (= the next rank's seed
 * buffers, peer memory) instead of any output. Row windows: buffer row 0 is canvas row seed_y0 / spill_y0, all Wp columns.
 * phase_mask selects which block classes one launch processes, so that the caller can order them around the flags:
 *   1 final unseeded | 2 spill unseeded | 4 final seeded | 8 spill seeded        (tb_peer_signal / tb_peer_wait below)
 * Only the lattice kernels (TMA ring, single view) implement it: TB_ERR_UNSUPPORTED otherwise. */
typedef struct tb_peer_plan {
  int phase_mask;
  int t0, t1;
  const float* seed_num; /* f32 [K, seed_rows, Wp] */
  const float* seed_den; /* f32 [seed_rows, Wp] */
  int seed_y0, seed_rows;
  float* spill_num;      /* f32 [K, spill_rows, Wp] */
  float* spill_den;      /* f32 [spill_rows, Wp] */
  int spill_y0, spill_rows;
  /* fused != 0 (TMA kernel, i.e. TTA groups with >= 2 views): ONE launch per step handles all classes in phase_mask in the order
   * spill -> final -> seeded AND the flags: waits `*ack_flag >= epoch - 2` before its first spill store, publishes
   * `*next_ready = epoch` after the last spill block of the grid (after the seeded pass when `chained`), waits
   * `*ready_flag >= epoch` before its first seeded block, publishes `*prev_ack = epoch` after the last one. counters: 3 local
   * uint64 words, zero-initialised once, never reset (they count arrivals; target = epoch * pipelines). epoch starts at 1. */
  int fused, chained, has_next, has_prev;
  uint64_t epoch;
  uint64_t* counters;
  const uint64_t* ready_flag;
  const uint64_t* ack_flag;
  uint64_t* next_ready;
  uint64_t* prev_ack;
  uint64_t* status;
  int timeout_ms;
} tb_peer_plan;

typedef struct tb_merge_extra {
  int pre;
  float temperature;
  int out_mode;
  float threshold;
  uint8_t* out_u8;
  const tb_peer_plan* peer; /* NULL: single-GPU / legacy band mode */
} tb_merge_extra;

TB_API int tb_merge_all_ex(const float* pred, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight,
                    int Hp, int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den,
                    int raw, int top, int left, int H, int W, float* out_num, float* out_den, const tb_merge_extra* ex,
                    void* stream);
TB_API int tb_merge_all_batched_ex(const float* const* batch_ptrs_host, int n_batches, int batch_tiles, int last_batch_tiles,
                                   int first_tile, const tb_grid* grid, int K, int tta, int reduction, const tb_weight* weight,
                                   int Hp, int Wp, int y0, int y1, int r0, int r1, const float* init_num, const float* init_den,
                                   int raw, int top, int left, int H, int W, float* out_num, float* out_den,
                                   const tb_merge_extra* ex, void* stream);

/* SURVEY 8f rank 2, model ensembles (inference/ensembling.py:71-123): out[i] = reduction over m of inputs[m][i] for M
 * separately allocated f32 tensors of n elements (inputs_host: HOST array of M <= 64 device pointers). One pass, no [M,...]
 * stack; reductions and summation order (m = 0..M-1) as in tb_tta_deaugment / _deaugment_averaging (tta.py:63-96). */
TB_API int tb_stack_reduce(const float* const* inputs_host, int M, int64_t n, int reduction, float* out, void* stream);
/* ApplySigmoidTo / ApplySoftmaxTo (inference/ensembling.py:12-68) in place on f32 [B,C,inner]:
 * mode 0: x.mul(temperature).sigmoid_();  mode 1: x.mul(temperature).softmax(dim=1). */
TB_API int tb_activation_inplace(float* x, int64_t B, int C, int64_t inner, int mode, float temperature, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3-D TILES  (reference inference/tiles_3d.py)
 * Volumes are [D,H,W] (W fastest), tiles [d,h,w], origins int32 [n,3] = (z,y,x).
 * ---------------------------------------------------------------------------------------------- */
/* VolumeSlicer.split / iter_split (tiles_3d.py:106-133): constant-padded sub-volume gather in the native element type
 * (elem_bytes 1/2/4/8; fill_bits = the padding value's bit pattern). zyx are origins in the UNPADDED volume frame
 * (crop origin - pad_before; may be negative / overhang). out [n,d,h,w]. */
TB_API int tb_volume_crop_tiles(const void* vol, int elem_bytes, int D, int H, int W, const int32_t* zyx, int n, int d, int h,
                         int w, uint64_t fill_bits, void* out, void* stream);
/* VolumeMerger.integrate_batch (tiles_3d.py:195-208): volume[:, roi] += tile * weight; norm[:, roi] += weight for the n
 * tiles of batch f32 [n,K,d,h,w]; zyx in the padded frame; float reductions in memory (order-nondeterministic rounding,
 * like any parallel scatter; tb_volume_merge_all is the deterministic form). */
TB_API int tb_volume_integrate(const float* batch, const int32_t* zyx, int n, int K, int d, int h, int w, const float* weight,
                        float* volume, float* norm, int Dp, int Hp, int Wp, void* stream);
/* VolumeMerger.merge (tiles_3d.py:210-211): out = volume / norm (no eps), nvox = Dp*Hp*Wp. */
TB_API int tb_volume_normalize(const float* volume, const float* norm, int K, int64_t nvox, float* out, void* stream);
/* The whole loop integrate_batch(all tiles, crops order) -> merge -> crop_to_orignal_size, output-stationary: pred f32
 * [nd*nh*nw, K, d, h, w] with tile (i,j,k) at index (i*nh + j)*nw + k and padded origin (i*sd, j*sh, k*sw) (tiles_3d.py:76-96);
 * out f32 [K,D,H,W] = the [pd:pd+D, ph:ph+H, pw:pw+W] window of the merged padded volume. Deterministic; bit-identical to
 * integrate_batch fed in crops order followed by merge. */
TB_API int tb_volume_merge_all(const float* pred, const float* weight, int K, int d, int h, int w, int sd, int sh, int sw, int nd,
                        int nh, int nw, int pd, int ph, int pw, int D, int H, int W, float* out, void* stream);

/* ---- peer memory for the multi-GPU halo: the one place where the library owns device memory (cudaMalloc'ed so that it
 * can be exported with CUDA IPC; torch's caching allocator sub-allocates and cannot). One process per GPU; a rank
 * allocates its mailbox (seed buffers + flags), exports the 64-byte handle, the neighbour opens it and gets a pointer
 * it can store to from its kernels over NVLink (or into the same GPU when two ranks share one in tests).
 *   tb_peer_alloc   cudaMalloc + zero fill;  tb_peer_free
 *   tb_peer_export  handle_out: 64 bytes (cudaIpcMemHandle_t);  tb_peer_open / tb_peer_close in ANOTHER process
 *   tb_peer_signal  stream-ordered: after everything queued before it on `stream` has completed, store `value` to *flag
 *                   with system scope (flag may be peer memory)
 *   tb_peer_wait    stream-ordered: hold the stream until *flag >= value (acquire, system scope). Gives up after timeout_ms
 *                   (then stores the observed value + 1 into *status, if given, and lets the stream continue, so that a
 *                   dead neighbour cannot hang the GPU); *status stays 0 on success. */
TB_API int tb_peer_alloc(size_t bytes, void** dev_ptr_out);
TB_API int tb_peer_free(void* dev_ptr);
TB_API int tb_peer_export(const void* dev_ptr, void* handle_out_64);
TB_API int tb_peer_open(const void* handle_64, void** dev_ptr_out);
TB_API int tb_peer_close(void* dev_ptr);
TB_API int tb_peer_signal(uint64_t* flag, uint64_t value, void* stream);
TB_API int tb_peer_wait(const uint64_t* flag, uint64_t value, int timeout_ms, uint64_t* status, void* stream);
/* synchronous read of one flag / status word (device or mapped peer memory) into host memory */
TB_API int tb_peer_read(const uint64_t* word, uint64_t* value_out_host);

/* ------------------------------------------------------------------------------------------------
 * LOSSES  (forward value + gradient w.r.t. the logits in one call)
 * Layout for all of them: x is [B,C,inner] contiguous (inner = product of spatial dims).
 * Workspaces: query the size with the *_workspace_bytes twin; contents need no initialisation.
 * loss_out / stats_out are device pointers; the value is complete when the stream reaches that point.
 * ---------------------------------------------------------------------------------------------- */

enum tb_loss_reduction { TB_LOSS_NONE = 0, TB_LOSS_MEAN = 1, TB_LOSS_SUM = 2, TB_LOSS_BATCHWISE = 3, TB_LOSS_NORMALIZED = 4 };
enum tb_target_kind { TB_T_DENSE_F32 = 0, TB_T_LABELS_I64 = 1 };

typedef struct tb_focal_params {
  float gamma;
  float alpha;             /* used when has_alpha */
  int has_alpha;
  float reduced_threshold; /* used when has_reduced_threshold */
  int has_reduced_threshold;
  float ignore_index;      /* compared against the float target; used when has_ignore */
  int has_ignore;
  int normalized;
  float eps;               /* clamp_min of the normalisation factor (1e-6) */
  int reduction;           /* enum tb_loss_reduction */
  int softmax;             /* 0: p = sigmoid(x); 1: p = softmax over C (activation="softmax", softmax_dim=1) */
} tb_focal_params;

TB_API size_t tb_loss_workspace_bytes(int64_t B, int C, int64_t inner);

/* Sigmoid (binary) focal loss. Replaces focal_loss_with_logits (losses/functional.py:19-107) and the
 * one-hot expansion of BinaryFocalLoss (losses/focal.py:77-105) when t_kind = TB_T_LABELS_I64.
 *   x f32 [B,C,inner]; t f32 [B,C,inner] or i64 [B,inner]; class_w NULL or f32 [C]
 *   loss_out: f32 scalar (MEAN/SUM), [B,C,inner] (NONE), or for BATCHWISE a buffer of (B+1)*C*inner floats: the
 *             first C*inner receive the result, the rest is scratch for the element losses
 *   grad: NULL or f32 [B,C,inner] = d(sum-reduced or mean-reduced scalar)/dx; for NONE/BATCHWISE it is the
 *         per-element derivative d loss_e / d x_e (diagonal), not available with normalized
 */
TB_API int tb_focal_fwd_bwd(const float* x, const void* t, int t_kind, int64_t B, int C, int64_t inner,
                     const tb_focal_params* p, const float* class_w, float* loss_out, float* grad, void* ws,
                     size_t ws_bytes, void* stream);

/* Softmax focal loss. Replaces softmax_focal_loss_with_logits (losses/functional.py:110-173).
 *   x f32 [B,C,inner]; t i64 [B,inner]; ignore_index (int64) always active (default -100)
 *   loss_out: scalar (MEAN/SUM), [B,inner] (NONE), or for BATCHWISE (B+1)*inner floats (result first, scratch after) */
TB_API int tb_softmax_focal_fwd_bwd(const float* x, const int64_t* t, int64_t B, int C, int64_t inner, float gamma,
                             float reduced_threshold, int has_reduced_threshold, int64_t ignore_index,
                             int normalized, float eps, int reduction, const float* class_w, float* loss_out,
                             float* grad, void* ws, size_t ws_bytes, void* stream);

/* Dice / Jaccard. mode: 0 binary, 1 multiclass (softmax over C, i64 labels), 2 multilabel.
 * Replaces DiceLoss.forward (losses/dice.py:59-131) / JaccardLoss.forward (losses/jaccard.py:48-103) and
 * soft_dice_score / soft_jaccard_score (losses/functional.py:188-247).
 * Phase 1 (stats): per class c: I = sum p*t, S = sum (p+t), Tsum = sum t over (B,inner) -> stats_out f64 [C,3]
 * Phase 2 (loss + grad): loss_out f32 scalar; grad f32 [B,C,inner] (NULL to skip).
 */
typedef struct tb_dice_params {
  int jaccard;       /* 0 dice, 1 jaccard */
  int mode;          /* 0 binary, 1 multiclass, 2 multilabel */
  int from_logits;
  int log_loss;
  float smooth;
  float eps;
  int has_ignore;    /* dice only (dice.py:85-111) */
  float ignore_index;
} tb_dice_params;

TB_API int tb_dice_stats(const float* x, const void* t, int t_kind, int64_t B, int C, int64_t inner,
                  const tb_dice_params* p, double* stats_out /*[C,3]*/, void* ws, size_t ws_bytes, void* stream);
TB_API int tb_dice_loss_grad(const float* x, const void* t, int t_kind, int64_t B, int C, int64_t inner,
                      const tb_dice_params* p, const double* stats /*[C,3]*/, const int32_t* classes /*NULL or [n_classes]*/,
                      int n_classes, float* loss_out, float* grad, void* stream);

/* Joint sigmoid focal + dice / jaccard on a dense f32 target [B,C,inner] (BASELINE config 3:
 * BinaryFocalLoss + DiceLoss summed by the caller, focal.py:13-105 + dice.py:18-131): loss = w_focal * focal +
 * w_dice * region, and its gradient, in two passes over (x, t) and one gradient write.
 * Restrictions (TB_ERR_UNSUPPORTED otherwise): focal sigmoid, not normalised, reduction MEAN or SUM; region mode
 * binary / multilabel with from_logits. loss_out f32 [3] = (total, focal term, region term), unweighted terms.
 * Validated on a B200 in round 2 (tests/test_joint_labels_gpu.py). */
TB_API int tb_focal_dice_fwd_bwd(const float* x, const float* t, int64_t B, int C, int64_t inner, const tb_focal_params* fp,
                          const float* class_w, const tb_dice_params* dp, const int32_t* classes, int n_classes,
                          float w_focal, float w_dice, float* loss_out, float* grad, void* ws, size_t ws_bytes,
                          void* stream);

/* Wing loss. Replaces wing_loss (losses/functional.py:250-277). reduction NONE/MEAN/SUM.
 * grad = d loss / d x (NULL to skip); target gets no gradient here (shim negates for it). */
TB_API int tb_wing_fwd_bwd(const float* x, const float* t, int64_t n, float width, float curvature, int reduction,
                    float* loss_out, float* grad, void* ws, size_t ws_bytes, void* stream);

/* data[i] *= *scale_dev for i < n, skipped entirely when *scale_dev == 1.0f: lets autograd apply the upstream
 * gradient (a device scalar) to a stored gradient buffer without a host synchronisation. */
TB_API int tb_scale_by_device_scalar(float* data, int64_t n, const float* scale_dev, void* stream);

/* the reference's remaining element-wise losses on the same fused forward+gradient skeleton.
 *   kind 0  SoftBCEWithLogitsLoss (soft_bce.py:9-50): label smoothing (has_smooth / smooth), ignore_index, `weight` and
 *           `pos_weight` (device f32, weight_n / pos_weight_n = 1 or C entries, indexed by channel), reduction none|mean|sum
 *   kind 1  balanced_binary_cross_entropy_with_logits (balanced_bce.py:10-49): the counts of t == 1 and t == 0 are taken
 *           by a first pass over t (device-side, no host sync), `gamma`, ignore_index, reduction none|mean|sum
 *   kind 2  QualityFocalLoss (quality_focal_loss.py:5-53): `beta`, reduction none|mean|sum|TB_LOSS_NORMALIZED (sum L / sum F,
 *           gradient through both sums: two passes)
 * x, t f32 [B,C,inner]; loss_out f32 [1] (or [B,C,inner] element losses for TB_LOSS_NONE); grad f32 [B,C,inner] or NULL
 * (= d loss / d x, already divided by numel for MEAN); ws: tb_loss_workspace_bytes(B, C, inner). */
typedef struct tb_elem_loss_params {
  int kind;
  int reduction;
  float smooth;
  int has_smooth;
  float ignore_index;
  int has_ignore;
  int weight_n;
  int pos_weight_n;
  float gamma;
  float beta;
} tb_elem_loss_params;
TB_API int tb_elem_loss_fwd_bwd(const float* x, const float* t, int64_t B, int C, int64_t inner, const tb_elem_loss_params* p,
                         const float* weight, const float* pos_weight, float* loss_out, float* grad, void* ws, size_t ws_bytes,
                         void* stream);

/* soft F1 losses (soft_f1.py:8-107) on [N,C] logits (element (n,c) at x[n*C+c]):
 *   mode 0  BinarySoftF1Loss: C == 1, t f32 [N]; rows whose target equals `ignore` are dropped; p = clamp(sigmoid(x), eps, 1-eps)
 *   mode 1  SoftF1Loss: t i64 [N] class labels, p = clamp(softmax(x, 1), eps, 1-eps), per-class soft F1, mean over classes
 *   mode 2  SoftF1Loss with ignore_index: the reference compares the ONE-HOT VALUES with ignore_index and flattens what is
 *           kept into one F1 (soft_f1.py:97-106); replicated literally
 * loss_out f32 [1]; grad f32 [N,C] or NULL. C <= 64. */
TB_API size_t tb_soft_f1_workspace_bytes(int64_t N, int C);
TB_API int tb_soft_f1_fwd_bwd(const float* x, const void* t, int64_t N, int C, int mode, int has_ignore, float ignore, float eps,
                       float* loss_out, float* grad, void* ws, size_t ws_bytes, void* stream);

/* Lovasz hinge (mode 0; BinaryLovaszLoss, lovasz.py:37-86,187-194) and Lovasz-softmax (mode 1; LovaszLoss with
 * classes="present", lovasz.py:92-160,197-204) as ONE launch sequence for all (image, class) segments: errors + sort keys,
 * segmented descending radix sort (4 x 8 bits, payload = ground-truth bit | sign | index), positives scan, Jaccard
 * increments, dot product, means over present classes / images, and d loss / d input scattered back through the sorted
 * permutation. No host loop, no host synchronisation, no library sort.
 *   x       hinge: logits f32 [B,inner] (C must be 1); softmax: class PROBABILITIES f32 [B,C,inner] (no softmax is applied,
 *           as in the reference; C == 1 means "sigmoid output": foreground = (label == 0), lovasz.py:129-132)
 *   labels  [B,inner], TB_T_DENSE_F32 or TB_T_LABELS_I64; hinge labels must be 0/1 (or the ignore value)
 *   per_image  one segment set per image and the mean over images (lovasz.py:45-49,101-105), else one over the batch
 *   has_ignore / ignore  pixels whose label equals `ignore` are dropped (lovasz.py:82-86,156-160): they sort last and
 *           get a zero gradient; an image / batch without valid pixels contributes 0
 *   loss_out f32 [1]; grad f32, same shape as x, or NULL (= d loss_out / d x)
 *   ws      tb_lovasz_fused_workspace_bytes(B, C, inner, per_image) bytes, 256-byte aligned (16 B per key + histograms)
 * Limits: <= 2^30 elements per segment, <= 65535 segments. Ties between equal errors are ordered by pixel index (the loss
 * is invariant to the order inside a tie group; torch.sort leaves it unspecified). */
TB_API size_t tb_lovasz_fused_workspace_bytes(int64_t B, int C, int64_t inner, int per_image);
TB_API int tb_lovasz_fwd_bwd(const float* x, const void* labels, int t_kind, int64_t B, int C, int64_t inner, int mode,
                      int per_image, int has_ignore, float ignore, float* loss_out, float* grad, void* ws, size_t ws_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bench helper: the deterministic, D4-non-equivariant stand-in "model" of SURVEY.md section 8(d):
 *   y[b,k,i,j] = sum_c Wm[k,c]*x[b,c,i,j] + 0.01*(i - 2j + 7k)/T     (not part of the reference)
 * ---------------------------------------------------------------------------------------------- */
TB_API int tb_standin_model(const float* x, int B, int Cin, int K, int Th, int Tw, const float* Wm_host /*[K,Cin]*/,
                     float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TB200_H_ */
