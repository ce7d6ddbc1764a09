/*
 * alscore.h -- C ABI of the B200-native pool-scoring path.
 *
 * Drop-in boundary for the acquisition step of
 * alfrunesiq/SemanticSegmentationActiveLearning.  The reference has no FFI of its
 * own (it is TensorFlow graph code inline in main()), so every entry point cites
 * the reference lines (relative to /root/reference) whose behaviour it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no framework types.
 *   - logits are dense NHWC, class innermost:  [N,H,W,C]  or  [T,N,H,W,C]
 *     (T Monte-Carlo samples outermost), float32 or bfloat16, 16-byte aligned.
 *     (reference: pseudo_logits, active_learning.py:231; models/enet/enet_modules.py:1376-1380)
 *   - "device" pointers must belong to the context's GPU; "host" pointers are
 *     staged to the GPU (never scored on the CPU -- there is no CPU fallback).
 *   - every function returns ALS_OK (0) or a negative als_status; the message is
 *     available from als_last_error().  Outputs are never partially filled on error.
 *   - calls are synchronous with respect to the host unless a function says otherwise.
 *   - a `stream` argument is a cudaStream_t: NULL is the legacy default stream (what frameworks call "the default
 *     stream"), ALS_STREAM_CTX is the context's stream.  Entries without a `stream` argument (the als_pool_*,
 *     als_mc_* and *_host families) run on the context's stream, see als_ctx_set_stream.
 *   - a context owns ONE set of per-image accumulators.  Calls on different streams are allowed: a call whose stream
 *     differs from the previous call's first waits (on the device) for the work queued on that stream, so the scratch is
 *     never shared by two launches in flight; the previous call's stream must still exist at that moment.  One context
 *     per host thread.
 */
#ifndef ALSCORE_H_
#define ALSCORE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALS_VERSION 110 /* 0.1.1 */

/* `stream` argument meaning "the context's stream" (NULL is the legacy default stream). */
#define ALS_STREAM_CTX ((void*)(intptr_t)-1)

#if defined(__GNUC__)
#define ALS_API __attribute__((visibility("default")))
#else
#define ALS_API
#endif

typedef struct als_ctx als_ctx;

typedef enum als_status {
  ALS_OK = 0,
  ALS_ERR_INVALID = -1,     /* bad shape / dtype / stride / alignment / NULL pointer   -> ValueError        */
  ALS_ERR_UNSUPPORTED = -2, /* unknown measure (active_learning.py:259-260)            -> NotImplementedError */
  ALS_ERR_CUDA = -3,        /* CUDA runtime error                                       -> RuntimeError      */
  ALS_ERR_NOMEM = -4,       /* allocation failed                                        -> MemoryError       */
  ALS_ERR_STATE = -5        /* call sequence error (e.g. pool not begun)                -> RuntimeError      */
} als_status;

/* alparams["measure"] (active_learning.py:240,252,256; conf/default_params.json:49).
 * ALS_VARIANCE is an extension (MC-dropout; needs T >= 2), absent from the reference. */
typedef enum als_measure {
  ALS_ENTROPY = 0,    /* 1 - H(p)/log(C)            active_learning.py:240-251 */
  ALS_MARGIN = 1,     /* p(1) - p(2)                active_learning.py:252-255 */
  ALS_CONFIDENCE = 2, /* max_c p                    active_learning.py:256-258 */
  ALS_VARIANCE = 3    /* 1 - sum_c Var_t[p_t,c]     (extension)               */
} als_measure;

typedef enum als_dtype { ALS_F32 = 0, ALS_BF16 = 1 } als_dtype;

/* ---- library / context ------------------------------------------------------------ */

ALS_API int als_version(void);

/* Number of CUDA devices visible, or a negative als_status. */
ALS_API int als_device_count(void);

/* One context per GPU and per host thread: owns a stream, scratch accumulators,
 * staging buffers and the pool score vector. */
ALS_API int als_ctx_create(int device, als_ctx** out);
ALS_API int als_ctx_destroy(als_ctx* ctx);

/* Make the context's stream the caller's cudaStream_t (e.g. the framework's current stream, so scoring is ordered
 * after the kernels that produced the logits).  The context does not take ownership.  NULL is the legacy default
 * stream.  No host synchronisation: work already queued on the previous stream is ordered before anything the
 * context queues on the new one (event wait on the device). */
ALS_API int als_ctx_set_stream(als_ctx* ctx, void* stream);

/* Last error message of this context (ctx == NULL: of the calling thread). */
ALS_API const char* als_last_error(const als_ctx* ctx);

/* "entropy" | "margin" | "confidence" | "variance" -> als_measure.
 * Unknown names return ALS_ERR_UNSUPPORTED with the reference's message
 * "Uncertainty function not implemented." (active_learning.py:259-260). */
ALS_API int als_measure_from_name(const char* name, int* measure);

/* Count of this library's kernel launches issued through ctx since creation. */
ALS_API int64_t als_launch_count(const als_ctx* ctx);

/* Observability: with timing on, every scoring launch sequence of the logits path (scoring kernel + finalize) is
 * bracketed by CUDA events on its stream; als_last_scoring_ms waits for the last one and returns its device time.
 * (The reference's only tracing hook is a commented-out tf.RunOptions(FULL_TRACE), train.py:293-294.) */
ALS_API int als_ctx_enable_timing(als_ctx* ctx, int on);
ALS_API int als_last_scoring_ms(als_ctx* ctx, float* ms);

/* ---- graph-level boundary (replaces active_learning.py:234-269) ------------------- */

/*
 * Score N images held in DEVICE memory.
 *   logits     device, [T,N,H,W,C] (T == 1: [N,H,W,C]), dtype f32/bf16, 16-byte aligned
 *   scores     device f64[N]     pseudo_mean_confidence (active_learning.py:261-263); required
 * Optional per-pixel outputs (device, may each be NULL) -- the training-path consumers:
 *   conf_map   f32[N,H,W]        pseudo_confidence      (active_learning.py:251/255/258)
 *   label      u8[N,H,W]         pseudo_label           (active_learning.py:234-236; sample 0 if T > 1)
 *   mask       u8[N,H,W]         pseudo_mask            (active_learning.py:265-269), conf < threshold ? 0 : 1
 * Asynchronous on `stream`; the caller synchronises.
 */
ALS_API int als_score(als_ctx* ctx, const void* logits, int dtype,
              int64_t T, int64_t N, int64_t H, int64_t W, int64_t C, int measure,
              double* scores, float* conf_map, uint8_t* label, uint8_t* mask, float threshold,
              void* stream);

/*
 * Same for logits in HOST memory (pinned or pageable): staged host->device in chunks on a
 * copy stream, overlapped with scoring; `scores` is host f64[N]; per-pixel outputs (host,
 * optional) are copied back.  Synchronous.
 */
ALS_API int als_score_host(als_ctx* ctx, const void* logits, int dtype,
                   int64_t T, int64_t N, int64_t H, int64_t W, int64_t C, int measure,
                   double* scores, float* conf_map, uint8_t* label, uint8_t* mask, float threshold);

/*
 * DLPack entry: `managed` is a DLManagedTensor* (dlpack.h, v0.x ABI) taken from a
 * "dltensor" capsule.  kDLCUDA / kDLCUDAManaged tensors are scored in place (zero copy),
 * kDLCPU / kDLCUDAHost are staged.  Checks dtype in {f32, bf16}, ndim in {4,5}, dense
 * C-order strides, 16-byte alignment; anything else is ALS_ERR_INVALID (never a silent copy).
 * `scores` is HOST f64[N].  The tensor is only borrowed; the caller keeps ownership and
 * runs the deleter.  Synchronous.
 * Stream exchange (DLPack protocol): device tensors are scored on `stream`, the stream the consumer handed to the
 * producer's __dlpack__(stream=...) -- the producer has made the data ready there, so no device-wide synchronisation
 * is needed.  Host tensors ignore it.
 */
ALS_API int als_score_dlpack(als_ctx* ctx, void* managed, int measure, double* scores, void* stream);

/* ---- fused classifier head (replaces Final.call + active_learning.py:234-269) --------- */

/*
 * The logits scored above are produced by ENet's `Final` layer, a 3x3 stride-2 transposed
 * convolution 16 -> C without bias (models/enet/enet_modules.py:1294-1381, called from
 * models/enet/enet.py:367).  These entries take its INPUT feature map and its kernel instead of
 * the logits: the contraction runs on the tensor cores (tcgen05, split-TF32 = fp32-level accuracy)
 * inside the scoring kernel, so the [N,2h,2w,C] logits tensor is never written to or read from HBM.
 *
 * als_head_prepare: upload the layer's kernel.  `kernel` is HOST float32 [3][3][C][16]
 * (tf.nn.conv2d_transpose filter layout [kh, kw, out_channels, in_channels], enet_modules.py:1341).
 * ALS_ERR_UNSUPPORTED if no fused kernel is built for this class count (see als_head_supported).
 */
ALS_API int als_head_prepare(als_ctx* ctx, const float* kernel, int64_t C);

/*
 * Host-only helpers (no GPU needed; used by the CPU tests of the operand packing):
 * als_head_geometry fills geom[14] = {C, CB, n[0..3], col0[0..3], row0[0..3]} followed by the total row count
 * in *rows: the accumulator has 4 blocks of CB = round_up(C,4) columns, one per pixel of the 2x2 output quad,
 * in the order (dy,dx) = (0,1) (0,0) (1,0) (1,1); operand o multiplies source pixel (i - (o&1), j - (o>>1)),
 * is n[o] columns wide, starts at accumulator column col0[o] and at row row0[o] of the packed image.
 * als_head_pack_weights writes out[2][4][rows][4] floats (tf32 hi part, lo part; 16-byte channel chunks).
 */
ALS_API int als_head_geometry(int64_t C, int32_t* geom14, int32_t* rows);
ALS_API int als_head_pack_weights(const float* kernel, int64_t C, float* out, int64_t out_floats);

/* 1 if a fused-head kernel exists for (C, measure, T samples), else 0: 2 <= C <= 32 for T == 1,
 * 2 <= C <= 24 for T >= 2 (the per-class Welford state of two pixels has to fit a thread's registers). */
ALS_API int als_head_supported(int64_t C, int measure, int64_t T);

/*
 * Score N images from their `Final`-layer input.
 *   features   device f32 [T,N,h,w,16] (T == 1: [N,h,w,16]), dense NHWC, 16-byte aligned; T > 1 = the layer's
 *              input under T Monte-Carlo-dropout forward passes (sample outermost, like the logits of als_score)
 *   scores     device f64[N]; conf_map / label / mask: optional device [N,2h,2w] as in als_score
 * Same results as als_score on conv2d_transpose(features[t], kernel), t = 0..T-1 (to fp32 rounding).
 * Asynchronous on `stream`.
 */
ALS_API int als_score_features(als_ctx* ctx, const void* features, int64_t T, int64_t N, int64_t h, int64_t w,
                       int measure, double* scores, float* conf_map, uint8_t* label, uint8_t* mask, float threshold,
                       void* stream);

/* :697-700 with the fused head: like als_pool_score_batch, from the `Final`-layer input [T,B,h,w,16] (device or host). */
ALS_API int als_pool_score_features_batch(als_ctx* ctx, const void* features, int features_on_host, int64_t T, int64_t B,
                                  int64_t h, int64_t w, int measure, const int64_t* example_index);

/* ---- loop-level boundary (replaces rank_confidence, active_learning.py:682-715) ---- */

/* :684-685  confidence = np.zeros(num_examples, float32)  (device-resident). */
ALS_API int als_pool_begin(als_ctx* ctx, int64_t num_examples);

/*
 * :697-700  one sess.run + scatter:  confidence[example_index[b]] = float32(score64[b]).
 *   logits         [T,B,H,W,C] in device (logits_on_host == 0) or host memory
 *   example_index  HOST int64[B], each in [0, num_examples)
 * Asynchronous w.r.t. the host for device logits; host logits return once staged.
 */
ALS_API int als_pool_score_batch(als_ctx* ctx, const void* logits, int logits_on_host, int dtype,
                         int64_t T, int64_t B, int64_t H, int64_t W, int64_t C, int measure,
                         const int64_t* example_index);

/* Copy the whole f32[num_examples] confidence vector to the host (synchronises). */
ALS_API int als_pool_scores(als_ctx* ctx, float* out, int64_t num_examples);

/*
 * :705-715  filter to `unlabelled`, pick the k = min(M, selection_size) LOWEST confidences.
 *   unlabelled              HOST int64[M], unique ids in [0, num_examples) (duplicates are ALS_ERR_INVALID)
 *   out_ids                 HOST int64[min(k, M)], ascending in (confidence, id)
 *   out_unlabelled_conf     HOST f32[M]  == confidence[unlabelled]           (may be NULL)
 *   out_count               number of ids written
 * Total order (the reference leaves these to np.argpartition's whim): -0.0 == +0.0,
 * NaN after +inf, ties broken by the lower example id; k >= M returns all M
 * (the reference raises ValueError there).  Synchronous.
 */
ALS_API int als_pool_select(als_ctx* ctx, const int64_t* unlabelled, int64_t M, int64_t selection_size,
                    int64_t* out_ids, float* out_unlabelled_conf, int64_t* out_count);

/*
 * :682-715 in ONE call, for a pool whose logits arrive as one tensor (device or host): als_pool_begin(num_examples) +
 * als_pool_score_batch(logits, example_index) + als_pool_select(unlabelled, selection_size), with all host-side work done
 * before the first launch so that a small pool's single scoring launch is not stretched by gaps between calls.
 *   example_index   HOST int64[N] or NULL (= 0 .. N-1)
 * Other arguments and results as in the three calls it replaces.  Synchronous.
 */
ALS_API int als_rank_pool(als_ctx* ctx, const void* logits, int logits_on_host, int dtype,
                          int64_t T, int64_t N, int64_t H, int64_t W, int64_t C, int measure,
                          const int64_t* example_index, int64_t num_examples,
                          const int64_t* unlabelled, int64_t M, int64_t selection_size,
                          int64_t* out_ids, float* out_unlabelled_conf, int64_t* out_count);

/*
 * Device-level selection primitive (block-radix select), also used to merge the
 * per-GPU candidates after the all-gather.
 *   keys  device f32[M];  ids  device int64[M] (unique: duplicated (key, id) pairs leave output slots unwritten);  k >= 0
 *   out_keys device f32[min(k,M)], out_ids device int64[min(k,M)], ascending in (key, id).
 * Asynchronous on `stream`.
 */
ALS_API int als_select_smallest(als_ctx* ctx, const float* keys, const int64_t* ids, int64_t M, int64_t k,
                        float* out_keys, int64_t* out_ids, void* stream);

/* ---- multi-GPU: pool sharded by image, one all-gather (no counterpart in the reference) ------------
 *
 * The reference scores the whole pool on GPU:0 (active_learning.py:221, :689-700).  Images are independent, so the pool
 * shards by image: every rank keeps a FULL-SIZE confidence vector (als_pool_begin(num_examples)), scores only the
 * examples it owns -- the contiguous id range [shard_lo, shard_hi); the ranges of all ranks must tile
 * [0, num_examples) -- and als_pool_select_global finishes :705-715 for the whole pool:
 *   1. on the device, one launch: the rank's k lowest (confidence, id) candidates among unlabelled ids it owns, packed
 *      with its score slice into one record  {lo, n, f32 key[k], i64 id[k], f32 score[max_shard]};
 *   2. ONE ncclAllGather of the records over NVLink (a few hundred KB at most: latency bound);
 *   3. on the device, one launch: the other ranks' score slices complete the local confidence vector, the world*k
 *      candidates are merged with the same block-radix select, unlabelled_confidence is gathered;
 *   4. the results {ids, unlabelled_confidence} are written by that launch straight into pinned host memory.
 * Every rank returns the same result.  Examples nobody visited keep 0.0 and are selected first, like :685/:701-702.
 */

/* 128-byte NCCL unique id for als_comm_init_rank (rank 0 creates it and hands it to the others out of band). */
ALS_API int als_comm_unique_id(void* out128);

/* One process per GPU: join a communicator of `world` ranks.  Collective (every rank calls it). */
ALS_API int als_comm_init_rank(als_ctx* ctx, int rank, int world, const void* unique_id128);

/* One process, n GPUs (the reference's own process model, active_learning.py:221,277): ctxs[i] becomes rank i. */
ALS_API int als_comm_init_all(als_ctx** ctxs, int n);

ALS_API int als_comm_destroy(als_ctx* ctx);

/*
 * :705-715 over the sharded pool.  Collective; same arguments on every rank except shard_lo / shard_hi.
 *   unlabelled       HOST int64[M], unique GLOBAL ids in [0, num_examples)        (same on every rank)
 *   shard_lo/hi      the id range this rank owns and has scored
 *   max_shard        upper bound on any rank's shard_hi - shard_lo, identical on all ranks; 0 = ceil(num_examples / world)
 *   out_*            as als_pool_select
 * Without a communicator (world 1) it is als_pool_select.  Synchronous.
 */
ALS_API int als_pool_select_global(als_ctx* ctx, const int64_t* unlabelled, int64_t M, int64_t selection_size,
                                   int64_t shard_lo, int64_t shard_hi, int64_t max_shard,
                                   int64_t* out_ids, float* out_unlabelled_conf, int64_t* out_count);

/* Single-process form: ctxs[i] owns [shard_lo[i], shard_hi[i]); results are read back from ctxs[0]. */
ALS_API int als_pool_select_global_all(als_ctx** ctxs, int n, const int64_t* unlabelled, int64_t M,
                                       int64_t selection_size, const int64_t* shard_lo, const int64_t* shard_hi,
                                       int64_t max_shard, int64_t* out_ids, float* out_unlabelled_conf,
                                       int64_t* out_count);

/* ---- streamed Monte-Carlo accumulation (extension) --------------------
 *
 * als_score takes all T dropout samples at once as [T,N,H,W,C] (948 GB for 2975 images at T = 8).  A producer that
 * makes one stochastic forward pass at a time -- the reference only builds its dropout under training=True
 * (models/util/extra_ops.py:137-151, models/enet/enet_modules.py:591-594) -- hands them over one by one instead:
 * the per-pixel Welford state (running mean per class + summed M2, float32) stays resident in HBM between samples
 * and [T,...] never materialises.  Results are bit-identical to als_score on the stacked samples.
 * Bytes per pixel and sample: C*sizeof(elem) read + (C+1)*4 read + (C+1)*4 written (sample 0 skips the state read).
 * All three run on the context's stream and are asynchronous.
 */

/* Open an accumulation over a batch of N images [N,H,W,C].  label (optional, device u8[N,H,W]) receives
 * pseudo_label = argmax of sample 0 (active_learning.py:234-236). */
ALS_API int als_mc_begin(als_ctx* ctx, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, uint8_t* label);

/* Fold one sample in: logits [N,H,W,C] of the dtype given to als_mc_begin, device (logits_on_host == 0) or host
 * memory (staged; returns once staged). */
ALS_API int als_mc_add_sample(als_ctx* ctx, const void* logits, int logits_on_host);

/* Close it: confidence of the predictive mean (entropy / margin / confidence) or 1 - summed variance (variance; needs
 * >= 2 samples), per-image f64 mean.
 *   scores         device f64[N], optional
 *   example_index  HOST int64[N], optional: scatter float32(score) into the pool vector like als_pool_score_batch
 *   conf_map / mask  optional device [N,H,W] as in als_score */
ALS_API int als_mc_finish(als_ctx* ctx, int measure, double* scores, const int64_t* example_index, float* conf_map,
                          uint8_t* mask, float threshold);

/* ---- synthetic pool (bench / tests) ------------------------------------------------ */

/*
 * Counter-based logits generator, bit-identical to oracle/synth.py.  Writes images
 * n0 .. n0+n_imgs-1 of the synthetic pool as device [T, n_imgs, H, W, C].
 * mc != 0 adds the per-sample noise term (use mc = (T > 1)).  Asynchronous on `stream`.
 */
ALS_API int als_synth_logits(als_ctx* ctx, void* out, int dtype, int64_t T, int64_t n0, int64_t n_imgs,
                     int64_t H, int64_t W, int64_t C, uint64_t seed, int mc, void* stream);

/* Write `bytes` of device scratch (> L2) so the next timed launch starts cold. */
ALS_API int als_flush_l2(als_ctx* ctx, void* stream);

/* Describe the kernel als_score would launch for this shape (for bench / profiles):
 * fills name (<= 127 chars), grid, block, dynamic smem bytes, pipeline stages, pixels per tile. */
ALS_API int als_describe_launch(als_ctx* ctx, int dtype, int64_t T, int64_t N, int64_t H, int64_t W, int64_t C,
                        int measure, char* name, int* grid, int* block, int* smem_bytes, int* stages,
                        int* tile_pixels);

/* Same for the fused-head kernel als_score_features would launch (needs als_head_prepare). */
ALS_API int als_describe_head_launch(als_ctx* ctx, int64_t T, int measure, char* name, int* grid, int* block,
                                     int* smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ALSCORE_H_ */
