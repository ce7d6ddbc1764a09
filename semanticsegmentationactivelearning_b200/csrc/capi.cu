// C ABI of the pool-scoring path (include/alscore.h): contexts, validation, host staging,
// DLPack borrowing, the rank_confidence-shaped pool API.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ctx.h"
#include "head.cuh"
#include "mc.cuh"
#include "score.cuh"
#include "select.cuh"
#include "synth.cuh"

// ---- minimal DLPack v0.x ABI mirror (dlpack.h: DLDevice, DLDataType, DLTensor, DLManagedTensor)
extern "C" {
typedef struct { int32_t device_type; int32_t device_id; } AlsDLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } AlsDLDataType;
typedef struct {
  void* data; AlsDLDevice device; int32_t ndim; AlsDLDataType dtype;
  int64_t* shape; int64_t* strides; uint64_t byte_offset;
} AlsDLTensor;
typedef struct AlsDLManagedTensor {
  AlsDLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct AlsDLManagedTensor*);
} AlsDLManagedTensor;
}
static_assert(sizeof(AlsDLTensor) == 48, "DLTensor ABI");
enum { kAlsDLCPU = 1, kAlsDLCUDA = 2, kAlsDLCUDAHost = 3, kAlsDLCUDAManaged = 13 };
enum { kAlsDLFloat = 2, kAlsDLBfloat = 4 };

static thread_local std::string g_tls_error;

namespace als {

int fail(als_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_tls_error = buf;
  if (ctx) ctx->error = buf;
  return code;
}

int grow_bytes(als_ctx* ctx, void** ptr, size_t* cap, size_t need) {
  if (need <= *cap) return ALS_OK;
  size_t n = *cap ? *cap : 4096;
  while (n < need) n *= 2;
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*ptr) ALS_CUDA(ctx, cudaFree(*ptr));
  *ptr = nullptr;
  *cap = 0;
  ALS_CUDA(ctx, cudaMalloc(ptr, n));
  *cap = n;
  return ALS_OK;
}

int grow_pinned(als_ctx* ctx, void** ptr, size_t* cap, size_t need) {
  if (need <= *cap) return ALS_OK;
  size_t n = *cap ? *cap : 4096;
  while (n < need) n *= 2;
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*ptr) ALS_CUDA(ctx, cudaFreeHost(*ptr));
  *ptr = nullptr;
  *cap = 0;
  ALS_CUDA(ctx, cudaMallocHost(ptr, n));
  *cap = n;
  return ALS_OK;
}

cudaStream_t resolve_stream(als_ctx* ctx, void* stream) {
  return stream == ALS_STREAM_CTX ? ctx->stream : static_cast<cudaStream_t>(stream);
}

// The accumulators / flags / tile counter are one set per context: a launch sequence on another stream than the
// previous one first waits for everything queued on that stream.  The event is recorded LAZILY, at the switch (it then
// covers the last launch sequence and whatever else the old stream holds): recording one after every call would put a
// stream operation between consecutive scoring launches and cost them their programmatic-dependent-launch overlap
// (measured: 52.0 -> 50 us per batch-of-8 call).  The stream a call ran on must therefore still exist when the next call
// arrives on a different one; if recording on it fails the context falls back to a device-wide synchronisation.
int scratch_begin(als_ctx* ctx, cudaStream_t st) {
  if (ctx->scratch_used && st != ctx->scratch_stream) {
    cudaError_t e = cudaEventRecord(ctx->ev_scratch, ctx->scratch_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ctx->ev_scratch, 0);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      ALS_CUDA(ctx, cudaDeviceSynchronize());
    }
  }
  return ALS_OK;
}
int scratch_end(als_ctx* ctx, cudaStream_t st) {
  ctx->scratch_stream = st;
  ctx->scratch_used = true;
  return ALS_OK;
}

int check_device_ptr(als_ctx* ctx, const void* p, const char* what) {
  if (!p) return ALS_OK;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(ctx, ALS_ERR_INVALID, "%s: not a CUDA pointer (%s)", what, cudaGetErrorString(e));
  }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
    return fail(ctx, ALS_ERR_INVALID, "%s must be device memory (use the *_host entry for host memory)", what);
  if (a.type == cudaMemoryTypeDevice && a.device != ctx->device)
    return fail(ctx, ALS_ERR_INVALID, "%s lives on GPU %d but the context is bound to GPU %d", what, a.device, ctx->device);
  return ALS_OK;
}

}  // namespace als

using als::check_device_ptr;
using als::DeviceGuard;
using als::fail;
using als::grow;
using als::grow_bytes;
using als::grow_pinned;
using als::resolve_stream;
using als::scratch_begin;
using als::scratch_end;

namespace {

// Fixed-point scale of the per-image sums.  f32 logits: as fine as 63 bits allow (2^44 at most: the quantisation of a
// confidence is then below 2^-45, far under an f32 half-ulp -- the sum is the exact f64 sum of the f32 map).  bf16
// logits: 2^22, the grid the FMA-pipe conversion of the hot kernel produces (pixel_math.cuh: ImageAcc::add_q22); with
// 8-bit inputs and a 1e-2 tolerance nothing is lost, and every bf16 path (tiled, generic, streamed) uses the same grid.
int ceil_log2_ll(long long v);
int fx_shift_for(int dtype, long long P) {
  if (dtype == ALS_BF16) return 22;
  int shift = 62 - ceil_log2_ll(P);
  return shift > 44 ? 44 : shift;
}

int ceil_log2_ll(long long v) {
  int b = 0;
  while ((1ll << b) < v) ++b;
  return b;
}

struct Shape {
  int64_t T, N, H, W, C;
  int64_t P() const { return H * W; }
  int64_t elems() const { return T * N * H * W * C; }
};

int check_common(als_ctx* ctx, const void* logits, int dtype, const Shape& s, int measure, bool want_label) {
  // (no alignment requirement: host logits are staged into an aligned buffer, misaligned device logits take the
  //  generic kernel, see score_device)
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (dtype != ALS_F32 && dtype != ALS_BF16) return fail(ctx, ALS_ERR_INVALID, "logits dtype must be float32 or bfloat16");
  if (measure < ALS_ENTROPY || measure > ALS_VARIANCE)
    return fail(ctx, ALS_ERR_UNSUPPORTED, "Uncertainty function not implemented.");
  if (s.T < 1 || s.N < 0 || s.H < 1 || s.W < 1) return fail(ctx, ALS_ERR_INVALID, "bad logits shape [T=%lld,N=%lld,H=%lld,W=%lld,C=%lld]",
                                                            (long long)s.T, (long long)s.N, (long long)s.H, (long long)s.W, (long long)s.C);
  if (s.C < 2) return fail(ctx, ALS_ERR_INVALID, "need at least 2 classes, got C=%lld", (long long)s.C);
  if (s.C > 65536) return fail(ctx, ALS_ERR_INVALID, "C=%lld is too large", (long long)s.C);
  if (measure == ALS_VARIANCE && s.T < 2) return fail(ctx, ALS_ERR_INVALID, "measure 'variance' needs T >= 2 Monte-Carlo samples");
  if (want_label && s.C > 256) return fail(ctx, ALS_ERR_INVALID, "uint8 pseudo_label needs C <= 256");
  if (s.P() > (1ll << 40)) return fail(ctx, ALS_ERR_INVALID, "image too large");
  if (s.N > 0x7fffffff) return fail(ctx, ALS_ERR_INVALID, "too many images in one call");
  if (s.N > 0 && !logits) return fail(ctx, ALS_ERR_INVALID, "logits pointer is NULL");
  return ALS_OK;
}

// Core: device logits -> fixed-point sums -> finalize.  scores64 / pool scatter optional.
int score_device(als_ctx* ctx, const void* logits, int dtype, const Shape& s, int measure, double* scores64,
                 float* pool32, const long long* example_index_dev, int64_t num_examples, float* conf_map,
                 uint8_t* label, uint8_t* mask, float threshold, cudaStream_t stream, int64_t index_base = 0) {
  if (s.N == 0) return ALS_OK;
  const long long P = s.P();
  const int es = dtype == ALS_F32 ? 4 : 2;
  const long long sample_stride = s.N * P * s.C;
  // the bulk copies need 16-byte aligned sample planes; anything else runs the generic kernel (same results)
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) % 16 == 0) && ((s.T == 1) || ((sample_stride * es) % 16 == 0));
  als::LaunchPlan plan = als::plan_score(dtype, static_cast<int>(s.C), measure, static_cast<int>(s.T), s.N * P, aligned,
                                         ctx->num_sms, ctx->max_smem);
  ALS_TRY(scratch_begin(ctx, stream));
  const int shift = fx_shift_for(dtype, P);
  als::ScoreParams p{};
  p.logits = logits;
  p.sample_stride = sample_stride;
  p.total_pixels = s.N * P;
  p.P = P;
  p.T = static_cast<int>(s.T);
  p.C = static_cast<int>(s.C);
  p.measure = measure;
  p.inv_log2_c = static_cast<float>(1.0 / log2(static_cast<double>(s.C)));
  p.threshold = threshold;
  p.inv_T = 1.0f / static_cast<float>(s.T);
  p.fx_scale = ldexpf(1.0f, shift);
  p.acc = ctx->acc;
  p.acc_stride = ctx->acc_cap;
  p.flags = ctx->flags;
  p.tile_counter = ctx->tile_counter;
  p.conf_map = conf_map;
  p.label = label;
  p.mask = mask;
  const double inv_scale_p = ldexp(1.0, -shift) / static_cast<double>(P);
  if (ctx->timing) ALS_CUDA(ctx, cudaEventRecord(ctx->ev_t0, stream));
  struct TimingEnd {  // closes the timed interval on every return path below
    als_ctx* c; cudaStream_t st;
    ~TimingEnd() { if (c->timing) c->timing_valid = cudaEventRecord(c->ev_t1, st) == cudaSuccess; }
  } timing_end{ctx, stream};
  if (plan.tiled && s.N <= als::kFusedFinalizeMaxImages) {
    // small launches: the scoring kernel's last CTA finalizes (one launch per batch instead of two)
    p.fin_n = static_cast<int>(s.N);
    p.done_counter = reinterpret_cast<unsigned int*>(ctx->tile_counter + 1);
    p.fin_inv_scale_p = inv_scale_p;
    p.fin_scores64 = scores64;
    p.fin_pool32 = pool32;
    p.fin_example_index = example_index_dev;
    p.fin_index_base = index_base;
    p.fin_num_examples = num_examples;
    ALS_CUDA(ctx, als::launch_score(plan, dtype, p, stream));
    ctx->launches += 1;
    return scratch_end(ctx, stream);
  }
  ALS_CUDA(ctx, als::launch_score(plan, dtype, p, stream));
  ALS_CUDA(ctx, als::launch_finalize(ctx->acc, ctx->acc_cap, ctx->flags, ctx->tile_counter, static_cast<int>(s.N), inv_scale_p,
                                     scores64, pool32, example_index_dev, index_base, num_examples, stream));
  ctx->launches += 2;
  return scratch_end(ctx, stream);
}

// example_index -> device, unless it is the run first, first+1, ...: then the scatter needs no index vector at all
// (the usual case when a pool is walked in order; saves the small host->device copy in front of every batch)
int stage_index(als_ctx* ctx, const int64_t* example_index, int64_t B, const long long** dev, int64_t* base) {
  bool run = true;
  for (int64_t i = 1; i < B && run; ++i) run = example_index[i] == example_index[0] + i;
  if (run) {
    *dev = nullptr;
    *base = example_index[0];
    return ALS_OK;
  }
  ALS_TRY(grow(ctx, &ctx->index_dev, &ctx->index_cap, B, false));
  // (pageable source: staged by the driver before the call returns)
  ALS_CUDA(ctx, cudaMemcpyAsync(ctx->index_dev, example_index, static_cast<size_t>(B) * sizeof(int64_t), cudaMemcpyHostToDevice,
                                ctx->stream));
  *dev = ctx->index_dev;
  *base = 0;
  return ALS_OK;
}

int ensure_acc(als_ctx* ctx, int64_t n) {
  if (n <= ctx->acc_cap) return ALS_OK;
  int64_t cap = ctx->acc_cap > 0 ? ctx->acc_cap : 1024;
  while (cap < n) cap *= 2;
  // the old set may still be in use on whichever stream ran the last launch sequence
  if (ctx->scratch_used && cudaStreamSynchronize(ctx->scratch_stream) != cudaSuccess) {
    (void)cudaGetLastError();
    ALS_CUDA(ctx, cudaDeviceSynchronize());
  }
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->acc) ALS_CUDA(ctx, cudaFree(ctx->acc));
  if (ctx->flags) ALS_CUDA(ctx, cudaFree(ctx->flags));
  ctx->acc = nullptr;
  ctx->flags = nullptr;
  ctx->acc_cap = 0;
  ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->acc), static_cast<size_t>(cap) * als::kAccReplicas * sizeof(long long)));
  ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->flags), static_cast<size_t>(cap) * sizeof(unsigned int)));
  // zero them before any stream can launch on them (cudaMemset on device memory is asynchronous to the host and
  // a non-blocking stream does not order behind the NULL stream: wait for it here, this is a rare grow)
  ALS_CUDA(ctx, cudaMemset(ctx->acc, 0, static_cast<size_t>(cap) * als::kAccReplicas * sizeof(long long)));
  ALS_CUDA(ctx, cudaMemset(ctx->flags, 0, static_cast<size_t>(cap) * sizeof(unsigned int)));
  ALS_CUDA(ctx, cudaDeviceSynchronize());
  ctx->acc_cap = cap;
  return ALS_OK;
}

constexpr size_t kStageDefault = 256ull << 20;

int ensure_stage(als_ctx* ctx, size_t need) {
  if (need <= ctx->stage_cap) return ALS_OK;
  size_t cap = ctx->stage_cap ? ctx->stage_cap : kStageDefault;
  while (cap < need) cap *= 2;
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  for (int b = 0; b < 2; ++b) {
    if (ctx->stage[b]) ALS_CUDA(ctx, cudaFree(ctx->stage[b]));
    ctx->stage[b] = nullptr;
  }
  ctx->stage_cap = 0;
  for (int b = 0; b < 2; ++b) ALS_CUDA(ctx, cudaMalloc(&ctx->stage[b], cap));
  ctx->stage_cap = cap;
  return ALS_OK;
}

// Stage images [n0, n0+nb) of host logits [T,N,P,C] into a staging buffer as [T,nb,P,C]; returns the buffer index.
int stage_chunk(als_ctx* ctx, const unsigned char* host, const Shape& s, int es, int64_t n0, int64_t nb, int* buf_out) {
  const int b = ctx->stage_next;
  ctx->stage_next ^= 1;
  const size_t img_bytes = static_cast<size_t>(s.P()) * s.C * es;
  // the previous user of this buffer must have been scored
  ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_scored[b], 0));
  // (als::stage_copy: pinned sources are one cudaMemcpyAsync, large pageable ones go through the parallel bounce pipeline)
  if (nb == s.N) {
    // the whole batch: [T, nb, P, C] is one dense range on both sides -> one copy instead of T
    ALS_TRY(als::stage_copy(ctx, ctx->stage[b], host, static_cast<size_t>(s.T) * nb * img_bytes));
  } else {
    for (int64_t t = 0; t < s.T; ++t) {
      const unsigned char* src = host + (static_cast<size_t>(t) * s.N + n0) * img_bytes;
      unsigned char* dst = static_cast<unsigned char*>(ctx->stage[b]) + static_cast<size_t>(t) * nb * img_bytes;
      ALS_TRY(als::stage_copy(ctx, dst, src, static_cast<size_t>(nb) * img_bytes));
    }
  }
  ALS_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
  *buf_out = b;
  return ALS_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int als_version(void) { return ALS_VERSION; }

int als_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(nullptr, ALS_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  return n;
}

const char* als_last_error(const als_ctx* ctx) { return ctx ? ctx->error.c_str() : g_tls_error.c_str(); }

int als_measure_from_name(const char* name, int* measure) {
  if (!name || !measure) return fail(nullptr, ALS_ERR_INVALID, "NULL argument");
  static const char* names[] = {"entropy", "margin", "confidence", "variance"};
  for (int i = 0; i < 4; ++i)
    if (strcmp(name, names[i]) == 0) {
      *measure = i;
      return ALS_OK;
    }
  // active_learning.py:259-260
  return fail(nullptr, ALS_ERR_UNSUPPORTED, "Uncertainty function not implemented.");
}

int als_ctx_create(int device, als_ctx** out) {
  if (!out) return fail(nullptr, ALS_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return fail(nullptr, ALS_ERR_CUDA, "no CUDA device available (%s); alscore has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(nullptr, ALS_ERR_INVALID, "device %d out of range [0, %d)", device, n);
  als_ctx* ctx = new als_ctx();
  ctx->device = device;
  DeviceGuard g(device);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete ctx;
    return fail(nullptr, ALS_ERR_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
  }
  if (prop.major != 10) {
    delete ctx;
    return fail(nullptr, ALS_ERR_CUDA, "alscore kernels are built for sm_100a (B200); device %d is sm_%d%d", device,
                prop.major, prop.minor);
  }
  ctx->num_sms = prop.multiProcessorCount;
  ctx->max_smem = static_cast<int>(prop.sharedMemPerBlockOptin);
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int b = 0; b < 2 && ok; ++b) {
    ok = cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_scored[b], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_scratch, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&ctx->ev_unl, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&ctx->tile_counter), 128) == cudaSuccess &&
       cudaMemset(ctx->tile_counter, 0, 128) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  if (!ok) {
    (void)cudaGetLastError();
    als_ctx_destroy(ctx);
    return fail(nullptr, ALS_ERR_CUDA, "failed to create streams/events");
  }
  *out = ctx;
  return ALS_OK;
}

int als_ctx_destroy(als_ctx* ctx) {
  if (!ctx) return ALS_OK;
  DeviceGuard g(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->scratch_used && cudaStreamSynchronize(ctx->scratch_stream) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaDeviceSynchronize();
  }
  als::stage_destroy(ctx);
  als_comm_destroy(ctx);
  void* ptrs[] = {ctx->acc, ctx->flags, ctx->tile_counter, ctx->scores_dev, ctx->index_dev, ctx->pool32, ctx->sel_ids,
                  ctx->sel_out, ctx->sel_tmp_keys, ctx->sel_tmp_ids, ctx->stage[0], ctx->stage[1], ctx->maps_dev,
                  ctx->flush_buf, ctx->head_weights, ctx->mc_state, ctx->xchg_send, ctx->xchg_recv};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (ctx->sel_ids_host) cudaFreeHost(ctx->sel_ids_host);
  if (ctx->sel_out_host) cudaFreeHost(ctx->sel_out_host);
  if (ctx->ev_scratch) cudaEventDestroy(ctx->ev_scratch);
  if (ctx->ev_unl) cudaEventDestroy(ctx->ev_unl);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
  for (int b = 0; b < 2; ++b) {
    if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
    if (ctx->ev_scored[b]) cudaEventDestroy(ctx->ev_scored[b]);
  }
  if (ctx->stream && ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  (void)cudaGetLastError();
  delete ctx;
  return ALS_OK;
}

int64_t als_launch_count(const als_ctx* ctx) { return ctx ? ctx->launches : 0; }

int als_ctx_enable_timing(als_ctx* ctx, int on) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  if (on && !ctx->ev_t0) {
    ALS_CUDA(ctx, cudaEventCreate(&ctx->ev_t0));
    ALS_CUDA(ctx, cudaEventCreate(&ctx->ev_t1));
  }
  ctx->timing = on != 0;
  ctx->timing_valid = false;
  return ALS_OK;
}

int als_last_scoring_ms(als_ctx* ctx, float* ms) {
  if (!ctx || !ms) return fail(ctx, ALS_ERR_INVALID, "NULL argument");
  if (!ctx->timing || !ctx->timing_valid) return fail(ctx, ALS_ERR_STATE, "no timed scoring launch (als_ctx_enable_timing)");
  DeviceGuard g(ctx->device);
  ALS_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
  ALS_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
  return ALS_OK;
}

int als_ctx_set_stream(als_ctx* ctx, void* stream) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  cudaStream_t next = static_cast<cudaStream_t>(stream);
  if (next == ctx->stream && !ctx->owns_stream) return ALS_OK;
  // order the new stream behind everything queued on the old one (pool vector, staging, selection scratch)
  cudaEvent_t ev;
  ALS_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(next, ev, 0);
  cudaEventDestroy(ev);
  ALS_CUDA(ctx, e);
  if (ctx->owns_stream && ctx->stream) ALS_CUDA(ctx, cudaStreamDestroy(ctx->stream));  // released once its work is done
  ctx->stream = next;
  ctx->owns_stream = false;
  return ALS_OK;
}

int als_score(als_ctx* ctx, const void* logits, int dtype, int64_t T, int64_t N, int64_t H, int64_t W, int64_t C,
              int measure, double* scores, float* conf_map, uint8_t* label, uint8_t* mask, float threshold,
              void* stream) {
  const Shape s{T, N, H, W, C};
  ALS_TRY(check_common(ctx, logits, dtype, s, measure, label != nullptr));
  if (N > 0 && !scores) return fail(ctx, ALS_ERR_INVALID, "scores pointer is NULL");
  DeviceGuard g(ctx->device);
  if (N > 0) {
    ALS_TRY(check_device_ptr(ctx, logits, "logits"));
    ALS_TRY(check_device_ptr(ctx, scores, "scores"));
    ALS_TRY(check_device_ptr(ctx, conf_map, "conf_map"));
    ALS_TRY(check_device_ptr(ctx, label, "label"));
    ALS_TRY(check_device_ptr(ctx, mask, "mask"));
  }
  ALS_TRY(ensure_acc(ctx, N));
  return score_device(ctx, logits, dtype, s, measure, scores, nullptr, nullptr, 0, conf_map, label, mask, threshold,
                      resolve_stream(ctx, stream));
}

int als_score_host(als_ctx* ctx, const void* logits, int dtype, int64_t T, int64_t N, int64_t H, int64_t W, int64_t C,
                   int measure, double* scores, float* conf_map, uint8_t* label, uint8_t* mask, float threshold) {
  const Shape s{T, N, H, W, C};
  ALS_TRY(check_common(ctx, logits, dtype, s, measure, label != nullptr));
  if (N == 0) return ALS_OK;
  if (!scores) return fail(ctx, ALS_ERR_INVALID, "scores pointer is NULL");
  DeviceGuard g(ctx->device);
  const int es = dtype == ALS_F32 ? 4 : 2;
  const size_t img_bytes = static_cast<size_t>(s.P()) * C * es;
  const size_t per_img_all_t = img_bytes * T;
  int64_t nb_max = static_cast<int64_t>(kStageDefault / per_img_all_t);
  if (nb_max < 1) nb_max = 1;
  if (nb_max > N) nb_max = N;
  // keep every sample plane of a chunk 16-byte aligned inside the staging buffer
  ALS_TRY(ensure_stage(ctx, static_cast<size_t>(nb_max) * per_img_all_t));
  ALS_TRY(ensure_acc(ctx, nb_max));
  ALS_TRY(grow(ctx, &ctx->scores_dev, &ctx->scores_cap, N, false));
  const bool maps = conf_map || label || mask;
  const size_t P = static_cast<size_t>(s.P());
  if (maps) ALS_TRY(grow_bytes(ctx, &ctx->maps_dev, &ctx->maps_cap, static_cast<size_t>(nb_max) * P * 6));
  const unsigned char* host = static_cast<const unsigned char*>(logits);
  for (int64_t n0 = 0; n0 < N; n0 += nb_max) {
    const int64_t nb = (N - n0) < nb_max ? (N - n0) : nb_max;
    int b = 0;
    ALS_TRY(stage_chunk(ctx, host, s, es, n0, nb, &b));
    ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    float* d_conf = nullptr;
    uint8_t* d_label = nullptr;
    uint8_t* d_mask = nullptr;
    if (maps) {
      unsigned char* m = static_cast<unsigned char*>(ctx->maps_dev);
      if (conf_map) d_conf = reinterpret_cast<float*>(m);
      if (label) d_label = m + static_cast<size_t>(nb_max) * P * 4;
      if (mask) d_mask = m + static_cast<size_t>(nb_max) * P * 5;
    }
    const Shape cs{T, nb, H, W, C};
    ALS_TRY(score_device(ctx, ctx->stage[b], dtype, cs, measure, ctx->scores_dev + n0, nullptr, nullptr, 0, d_conf,
                         d_label, d_mask, threshold, ctx->stream));
    ALS_CUDA(ctx, cudaEventRecord(ctx->ev_scored[b], ctx->stream));
    if (d_conf) ALS_CUDA(ctx, cudaMemcpyAsync(conf_map + n0 * P, d_conf, nb * P * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_label) ALS_CUDA(ctx, cudaMemcpyAsync(label + n0 * P, d_label, nb * P, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_mask) ALS_CUDA(ctx, cudaMemcpyAsync(mask + n0 * P, d_mask, nb * P, cudaMemcpyDeviceToHost, ctx->stream));
  }
  ALS_CUDA(ctx, cudaMemcpyAsync(scores, ctx->scores_dev, static_cast<size_t>(N) * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx->stream));
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ALS_OK;
}

int als_score_dlpack(als_ctx* ctx, void* managed, int measure, double* scores, void* stream) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (!managed) return fail(ctx, ALS_ERR_INVALID, "DLManagedTensor is NULL");
  const AlsDLTensor& t = static_cast<AlsDLManagedTensor*>(managed)->dl_tensor;
  int dtype;
  if (t.dtype.lanes != 1) return fail(ctx, ALS_ERR_INVALID, "vector dtypes are not supported");
  if (t.dtype.code == kAlsDLFloat && t.dtype.bits == 32) dtype = ALS_F32;
  else if (t.dtype.code == kAlsDLBfloat && t.dtype.bits == 16) dtype = ALS_BF16;
  else return fail(ctx, ALS_ERR_INVALID, "logits dtype must be float32 or bfloat16 (DLPack code %d, bits %d)", t.dtype.code, t.dtype.bits);
  if (t.ndim != 4 && t.ndim != 5) return fail(ctx, ALS_ERR_INVALID, "logits must be [N,H,W,C] or [T,N,H,W,C], got ndim=%d", t.ndim);
  const int64_t* sh = t.shape;
  const int o = t.ndim - 4;
  const Shape s{o ? sh[0] : 1, sh[o], sh[o + 1], sh[o + 2], sh[o + 3]};
  if (t.strides) {  // must be dense C-order (NHWC, class innermost)
    int64_t expect = 1;
    for (int d = t.ndim - 1; d >= 0; --d) {
      if (sh[d] != 1 && t.strides[d] != expect)
        return fail(ctx, ALS_ERR_INVALID, "logits must be dense C-contiguous NHWC (stride[%d]=%lld, expected %lld)", d,
                    (long long)t.strides[d], (long long)expect);
      expect *= sh[d];
    }
  }
  const void* data = static_cast<const unsigned char*>(t.data) + t.byte_offset;
  if (!scores) return fail(ctx, ALS_ERR_INVALID, "scores pointer is NULL");
  switch (t.device.device_type) {
    case kAlsDLCUDA:
    case kAlsDLCUDAManaged: {
      if (t.device.device_type == kAlsDLCUDA && t.device.device_id != ctx->device)
        return fail(ctx, ALS_ERR_INVALID, "logits live on GPU %d but the context is bound to GPU %d", t.device.device_id,
                    ctx->device);
      ALS_TRY(check_common(ctx, data, dtype, s, measure, false));
      if (s.N == 0) return ALS_OK;
      DeviceGuard g(ctx->device);
      ALS_TRY(ensure_acc(ctx, s.N));
      // scores_dev belongs to the context's stream: nothing of it may still be in flight when `st` writes it
      ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      ALS_TRY(grow(ctx, &ctx->scores_dev, &ctx->scores_cap, s.N, false));
      // DLPack stream exchange: the producer made the tensor ready on `stream` (the stream handed to its
      // __dlpack__(stream=...)), so enqueueing there is all the ordering that is needed.
      cudaStream_t st = resolve_stream(ctx, stream);
      ALS_TRY(score_device(ctx, data, dtype, s, measure, ctx->scores_dev, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                           0.f, st));
      ALS_CUDA(ctx, cudaMemcpyAsync(scores, ctx->scores_dev, static_cast<size_t>(s.N) * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
      ALS_CUDA(ctx, cudaStreamSynchronize(st));
      return ALS_OK;
    }
    case kAlsDLCPU:
    case kAlsDLCUDAHost:
      return als_score_host(ctx, data, dtype, s.T, s.N, s.H, s.W, s.C, measure, scores, nullptr, nullptr, nullptr, 0.f);
    default:
      return fail(ctx, ALS_ERR_INVALID, "unsupported DLPack device type %d", t.device.device_type);
  }
}

}  // extern "C"

// ---- fused classifier head -----------------------------------------------------------------------

namespace {

int check_features(als_ctx* ctx, const void* features, int64_t T, int64_t N, int64_t h, int64_t w, int measure, bool want_label) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (measure < ALS_ENTROPY || measure > ALS_VARIANCE)
    return fail(ctx, ALS_ERR_UNSUPPORTED, "Uncertainty function not implemented.");
  if (ctx->head_C <= 0) return fail(ctx, ALS_ERR_STATE, "als_head_prepare has not been called");
  if (T < 1 || T > 4096) return fail(ctx, ALS_ERR_INVALID, "bad sample count T=%lld", (long long)T);
  if (measure == ALS_VARIANCE && T < 2) return fail(ctx, ALS_ERR_INVALID, "measure 'variance' needs T >= 2 Monte-Carlo samples");
  if (!als_head_supported(ctx->head_C, measure, T))
    return fail(ctx, ALS_ERR_UNSUPPORTED, "no fused-head kernel for C=%lld with T=%lld (use als_score on the logits)",
                (long long)ctx->head_C, (long long)T);
  if (N < 0 || h < 1 || w < 1) return fail(ctx, ALS_ERR_INVALID, "bad feature shape [N=%lld,h=%lld,w=%lld,16]", (long long)N, (long long)h, (long long)w);
  if (h > (1 << 20) || w > (1 << 20) || N > 0x7fffffff || T * N > 0x7fffffff)  // (sample, image) is one int32 TMA coordinate
    return fail(ctx, ALS_ERR_INVALID, "feature map too large");
  if (want_label && ctx->head_C > 256) return fail(ctx, ALS_ERR_INVALID, "uint8 pseudo_label needs C <= 256");
  if (N > 0 && !features) return fail(ctx, ALS_ERR_INVALID, "features pointer is NULL");
  if (reinterpret_cast<uintptr_t>(features) % 16 != 0) return fail(ctx, ALS_ERR_INVALID, "features must be 16-byte aligned");
  return ALS_OK;
}

// device features [T,N,h,w,16] -> fixed-point sums -> finalize (scores64 and/or pool scatter)
int score_features_device(als_ctx* ctx, const void* features, int64_t T, int64_t N, int64_t h, int64_t w, int measure, double* scores64,
                          float* pool32, const long long* example_index_dev, int64_t num_examples, float* conf_map,
                          uint8_t* label, uint8_t* mask, float threshold, cudaStream_t stream, int64_t index_base = 0) {
  if (N == 0) return ALS_OK;
  const int C = static_cast<int>(ctx->head_C);
  als::HeadPlan plan = als::plan_head(C, measure, static_cast<int>(T), ctx->num_sms);
  if (!plan.func) return fail(ctx, ALS_ERR_UNSUPPORTED, "no fused-head kernel for C=%d, T=%lld", C, (long long)T);
  if (plan.smem_bytes > ctx->max_smem) return fail(ctx, ALS_ERR_CUDA, "fused head needs %d bytes of shared memory", plan.smem_bytes);
  const long long P = 4ll * h * w;  // output pixels per image (2h x 2w)
  int shift = 62 - ceil_log2_ll(P);
  if (shift > 44) shift = 44;
  als::HeadParams p{};
  p.sp.P = P;
  p.sp.total_pixels = N * P;
  p.sp.T = static_cast<int>(T);
  p.sp.C = C;
  p.sp.measure = measure;
  p.sp.inv_log2_c = static_cast<float>(1.0 / log2(static_cast<double>(C)));
  p.sp.threshold = threshold;
  p.sp.inv_T = 1.0f / static_cast<float>(T);
  p.sp.fx_scale = ldexpf(1.0f, shift);
  p.sp.acc = ctx->acc;
  p.sp.acc_stride = ctx->acc_cap;
  p.sp.flags = ctx->flags;
  p.sp.tile_counter = ctx->tile_counter;
  p.sp.conf_map = conf_map;
  p.sp.label = label;
  p.sp.mask = mask;
  p.features = static_cast<const float*>(features);
  p.T = static_cast<int>(T);
  p.weights = ctx->head_weights;
  p.h = static_cast<int>(h);
  p.w = static_cast<int>(w);
  p.n_images = static_cast<int>(N);
  p.n_strips = static_cast<int>((w + als::kHeadTileQuads - 1) / als::kHeadTileQuads);
  // rows per unit: about 12 units per SM, 8..64 rows (one halo row is re-read per unit), evenly split.  Measured: longer
  // units win for T > 1 as well (49 rows 8.94 Gpix/s, 6 rows 8.90, 1 row 8.55 on cfg2h): a unit change drains the
  // pipeline, the static round-robin's tail is not what limits the kernel.
  long long r = (N * h * p.n_strips) / (12ll * ctx->num_sms);
  if (r < 8) r = 8;
  if (r > 64) r = 64;
  if (r > h) r = h;
  const long long nrb = (h + r - 1) / r;
  p.rows_per_unit = static_cast<int>((h + nrb - 1) / nrb);
  p.n_rowblocks = static_cast<int>((h + p.rows_per_unit - 1) / p.rows_per_unit);
  p.n_units = N * p.n_rowblocks * p.n_strips;
  p.g = als::head_geometry(C);
  ALS_TRY(scratch_begin(ctx, stream));
  ALS_CUDA(ctx, als::launch_head(plan, p, stream));
  ALS_CUDA(ctx, als::launch_finalize(ctx->acc, ctx->acc_cap, ctx->flags, ctx->tile_counter, static_cast<int>(N),
                                     ldexp(1.0, -shift) / static_cast<double>(P), scores64, pool32, example_index_dev,
                                     index_base, num_examples, stream));
  ctx->launches += 2;
  return scratch_end(ctx, stream);
}

}  // namespace

extern "C" {

int als_head_supported(int64_t C, int measure, int64_t T) {
  if (C < 2 || C > als::kHeadMaxClasses || T < 1 || T > 4096) return 0;
  return als::plan_head(static_cast<int>(C), measure, static_cast<int>(T), 1).func != nullptr ? 1 : 0;
}

int als_head_geometry(int64_t C, int32_t* geom14, int32_t* rows) {
  if (!geom14 || !rows) return fail(nullptr, ALS_ERR_INVALID, "NULL argument");
  if (C < 2 || C > 32) return fail(nullptr, ALS_ERR_INVALID, "the fused head handles 2 <= C <= 32 classes, got %lld", (long long)C);
  const als::HeadGeom g = als::head_geometry(static_cast<int>(C));
  geom14[0] = g.C;
  geom14[1] = g.CB;
  for (int o = 0; o < 4; ++o) {
    geom14[2 + o] = g.n[o];
    geom14[6 + o] = g.col0[o];
    geom14[10 + o] = g.row0[o];
  }
  *rows = g.rows;
  return ALS_OK;
}

int als_head_pack_weights(const float* kernel, int64_t C, float* out, int64_t out_floats) {
  if (!kernel || !out) return fail(nullptr, ALS_ERR_INVALID, "NULL argument");
  if (C < 2 || C > 32) return fail(nullptr, ALS_ERR_INVALID, "the fused head handles 2 <= C <= 32 classes, got %lld", (long long)C);
  const als::HeadGeom g = als::head_geometry(static_cast<int>(C));
  if (out_floats < static_cast<int64_t>(2) * 4 * g.rows * 4) return fail(nullptr, ALS_ERR_INVALID, "output buffer too small");
  als::pack_head_weights(kernel, static_cast<int>(C), out);
  return ALS_OK;
}

int als_head_prepare(als_ctx* ctx, const float* kernel, int64_t C) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (!kernel) return fail(ctx, ALS_ERR_INVALID, "kernel pointer is NULL");
  if (C < 2) return fail(ctx, ALS_ERR_INVALID, "need at least 2 classes, got C=%lld", (long long)C);
  if (!als_head_supported(C, ALS_ENTROPY, 1))
    return fail(ctx, ALS_ERR_UNSUPPORTED, "no fused-head kernel is built for C=%lld (use als_score on the logits)", (long long)C);
  DeviceGuard g(ctx->device);
  const als::HeadGeom geom = als::head_geometry(static_cast<int>(C));
  const size_t n = static_cast<size_t>(2) * 4 * geom.rows * 4;
  std::vector<float> packed(n);
  als::pack_head_weights(kernel, static_cast<int>(C), packed.data());
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->head_weights) ALS_CUDA(ctx, cudaFree(ctx->head_weights));
  ctx->head_weights = nullptr;
  ctx->head_C = 0;
  ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->head_weights), n * sizeof(float)));
  ALS_CUDA(ctx, cudaMemcpy(ctx->head_weights, packed.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  ctx->head_C = C;
  return ALS_OK;
}

int als_score_features(als_ctx* ctx, const void* features, int64_t T, int64_t N, int64_t h, int64_t w, int measure, double* scores,
                       float* conf_map, uint8_t* label, uint8_t* mask, float threshold, void* stream) {
  ALS_TRY(check_features(ctx, features, T, N, h, w, measure, label != nullptr));
  if (N == 0) return ALS_OK;
  if (!scores) return fail(ctx, ALS_ERR_INVALID, "scores pointer is NULL");
  DeviceGuard g(ctx->device);
  ALS_TRY(check_device_ptr(ctx, features, "features"));
  ALS_TRY(check_device_ptr(ctx, scores, "scores"));
  ALS_TRY(check_device_ptr(ctx, conf_map, "conf_map"));
  ALS_TRY(check_device_ptr(ctx, label, "label"));
  ALS_TRY(check_device_ptr(ctx, mask, "mask"));
  ALS_TRY(ensure_acc(ctx, N));
  return score_features_device(ctx, features, T, N, h, w, measure, scores, nullptr, nullptr, 0, conf_map, label, mask, threshold,
                               resolve_stream(ctx, stream));
}

int als_pool_score_features_batch(als_ctx* ctx, const void* features, int features_on_host, int64_t T, int64_t B, int64_t h,
                                  int64_t w, int measure, const int64_t* example_index) {
  ALS_TRY(check_features(ctx, features, T, B, h, w, measure, false));
  if (ctx->pool_n < 0) return fail(ctx, ALS_ERR_STATE, "als_pool_begin has not been called");
  if (B == 0) return ALS_OK;
  if (!example_index) return fail(ctx, ALS_ERR_INVALID, "example_index is NULL");
  for (int64_t i = 0; i < B; ++i)
    if (example_index[i] < 0 || example_index[i] >= ctx->pool_n)
      return fail(ctx, ALS_ERR_INVALID, "example_index[%lld]=%lld out of range [0, %lld)", (long long)i,
                  (long long)example_index[i], (long long)ctx->pool_n);
  DeviceGuard g(ctx->device);
  ALS_TRY(ensure_acc(ctx, B));
  const void* dev = features;
  int b = -1;
  if (features_on_host) {
    const Shape s{T, B, h, w, als::kHeadChannels};
    ALS_TRY(ensure_stage(ctx, static_cast<size_t>(s.elems()) * 4));
    ALS_TRY(stage_chunk(ctx, static_cast<const unsigned char*>(features), s, 4, 0, B, &b));
    ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    dev = ctx->stage[b];
  } else {
    ALS_TRY(check_device_ptr(ctx, features, "features"));
  }
  const long long* idx_dev = nullptr;
  int64_t idx_base = 0;
  ALS_TRY(stage_index(ctx, example_index, B, &idx_dev, &idx_base));
  ALS_TRY(score_features_device(ctx, dev, T, B, h, w, measure, nullptr, ctx->pool32, idx_dev, ctx->pool_n, nullptr, nullptr,
                                nullptr, 0.f, ctx->stream, idx_base));
  if (b >= 0) {
    ALS_CUDA(ctx, cudaEventRecord(ctx->ev_scored[b], ctx->stream));
    ALS_CUDA(ctx, cudaEventSynchronize(ctx->ev_copied[b]));
  }
  return ALS_OK;
}

}  // extern "C"

extern "C" {

// ---- rank_confidence-shaped pool API -----------------------------------------------------------

int als_pool_begin(als_ctx* ctx, int64_t num_examples) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (num_examples < 0) return fail(ctx, ALS_ERR_INVALID, "num_examples must be >= 0");
  DeviceGuard g(ctx->device);
  ALS_TRY(grow(ctx, &ctx->pool32, &ctx->pool_cap, num_examples > 0 ? num_examples : 1, false));
  // :685  np.zeros(num_examples, dtype=np.float32)
  ALS_CUDA(ctx, cudaMemsetAsync(ctx->pool32, 0, static_cast<size_t>(ctx->pool_cap) * sizeof(float), ctx->stream));
  ctx->pool_n = num_examples;
  return ALS_OK;
}

int als_pool_score_batch(als_ctx* ctx, const void* logits, int logits_on_host, int dtype, int64_t T, int64_t B,
                         int64_t H, int64_t W, int64_t C, int measure, const int64_t* example_index) {
  const Shape s{T, B, H, W, C};
  ALS_TRY(check_common(ctx, logits, dtype, s, measure, false));
  if (ctx->pool_n < 0) return fail(ctx, ALS_ERR_STATE, "als_pool_begin has not been called");
  if (B == 0) return ALS_OK;
  if (!example_index) return fail(ctx, ALS_ERR_INVALID, "example_index is NULL");
  for (int64_t i = 0; i < B; ++i)
    if (example_index[i] < 0 || example_index[i] >= ctx->pool_n)
      return fail(ctx, ALS_ERR_INVALID, "example_index[%lld]=%lld out of range [0, %lld)", (long long)i,
                  (long long)example_index[i], (long long)ctx->pool_n);
  DeviceGuard g(ctx->device);
  ALS_TRY(ensure_acc(ctx, B));
  static_assert(sizeof(long long) == sizeof(int64_t), "");
  const void* dev_logits = logits;
  int b = -1;
  if (logits_on_host) {
    const int es = dtype == ALS_F32 ? 4 : 2;
    ALS_TRY(ensure_stage(ctx, static_cast<size_t>(s.elems()) * es));
    ALS_TRY(stage_chunk(ctx, static_cast<const unsigned char*>(logits), s, es, 0, B, &b));
    ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    dev_logits = ctx->stage[b];
  } else {
    ALS_TRY(check_device_ptr(ctx, logits, "logits"));
  }
  const long long* idx_dev = nullptr;
  int64_t idx_base = 0;
  ALS_TRY(stage_index(ctx, example_index, B, &idx_dev, &idx_base));
  ALS_TRY(score_device(ctx, dev_logits, dtype, s, measure, nullptr, ctx->pool32, idx_dev, ctx->pool_n, nullptr,
                       nullptr, nullptr, 0.f, ctx->stream, idx_base));
  if (b >= 0) {
    ALS_CUDA(ctx, cudaEventRecord(ctx->ev_scored[b], ctx->stream));
    // "returns once staged": the caller may reuse its host buffer after this call
    ALS_CUDA(ctx, cudaEventSynchronize(ctx->ev_copied[b]));
  }
  return ALS_OK;
}

int als_pool_scores(als_ctx* ctx, float* out, int64_t num_examples) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (ctx->pool_n < 0) return fail(ctx, ALS_ERR_STATE, "als_pool_begin has not been called");
  if (num_examples != ctx->pool_n) return fail(ctx, ALS_ERR_INVALID, "num_examples=%lld but the pool holds %lld",
                                               (long long)num_examples, (long long)ctx->pool_n);
  if (num_examples == 0) return ALS_OK;
  if (!out) return fail(ctx, ALS_ERR_INVALID, "out is NULL");
  DeviceGuard g(ctx->device);
  ALS_CUDA(ctx, cudaMemcpyAsync(out, ctx->pool32, static_cast<size_t>(num_examples) * sizeof(float),
                                cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ALS_OK;
}

}  // extern "C"

namespace als {

// :705  the ids must index the pool vector, and be unique (duplicates would make the k-smallest set ill-defined)
int validate_unlabelled(als_ctx* ctx, const int64_t* unlabelled, int64_t M) {
  std::vector<uint64_t> seen(static_cast<size_t>((ctx->pool_n + 63) / 64), 0);
  for (int64_t i = 0; i < M; ++i) {
    const int64_t id = unlabelled[i];
    if (id < 0 || id >= ctx->pool_n)
      return fail(ctx, ALS_ERR_INVALID, "unlabelled[%lld]=%lld out of range [0, %lld)", (long long)i, (long long)id,
                  (long long)ctx->pool_n);
    uint64_t& w = seen[static_cast<size_t>(id >> 6)];
    const uint64_t bit = 1ull << (id & 63);
    if (w & bit) return fail(ctx, ALS_ERR_INVALID, "unlabelled[%lld]=%lld appears twice (ids must be unique)", (long long)i, (long long)id);
    w |= bit;
  }
  return ALS_OK;
}

// unlabelled (host) -> pinned mirror -> ctx->sel_ids on the context's stream
int upload_unlabelled(als_ctx* ctx, const int64_t* unlabelled, int64_t M) {
  ALS_TRY(grow(ctx, &ctx->sel_ids, &ctx->sel_ids_cap, M, false));
  size_t cap = static_cast<size_t>(ctx->sel_ids_host_cap) * 8;
  void* hp = ctx->sel_ids_host;
  ALS_TRY(grow_pinned(ctx, &hp, &cap, static_cast<size_t>(M) * 8));
  ctx->sel_ids_host = static_cast<long long*>(hp);
  ctx->sel_ids_host_cap = static_cast<int64_t>(cap / 8);
  // the previous selection's copy out of this pinned buffer completed before that call returned (it synchronises)
  memcpy(ctx->sel_ids_host, unlabelled, static_cast<size_t>(M) * 8);
  // on the copy stream: the upload overlaps the scoring kernels still running on the context's stream, only the
  // select launch waits for it (the previous selection has been waited for by the host, so sel_ids is free)
  ALS_CUDA(ctx, cudaMemcpyAsync(ctx->sel_ids, ctx->sel_ids_host, static_cast<size_t>(M) * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
  ALS_CUDA(ctx, cudaEventRecord(ctx->ev_unl, ctx->copy_stream));
  ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_unl, 0));
  return ALS_OK;
}

// Result block of a selection, device and pinned host mirror: {count, status} | ids[k] | keys[k] | uconf[M]
SelectBlock select_block(int64_t k, int64_t M) {
  SelectBlock b;
  b.off_ids = 16;
  b.off_keys = b.off_ids + static_cast<size_t>(k) * 8;
  b.off_uconf = (b.off_keys + static_cast<size_t>(k) * 4 + 15) & ~static_cast<size_t>(15);
  b.bytes = b.off_uconf + static_cast<size_t>(M) * 4;
  return b;
}

int ensure_select_block(als_ctx* ctx, const SelectBlock& b, int64_t kmax) {
  void* p = ctx->sel_out;
  ALS_TRY(grow_bytes(ctx, &p, &ctx->sel_out_cap, 64));  // device scratch for the {count, status} words of helper launches
  ctx->sel_out = static_cast<unsigned char*>(p);
  p = ctx->sel_out_host;
  ALS_TRY(grow_pinned(ctx, &p, &ctx->sel_out_host_cap, b.bytes));
  ctx->sel_out_host = static_cast<unsigned char*>(p);
  if (kmax > kSelFusedMaxK && kmax > ctx->sel_tmp_cap) {
    ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->sel_tmp_keys) ALS_CUDA(ctx, cudaFree(ctx->sel_tmp_keys));
    if (ctx->sel_tmp_ids) ALS_CUDA(ctx, cudaFree(ctx->sel_tmp_ids));
    ctx->sel_tmp_keys = nullptr;
    ctx->sel_tmp_ids = nullptr;
    ctx->sel_tmp_cap = 0;
    int64_t cap = 4096;
    while (cap < kmax) cap *= 2;
    ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sel_tmp_keys), static_cast<size_t>(cap) * 4));
    ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sel_tmp_ids), static_cast<size_t>(cap) * 8));
    ctx->sel_tmp_cap = cap;
  }
  return ALS_OK;
}

// The result block lives in PINNED HOST memory (device-accessible under unified addressing): the select kernel writes
// count / ids / keys / unlabelled_confidence straight over PCIe, so a pass ends with one kernel and a stream
// synchronisation -- no separate device->host copy operation.
SelectOut select_out_of(als_ctx* ctx, const SelectBlock& b, int64_t M) {
  SelectOut o{};
  o.count = reinterpret_cast<long long*>(ctx->sel_out_host);
  o.ids = reinterpret_cast<long long*>(ctx->sel_out_host + b.off_ids);
  o.keys = reinterpret_cast<float*>(ctx->sel_out_host + b.off_keys);
  o.pad_base = -1;
  o.uconf = reinterpret_cast<float*>(ctx->sel_out_host + b.off_uconf);
  o.uconf_ids = ctx->sel_ids;
  o.uconf_M = M;
  o.uconf_pool = ctx->pool32;
  return o;
}

// Wait for the select kernel (its writes to the pinned block are visible once the stream is idle), then unpack.
int fetch_select_block(als_ctx* ctx, const SelectBlock& b, int64_t k, int64_t M, int64_t* out_ids, float* out_unlabelled_conf,
                       int64_t* out_count) {
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const long long* hdr = reinterpret_cast<const long long*>(ctx->sel_out_host);
  if (hdr[1] != 0)
    return fail(ctx, ALS_ERR_INVALID, "the ranks' shards [shard_lo, shard_hi) do not tile [0, %lld): every example needs exactly one owner",
                (long long)ctx->pool_n);
  const int64_t n = hdr[0] < k ? hdr[0] : k;
  if (n > 0) memcpy(out_ids, ctx->sel_out_host + b.off_ids, static_cast<size_t>(n) * 8);
  if (out_unlabelled_conf && M > 0) memcpy(out_unlabelled_conf, ctx->sel_out_host + b.off_uconf, static_cast<size_t>(M) * 4);
  *out_count = n;
  return ALS_OK;
}

}  // namespace als

extern "C" {

int als_pool_select(als_ctx* ctx, const int64_t* unlabelled, int64_t M, int64_t selection_size, int64_t* out_ids,
                    float* out_unlabelled_conf, int64_t* out_count) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (ctx->pool_n < 0) return fail(ctx, ALS_ERR_STATE, "als_pool_begin has not been called");
  if (M < 0) return fail(ctx, ALS_ERR_INVALID, "M must be >= 0");
  if (!out_count) return fail(ctx, ALS_ERR_INVALID, "out_count is NULL");
  *out_count = 0;
  if (M == 0) return ALS_OK;
  if (!unlabelled) return fail(ctx, ALS_ERR_INVALID, "unlabelled is NULL");
  ALS_TRY(als::validate_unlabelled(ctx, unlabelled, M));
  // :707-708  selection_size = min(len(unlabelled), selection_size); negative sizes never reach here (:779)
  const int64_t k = selection_size < 0 ? 0 : (selection_size < M ? selection_size : M);
  if (k > 0 && !out_ids) return fail(ctx, ALS_ERR_INVALID, "out_ids is NULL");
  DeviceGuard g(ctx->device);
  const als::SelectBlock blk = als::select_block(k, M);
  ALS_TRY(als::ensure_select_block(ctx, blk, k));
  ALS_TRY(als::upload_unlabelled(ctx, unlabelled, M));
  // one launch: gather confidence[unlabelled] (:705), radix-select the k lowest (:710-712), order them, pack
  als::SelectSrc src{};
  src.mode = 1;
  src.ids = ctx->sel_ids;
  src.pool = ctx->pool32;
  src.lo = INT64_MIN;
  src.hi = INT64_MAX;
  const als::SelectOut out = als::select_out_of(ctx, blk, M);
  int nl = 0;
  ALS_CUDA(ctx, als::launch_select(src, M, k, als::ScatterDesc{}, als::ExportDesc{}, out, ctx->sel_tmp_keys, ctx->sel_tmp_ids,
                                   ctx->stream, &nl));
  ctx->launches += nl;
  return als::fetch_select_block(ctx, blk, k, M, out_ids, out_unlabelled_conf, out_count);
}

// :682-715 in one call, for a pool that arrives as one tensor.  Same results as begin + score_batch + select; the point
// is the order of the host work: the scoring launch is queued FIRST, and everything the selection needs from the host
// (validating and uploading `unlabelled`, growing buffers) happens while the GPU scores, so a small pool -- one scoring
// launch of a few hundred microseconds -- is followed by the select launch without a host gap.
int als_rank_pool(als_ctx* ctx, const void* logits, int logits_on_host, int dtype, int64_t T, int64_t N, int64_t H, int64_t W,
                  int64_t C, int measure, const int64_t* example_index, int64_t num_examples, const int64_t* unlabelled,
                  int64_t M, int64_t selection_size, int64_t* out_ids, float* out_unlabelled_conf, int64_t* out_count) {
  const Shape s{T, N, H, W, C};
  ALS_TRY(check_common(ctx, logits, dtype, s, measure, false));
  if (num_examples < 0) return fail(ctx, ALS_ERR_INVALID, "num_examples must be >= 0");
  if (M < 0) return fail(ctx, ALS_ERR_INVALID, "M must be >= 0");
  if (!out_count) return fail(ctx, ALS_ERR_INVALID, "out_count is NULL");
  *out_count = 0;
  if (M > 0 && !unlabelled) return fail(ctx, ALS_ERR_INVALID, "unlabelled is NULL");
  if (!example_index && N > num_examples) return fail(ctx, ALS_ERR_INVALID, "the pool holds %lld examples but %lld images were given",
                                                      (long long)num_examples, (long long)N);
  const int64_t k = selection_size < 0 ? 0 : (selection_size < M ? selection_size : M);  // :707-708
  if (k > 0 && !out_ids) return fail(ctx, ALS_ERR_INVALID, "out_ids is NULL");
  DeviceGuard g(ctx->device);
  ALS_TRY(als_pool_begin(ctx, num_examples));  // :684-685
  // A small `unlabelled` goes up BEFORE the scoring launch (a few microseconds of host work), so that no stream operation
  // sits between the scoring and the select launch and the latter keeps its programmatic-dependent-launch overlap; a
  // large one is validated and uploaded while the GPU scores (als_pool_select below).
  const bool early_upload = M > 0 && M <= 4096;
  als::SelectBlock blk = als::select_block(k, M);
  if (early_upload) {
    ALS_TRY(als::validate_unlabelled(ctx, unlabelled, M));
    ALS_TRY(als::ensure_select_block(ctx, blk, k));
    ALS_TRY(als::upload_unlabelled(ctx, unlabelled, M));
  }
  if (N > 0) {
    std::vector<int64_t> iota;
    const int64_t* idx = example_index;
    if (!idx) {  // examples 0 .. N-1 in order
      iota.resize(static_cast<size_t>(N));
      for (int64_t i = 0; i < N; ++i) iota[static_cast<size_t>(i)] = i;
      idx = iota.data();
    }
    ALS_TRY(als_pool_score_batch(ctx, logits, logits_on_host, dtype, T, N, H, W, C, measure, idx));  // :697-700
  }
  if (!early_upload)  // the GPU is scoring; the host prepares the selection meanwhile
    return als_pool_select(ctx, unlabelled, M, selection_size, out_ids, out_unlabelled_conf, out_count);  // :705-715
  als::SelectSrc src{};
  src.mode = 1;
  src.ids = ctx->sel_ids;
  src.pool = ctx->pool32;
  src.lo = INT64_MIN;
  src.hi = INT64_MAX;
  const als::SelectOut out = als::select_out_of(ctx, blk, M);
  int nl = 0;
  ALS_CUDA(ctx, als::launch_select(src, M, k, als::ScatterDesc{}, als::ExportDesc{}, out, ctx->sel_tmp_keys, ctx->sel_tmp_ids,
                                   ctx->stream, &nl));  // :705-714
  ctx->launches += nl;
  return als::fetch_select_block(ctx, blk, k, M, out_ids, out_unlabelled_conf, out_count);
}

int als_select_smallest(als_ctx* ctx, const float* keys, const int64_t* ids, int64_t M, int64_t k, float* out_keys,
                        int64_t* out_ids, void* stream) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (M < 0 || k < 0) return fail(ctx, ALS_ERR_INVALID, "M and k must be >= 0");
  const int64_t kk = k < M ? k : M;
  if (kk == 0) return ALS_OK;
  if (!keys || !ids || !out_keys || !out_ids) return fail(ctx, ALS_ERR_INVALID, "NULL pointer");
  DeviceGuard g(ctx->device);
  ALS_TRY(check_device_ptr(ctx, keys, "keys"));
  ALS_TRY(check_device_ptr(ctx, ids, "ids"));
  ALS_TRY(check_device_ptr(ctx, out_keys, "out_keys"));
  ALS_TRY(check_device_ptr(ctx, out_ids, "out_ids"));
  const als::SelectBlock blk = als::select_block(0, 0);  // only the {count, status} header is used
  ALS_TRY(als::ensure_select_block(ctx, blk, kk));
  cudaStream_t st = resolve_stream(ctx, stream);
  ALS_TRY(scratch_begin(ctx, st));  // the header words and the large-k scratch are per context
  als::SelectSrc src{};
  src.mode = 0;
  src.keys = keys;
  src.ids = reinterpret_cast<const long long*>(ids);
  als::SelectOut out{};
  out.count = reinterpret_cast<long long*>(ctx->sel_out);
  out.keys = out_keys;
  out.ids = reinterpret_cast<long long*>(out_ids);
  out.pad_base = -1;
  int nl = 0;
  ALS_CUDA(ctx, als::launch_select(src, M, kk, als::ScatterDesc{}, als::ExportDesc{}, out, ctx->sel_tmp_keys, ctx->sel_tmp_ids, st, &nl));
  ctx->launches += nl;
  return scratch_end(ctx, st);
}

// ---- synthetic pool / bench helpers ----------------------------------------------------------------

int als_synth_logits(als_ctx* ctx, void* out, int dtype, int64_t T, int64_t n0, int64_t n_imgs, int64_t H, int64_t W,
                     int64_t C, uint64_t seed, int mc, void* stream) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (dtype != ALS_F32 && dtype != ALS_BF16) return fail(ctx, ALS_ERR_INVALID, "dtype must be float32 or bfloat16");
  if (T < 1 || n0 < 0 || n_imgs < 0 || H < 1 || W < 1 || C < 1 || C > 4096)
    return fail(ctx, ALS_ERR_INVALID, "bad synth shape");
  if (n_imgs == 0) return ALS_OK;
  DeviceGuard g(ctx->device);
  ALS_TRY(check_device_ptr(ctx, out, "out"));
  if (!out) return fail(ctx, ALS_ERR_INVALID, "out is NULL");
  cudaStream_t st = resolve_stream(ctx, stream);
  ALS_CUDA(ctx, als::launch_synth(out, dtype, T, n0, n_imgs, H * W, static_cast<int>(C), seed, mc, st));
  ctx->launches += 1;
  return ALS_OK;
}

int als_flush_l2(als_ctx* ctx, void* stream) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  if (!ctx->flush_buf) {
    ctx->flush_bytes = 512ull << 20;
    ALS_CUDA(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  cudaStream_t st = resolve_stream(ctx, stream);
  ALS_CUDA(ctx, als::launch_fill(ctx->flush_buf, ctx->flush_bytes, st));
  return ALS_OK;
}

int als_describe_launch(als_ctx* ctx, int dtype, int64_t T, int64_t N, int64_t H, int64_t W, int64_t C, int measure,
                        char* name, int* grid, int* block, int* smem_bytes, int* stages, int* tile_pixels) {
  const Shape s{T, N, H, W, C};
  ALS_TRY(check_common(ctx, reinterpret_cast<const void*>(16), dtype, s, measure, false));
  DeviceGuard g(ctx->device);
  const int es = dtype == ALS_F32 ? 4 : 2;
  const bool aligned = (T == 1) || ((N * s.P() * C * es) % 16 == 0);
  als::LaunchPlan plan = als::plan_score(dtype, static_cast<int>(C), measure, static_cast<int>(T), N * s.P(), aligned,
                                         ctx->num_sms, ctx->max_smem);
  if (name) snprintf(name, 128, "%s dtype=%s C=%lld lanes/pixel=%d pixels/thread=%d", plan.name, dtype == ALS_F32 ? "f32" : "bf16",
                     (long long)C, plan.lanes_per_pixel, plan.pixels_per_thread);
  if (grid) *grid = plan.grid;
  if (block) *block = plan.block;
  if (smem_bytes) *smem_bytes = plan.smem_bytes;
  if (stages) *stages = plan.stages;
  if (tile_pixels) *tile_pixels = plan.tile_pixels;
  return ALS_OK;
}

int als_describe_head_launch(als_ctx* ctx, int64_t T, int measure, char* name, int* grid, int* block, int* smem_bytes) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (ctx->head_C <= 0) return fail(ctx, ALS_ERR_STATE, "als_head_prepare has not been called");
  const als::HeadPlan plan = als::plan_head(static_cast<int>(ctx->head_C), measure, static_cast<int>(T), ctx->num_sms);
  if (!plan.func) return fail(ctx, ALS_ERR_UNSUPPORTED, "no fused-head kernel for C=%lld, T=%lld", (long long)ctx->head_C, (long long)T);
  if (name) snprintf(name, 128, "%s C=%lld: tcgen05 split-TF32 UMMA 128xNx8 + TMEM epilogue", plan.name, (long long)ctx->head_C);
  if (grid) *grid = plan.grid;
  if (block) *block = plan.block;
  if (smem_bytes) *smem_bytes = plan.smem_bytes;
  return ALS_OK;
}

// ---- streamed Monte-Carlo accumulation ---------------------------------------------------------------

int als_mc_begin(als_ctx* ctx, int dtype, int64_t N, int64_t H, int64_t W, int64_t C, uint8_t* label) {
  const Shape s{1, N, H, W, C};
  ALS_TRY(check_common(ctx, reinterpret_cast<const void*>(16), dtype, s, ALS_ENTROPY, label != nullptr));
  DeviceGuard g(ctx->device);
  ALS_TRY(check_device_ptr(ctx, label, "label"));
  ctx->mc_samples = -1;
  const als::McPlan plan = als::plan_mc(dtype, static_cast<int>(C), N * s.P(), true, ctx->num_sms, ctx->max_smem);
  const als::McPlan generic = als::plan_mc(dtype, static_cast<int>(C), N * s.P(), false, ctx->num_sms, ctx->max_smem);
  const long long need = plan.state_floats > generic.state_floats ? plan.state_floats : generic.state_floats;
  if (static_cast<size_t>(need) > ctx->mc_state_cap) {
    ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->mc_state) ALS_CUDA(ctx, cudaFree(ctx->mc_state));
    ctx->mc_state = nullptr;
    ctx->mc_state_cap = 0;
    ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->mc_state), static_cast<size_t>(need) * sizeof(float)));
    ctx->mc_state_cap = static_cast<size_t>(need);
  }
  ALS_TRY(ensure_acc(ctx, N));
  ctx->mc_dtype = dtype;
  ctx->mc_N = N;
  ctx->mc_H = H;
  ctx->mc_W = W;
  ctx->mc_C = C;
  ctx->mc_label = label;
  ctx->mc_samples = 0;
  ctx->mc_tiled = -1;
  return ALS_OK;
}

int als_mc_add_sample(als_ctx* ctx, const void* logits, int logits_on_host) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (ctx->mc_samples < 0) return fail(ctx, ALS_ERR_STATE, "als_mc_begin has not been called");
  if (ctx->mc_N == 0) { ++ctx->mc_samples; return ALS_OK; }
  if (!logits) return fail(ctx, ALS_ERR_INVALID, "logits pointer is NULL");
  if (ctx->mc_samples >= (1 << 24)) return fail(ctx, ALS_ERR_INVALID, "too many samples");
  DeviceGuard g(ctx->device);
  const Shape s{1, ctx->mc_N, ctx->mc_H, ctx->mc_W, ctx->mc_C};
  const int es = ctx->mc_dtype == ALS_F32 ? 4 : 2;
  const void* dev = logits;
  int b = -1, d2d = -1;
  if (logits_on_host) {
    ALS_TRY(ensure_stage(ctx, static_cast<size_t>(s.elems()) * es));
    ALS_TRY(stage_chunk(ctx, static_cast<const unsigned char*>(logits), s, es, 0, s.N, &b));
    ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    dev = ctx->stage[b];
  } else {
    ALS_TRY(check_device_ptr(ctx, logits, "logits"));
  }
  // One state layout for the whole accumulation (the tiled one whenever a specialised kernel exists).  The bulk copies
  // of the tiled kernel need a 16-byte aligned sample: a misaligned device sample (e.g. slice t of an odd-sized
  // [T,N,H,W,C] stack) is first copied, device to device, into the aligned staging buffer -- a rare path.
  als::McPlan plan = als::plan_mc(ctx->mc_dtype, static_cast<int>(s.C), s.N * s.P(), true, ctx->num_sms, ctx->max_smem);
  ctx->mc_tiled = plan.tiled ? 1 : 0;
  if (plan.tiled && reinterpret_cast<uintptr_t>(dev) % 16 != 0) {
    const size_t bytes = static_cast<size_t>(s.elems()) * es;
    ALS_TRY(ensure_stage(ctx, bytes));
    const int sb = ctx->stage_next;
    ctx->stage_next ^= 1;
    ALS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_scored[sb], 0));
    ALS_CUDA(ctx, cudaMemcpyAsync(ctx->stage[sb], dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    dev = ctx->stage[sb];
    d2d = sb;  // ev_scored[sb] is recorded below so that the next user of this staging buffer waits for the kernel
  }
  als::ScoreParams p{};
  p.logits = dev;
  p.total_pixels = s.N * s.P();
  p.P = s.P();
  p.T = 1;
  p.C = static_cast<int>(s.C);
  p.tile_counter = ctx->tile_counter;
  p.done_counter = reinterpret_cast<unsigned int*>(ctx->tile_counter + 1);
  p.acc = ctx->acc;
  p.acc_stride = ctx->acc_cap;
  p.flags = ctx->flags;
  p.label = ctx->mc_samples == 0 ? ctx->mc_label : nullptr;
  ALS_TRY(scratch_begin(ctx, ctx->stream));
  ALS_CUDA(ctx, als::launch_mc_update(plan, ctx->mc_dtype, p, ctx->mc_state, static_cast<int>(ctx->mc_samples), ctx->stream));
  ctx->launches += 1;
  ALS_TRY(scratch_end(ctx, ctx->stream));
  if (d2d >= 0) ALS_CUDA(ctx, cudaEventRecord(ctx->ev_scored[d2d], ctx->stream));
  if (b >= 0) {
    ALS_CUDA(ctx, cudaEventRecord(ctx->ev_scored[b], ctx->stream));
    ALS_CUDA(ctx, cudaEventSynchronize(ctx->ev_copied[b]));  // "returns once staged"
  }
  ++ctx->mc_samples;
  return ALS_OK;
}

int als_mc_finish(als_ctx* ctx, int measure, double* scores, const int64_t* example_index, float* conf_map, uint8_t* mask,
                  float threshold) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (ctx->mc_samples < 0) return fail(ctx, ALS_ERR_STATE, "als_mc_begin has not been called");
  if (measure < ALS_ENTROPY || measure > ALS_VARIANCE) return fail(ctx, ALS_ERR_UNSUPPORTED, "Uncertainty function not implemented.");
  const int64_t T = ctx->mc_samples;
  if (T < 1) return fail(ctx, ALS_ERR_STATE, "no sample has been added");
  if (measure == ALS_VARIANCE && T < 2) return fail(ctx, ALS_ERR_INVALID, "measure 'variance' needs T >= 2 Monte-Carlo samples");
  const Shape s{1, ctx->mc_N, ctx->mc_H, ctx->mc_W, ctx->mc_C};
  ctx->mc_samples = -1;
  if (s.N == 0) return ALS_OK;
  if (example_index && ctx->pool_n < 0) return fail(ctx, ALS_ERR_STATE, "als_pool_begin has not been called");
  if (example_index)
    for (int64_t i = 0; i < s.N; ++i)
      if (example_index[i] < 0 || example_index[i] >= ctx->pool_n)
        return fail(ctx, ALS_ERR_INVALID, "example_index[%lld]=%lld out of range [0, %lld)", (long long)i,
                    (long long)example_index[i], (long long)ctx->pool_n);
  DeviceGuard g(ctx->device);
  ALS_TRY(check_device_ptr(ctx, scores, "scores"));
  ALS_TRY(check_device_ptr(ctx, conf_map, "conf_map"));
  ALS_TRY(check_device_ptr(ctx, mask, "mask"));
  const long long* idx_dev = nullptr;
  int64_t idx_base = 0;
  if (example_index) ALS_TRY(stage_index(ctx, example_index, s.N, &idx_dev, &idx_base));
  const als::McPlan plan = als::plan_mc(ctx->mc_dtype, static_cast<int>(s.C), s.N * s.P(), ctx->mc_tiled == 1, ctx->num_sms,
                                        ctx->max_smem);
  const long long P = s.P();
  const int shift = fx_shift_for(ctx->mc_dtype, P);
  als::ScoreParams p{};
  p.total_pixels = s.N * P;
  p.P = P;
  p.T = static_cast<int>(T);
  p.C = static_cast<int>(s.C);
  p.measure = measure;
  p.inv_log2_c = static_cast<float>(1.0 / log2(static_cast<double>(s.C)));
  p.threshold = threshold;
  p.inv_T = 1.0f / static_cast<float>(T);
  p.fx_scale = ldexpf(1.0f, shift);
  p.acc = ctx->acc;
  p.acc_stride = ctx->acc_cap;
  p.flags = ctx->flags;
  p.tile_counter = ctx->tile_counter;
  p.conf_map = conf_map;
  p.mask = mask;
  ALS_TRY(scratch_begin(ctx, ctx->stream));
  ALS_CUDA(ctx, als::launch_mc_finish(plan, ctx->mc_dtype, p, ctx->mc_state, ctx->stream));
  ALS_CUDA(ctx, als::launch_finalize(ctx->acc, ctx->acc_cap, ctx->flags, ctx->tile_counter, static_cast<int>(s.N),
                                     ldexp(1.0, -shift) / static_cast<double>(P), scores, example_index ? ctx->pool32 : nullptr,
                                     idx_dev, idx_base, example_index ? ctx->pool_n : 0, ctx->stream));
  ctx->launches += 2;
  return scratch_end(ctx, ctx->stream);
}

}  // extern "C"
