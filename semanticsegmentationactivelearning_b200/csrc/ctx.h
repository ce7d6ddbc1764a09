// Internal: the context object behind include/alscore.h and the helpers its translation units share.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/alscore.h"
#include "select.cuh"

namespace als {
class HostStager;  // stage.cu: parallel staging of pageable host memory
}

struct als_ctx {
  int device = 0;
  int num_sms = 148;
  int max_smem = 227 * 1024;
  cudaStream_t stream = nullptr;  // the context's stream (own, or the caller's after als_ctx_set_stream)
  bool owns_stream = true;
  cudaStream_t copy_stream = nullptr;
  std::string error;
  int64_t launches = 0;
  // per-image fixed-point accumulators: one set, shared by every scoring entry.  ev_scratch marks the end of the
  // last launch sequence that used them; a call on another stream waits for it first (order_scratch).
  long long* acc = nullptr;
  unsigned int* flags = nullptr;
  unsigned long long* tile_counter = nullptr;
  int64_t acc_cap = 0;
  cudaEvent_t ev_scratch = nullptr;
  cudaEvent_t ev_unl = nullptr;  // unlabelled ids uploaded (copy stream)
  // optional timing of the scoring launches (als_ctx_enable_timing): events around the last scoring launch sequence
  bool timing = false;
  bool timing_valid = false;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  cudaStream_t scratch_stream = nullptr;
  bool scratch_used = false;
  // device scratch for scores / indices
  double* scores_dev = nullptr;
  int64_t scores_cap = 0;
  long long* index_dev = nullptr;
  int64_t index_cap = 0;
  // pool state (rank_confidence)
  float* pool32 = nullptr;
  int64_t pool_n = -1;
  int64_t pool_cap = 0;
  // selection: unlabelled ids (device + pinned host mirror), packed result block (device + pinned host mirror)
  long long* sel_ids = nullptr;
  int64_t sel_ids_cap = 0;
  long long* sel_ids_host = nullptr;
  int64_t sel_ids_host_cap = 0;
  unsigned char* sel_out = nullptr;
  size_t sel_out_cap = 0;
  unsigned char* sel_out_host = nullptr;
  size_t sel_out_host_cap = 0;
  float* sel_tmp_keys = nullptr;
  long long* sel_tmp_ids = nullptr;
  int64_t sel_tmp_cap = 0;
  // host -> device staging (double buffered)
  void* stage[2] = {nullptr, nullptr};
  size_t stage_cap = 0;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};
  cudaEvent_t ev_scored[2] = {nullptr, nullptr};
  int stage_next = 0;
  als::HostStager* stager = nullptr;  // created on the first large pageable source
  // per-pixel output staging for the host path
  void* maps_dev = nullptr;
  size_t maps_cap = 0;
  // fused classifier head: packed split-TF32 weights of `Final` (als_head_prepare)
  float* head_weights = nullptr;
  int64_t head_C = 0;
  // streamed Monte-Carlo accumulation (als_mc_*): Welford state resident in HBM between samples
  float* mc_state = nullptr;
  size_t mc_state_cap = 0;  // floats
  int mc_dtype = 0;
  int64_t mc_N = 0, mc_H = 0, mc_W = 0, mc_C = 0;
  int64_t mc_samples = -1;  // -1: no accumulation open
  int mc_tiled = -1;        // -1 undecided, 1 tiled kernels / layout, 0 generic
  uint8_t* mc_label = nullptr;
  // multi-GPU exchange (comm.cu): NCCL communicator + the all-gather records
  void* comm = nullptr;  // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  unsigned char* xchg_send = nullptr;
  unsigned char* xchg_recv = nullptr;
  size_t xchg_send_cap = 0, xchg_recv_cap = 0;
  // L2 flush scratch
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
};

namespace als {

int fail(als_ctx* ctx, int code, const char* fmt, ...);

#define ALS_CUDA(ctx, call)                                                                                  \
  do {                                                                                                       \
    cudaError_t _e = (call);                                                                                 \
    if (_e != cudaSuccess) {                                                                                 \
      (void)cudaGetLastError();                                                                              \
      return als::fail(ctx, _e == cudaErrorMemoryAllocation ? ALS_ERR_NOMEM : ALS_ERR_CUDA, "%s failed: %s", \
                       #call, cudaGetErrorString(_e));                                                       \
    }                                                                                                        \
  } while (0)

#define ALS_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != ALS_OK) return _rc; \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Device buffer that only grows (power-of-two steps); waits for the context's stream before freeing the old one.
template <typename T>
int grow(als_ctx* ctx, T** ptr, int64_t* cap, int64_t need, bool zero) {
  if (need <= *cap) return ALS_OK;
  int64_t n = *cap > 0 ? *cap : 1024;
  while (n < need) n *= 2;
  ALS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*ptr) ALS_CUDA(ctx, cudaFree(*ptr));
  *ptr = nullptr;
  *cap = 0;
  ALS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(ptr), static_cast<size_t>(n) * sizeof(T)));
  if (zero) ALS_CUDA(ctx, cudaMemsetAsync(*ptr, 0, static_cast<size_t>(n) * sizeof(T), ctx->stream));
  *cap = n;
  return ALS_OK;
}
int grow_bytes(als_ctx* ctx, void** ptr, size_t* cap, size_t need);
int grow_pinned(als_ctx* ctx, void** ptr, size_t* cap, size_t need);

// `stream` argument of the public entries: a cudaStream_t; NULL = the legacy default stream; ALS_STREAM_CTX = ctx->stream.
cudaStream_t resolve_stream(als_ctx* ctx, void* stream);
// Serialise users of the shared accumulator set across streams (see als_ctx::ev_scratch).
int scratch_begin(als_ctx* ctx, cudaStream_t st);
int scratch_end(als_ctx* ctx, cudaStream_t st);

int check_device_ptr(als_ctx* ctx, const void* p, const char* what);

// Host -> device copy on the context's copy stream (stage.cu): pinned sources directly, large pageable sources through
// the parallel bounce-buffer pipeline.  On return the source may be reused only once ev_copied has fired (pinned) --
// callers record and wait for it as before.
int stage_copy(als_ctx* ctx, void* dst, const void* src, size_t bytes);
void stage_destroy(als_ctx* ctx);

// ---- pieces of the :705-715 selection shared by als_pool_select (capi.cu) and the multi-GPU exchange (comm.cu) ----
int validate_unlabelled(als_ctx* ctx, const int64_t* unlabelled, int64_t M);
int upload_unlabelled(als_ctx* ctx, const int64_t* unlabelled, int64_t M);  // -> ctx->sel_ids, on ctx->stream
struct SelectBlock {  // byte layout of the packed result: {count, status} | ids[k] | keys[k] | uconf[M]
  size_t off_ids, off_keys, off_uconf, bytes;
};
SelectBlock select_block(int64_t k, int64_t M);
int ensure_select_block(als_ctx* ctx, const SelectBlock& b, int64_t kmax);
SelectOut select_out_of(als_ctx* ctx, const SelectBlock& b, int64_t M);
int fetch_select_block(als_ctx* ctx, const SelectBlock& b, int64_t k, int64_t M, int64_t* out_ids, float* out_unlabelled_conf,
                       int64_t* out_count);

}  // namespace als
