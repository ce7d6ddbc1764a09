// Multi-GPU exchange of the sharded pool pass: NCCL communicator behind the C ABI + als_pool_select_global.
//
// The reference has no collective; the pool shards by image, so the only exchange of a pass is
// the one this file implements: per-rank candidates + score slices, ONE ncclAllGather over NVLink, merged on the device
// (select.cu).  Payload: 16 + 12*k + 4*max_shard bytes per rank (k = 50, 2250-image shards: 9.6 KB) -- latency bound, so
// there is nothing to gain from fusing the collective into the scoring kernel; what matters is that the whole tail of
// a pass is three stream-ordered operations and one device->host copy instead of a chain of host round trips.
//
// NCCL is bound at run time (dlopen): the library the framework already loaded (torch bundles libnccl.so.2) is reused,
// and libalscore.so itself loads on machines without NCCL -- als_comm_* then fail with a message.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "ctx.h"
#include "select.cuh"

using als::DeviceGuard;
using als::fail;

namespace {

// ---- the slice of nccl.h this file uses (NCCL 2.x ABI) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
enum { ncclUint8 = 1 };    // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("ALS_NCCL_LIB");
    // a copy the process already holds (the framework's) wins, so that there is one NCCL per process
    n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!n.handle && env && *env) n.handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) n.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) {
      const char* e = dlerror();
      n.why = std::string("libnccl.so.2 not found (set ALS_NCCL_LIB to its path): ") + (e ? e : "");
      return;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(n.handle, name);
      if (!p) { ok = false; n.why = std::string("NCCL symbol missing: ") + name; }
      return p;
    };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) n.handle = nullptr;
  });
  return n;
}

#define ALS_NCCL(ctx, call)                                                                          \
  do {                                                                                               \
    ncclResult_t _r = (call);                                                                        \
    if (_r != 0) return fail(ctx, ALS_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(_r)); \
  } while (0)

int need_nccl(als_ctx* ctx) {
  if (!nccl().handle) return fail(ctx, ALS_ERR_CUDA, "NCCL is not available: %s", nccl().why.c_str());
  return ALS_OK;
}

// Exchange record of one rank: {int64 lo, int64 n} | f32 key[k] | i64 id[k] | f32 score[width]
struct RecordLayout {
  size_t keys_off, ids_off, scores_off, bytes;
};
RecordLayout record_layout(int64_t k, int64_t width) {
  RecordLayout r;
  r.keys_off = 16;
  r.ids_off = (r.keys_off + static_cast<size_t>(k) * 4 + 7) & ~static_cast<size_t>(7);
  r.scores_off = (r.ids_off + static_cast<size_t>(k) * 8 + 15) & ~static_cast<size_t>(15);
  r.bytes = (r.scores_off + static_cast<size_t>(width) * 4 + 15) & ~static_cast<size_t>(15);
  return r;
}

struct GlobalArgs {
  const int64_t* unlabelled;
  int64_t M, k, width;
  RecordLayout rec;
  als::SelectBlock blk;
};

// step 1: candidates of the ids this rank owns + its score slice -> send record (one launch on ctx->stream)
int global_local_phase(als_ctx* ctx, const GlobalArgs& a, int64_t lo, int64_t hi) {
  DeviceGuard g(ctx->device);
  void* p = ctx->xchg_send;
  ALS_TRY(als::grow_bytes(ctx, &p, &ctx->xchg_send_cap, a.rec.bytes));
  ctx->xchg_send = static_cast<unsigned char*>(p);
  p = ctx->xchg_recv;
  ALS_TRY(als::grow_bytes(ctx, &p, &ctx->xchg_recv_cap, a.rec.bytes * static_cast<size_t>(ctx->comm_world)));
  ctx->xchg_recv = static_cast<unsigned char*>(p);
  ALS_TRY(als::ensure_select_block(ctx, a.blk, a.k));
  ALS_TRY(als::upload_unlabelled(ctx, a.unlabelled, a.M));
  als::SelectSrc src{};
  src.mode = 1;
  src.ids = ctx->sel_ids;
  src.pool = ctx->pool32;
  src.lo = lo;
  src.hi = hi;
  als::SelectOut out{};
  out.count = reinterpret_cast<long long*>(ctx->sel_out);  // scratch here; the merge launch rewrites it
  out.keys = reinterpret_cast<float*>(ctx->xchg_send + a.rec.keys_off);
  out.ids = reinterpret_cast<long long*>(ctx->xchg_send + a.rec.ids_off);
  out.pad_base = als::kPadId + static_cast<long long>(ctx->comm_rank) * (a.k > 0 ? a.k : 1);
  als::ExportDesc ex{};
  ex.rec = ctx->xchg_send;
  ex.scores_off = static_cast<long long>(a.rec.scores_off);
  ex.lo = lo;
  ex.n = hi - lo;
  ex.width = a.width;
  ex.pool = ctx->pool32;
  int nl = 0;
  ALS_CUDA(ctx, als::launch_select(src, a.M, a.k, als::ScatterDesc{}, ex, out, ctx->sel_tmp_keys, ctx->sel_tmp_ids, ctx->stream, &nl));
  ctx->launches += nl;
  return ALS_OK;
}

// step 2: the one collective of a pool pass
int global_gather_phase(als_ctx* ctx, const GlobalArgs& a) {
  DeviceGuard g(ctx->device);
  ALS_NCCL(ctx, nccl().AllGather(ctx->xchg_send, ctx->xchg_recv, a.rec.bytes, ncclUint8, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
  return ALS_OK;
}

// step 3: other ranks' scores -> pool vector, merge world*k candidates, gather unlabelled_confidence (one launch)
int global_merge_phase(als_ctx* ctx, const GlobalArgs& a) {
  DeviceGuard g(ctx->device);
  als::SelectSrc src{};
  src.mode = 2;
  src.rec = ctx->xchg_recv;
  src.rec_stride = static_cast<long long>(a.rec.bytes);
  src.keys_off = static_cast<long long>(a.rec.keys_off);
  src.ids_off = static_cast<long long>(a.rec.ids_off);
  src.kc = a.k;
  als::ScatterDesc sc{};
  sc.rec = ctx->xchg_recv;
  sc.rec_stride = static_cast<long long>(a.rec.bytes);
  sc.scores_off = static_cast<long long>(a.rec.scores_off);
  sc.world = ctx->comm_world;
  sc.self = ctx->comm_rank;
  sc.pool = ctx->pool32;
  sc.pool_n = ctx->pool_n;
  const als::SelectOut out = als::select_out_of(ctx, a.blk, a.M);
  int nl = 0;
  ALS_CUDA(ctx, als::launch_select(src, a.k * ctx->comm_world, a.k, sc, als::ExportDesc{}, out, ctx->sel_tmp_keys, ctx->sel_tmp_ids,
                                   ctx->stream, &nl));
  ctx->launches += nl;
  return ALS_OK;
}

int check_global_args(als_ctx* ctx, const int64_t* unlabelled, int64_t M, int64_t lo, int64_t hi, int64_t* max_shard) {
  if (ctx->pool_n < 0) return fail(ctx, ALS_ERR_STATE, "als_pool_begin has not been called");
  if (M < 0) return fail(ctx, ALS_ERR_INVALID, "M must be >= 0");
  if (M > 0 && !unlabelled) return fail(ctx, ALS_ERR_INVALID, "unlabelled is NULL");
  if (lo < 0 || hi < lo || hi > ctx->pool_n)
    return fail(ctx, ALS_ERR_INVALID, "shard [%lld, %lld) is not inside [0, %lld)", (long long)lo, (long long)hi, (long long)ctx->pool_n);
  if (*max_shard <= 0) *max_shard = (ctx->pool_n + ctx->comm_world - 1) / ctx->comm_world;
  if (hi - lo > *max_shard)
    return fail(ctx, ALS_ERR_INVALID, "this rank owns %lld examples but max_shard is %lld", (long long)(hi - lo), (long long)*max_shard);
  return ALS_OK;
}

}  // namespace

extern "C" {

int als_comm_unique_id(void* out128) {
  if (!out128) return fail(nullptr, ALS_ERR_INVALID, "out128 is NULL");
  ALS_TRY(need_nccl(nullptr));
  ncclUniqueId id;
  ALS_NCCL(nullptr, nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return ALS_OK;
}

int als_comm_destroy(als_ctx* ctx) {
  if (!ctx || !ctx->comm) return ALS_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (nccl().handle) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  return ALS_OK;
}

int als_comm_init_rank(als_ctx* ctx, int rank, int world, const void* unique_id128) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, ALS_ERR_INVALID, "bad rank %d of %d", rank, world);
  if (!unique_id128) return fail(ctx, ALS_ERR_INVALID, "unique_id128 is NULL");
  ALS_TRY(need_nccl(ctx));
  als_comm_destroy(ctx);
  DeviceGuard g(ctx->device);
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ALS_NCCL(ctx, nccl().CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return ALS_OK;
}

int als_comm_init_all(als_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return fail(nullptr, ALS_ERR_INVALID, "need at least one context");
  for (int i = 0; i < n; ++i)
    if (!ctxs[i]) return fail(nullptr, ALS_ERR_INVALID, "ctxs[%d] is NULL", i);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (ctxs[i]->device == ctxs[j]->device)
        return fail(ctxs[0], ALS_ERR_INVALID, "ctxs[%d] and ctxs[%d] are bound to the same GPU %d", j, i, ctxs[i]->device);
  ALS_TRY(need_nccl(ctxs[0]));
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    als_comm_destroy(ctxs[i]);
    devs[i] = ctxs[i]->device;
  }
  std::vector<ncclComm_t> comms(n, nullptr);
  ALS_NCCL(ctxs[0], nccl().CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) {
    ctxs[i]->comm = comms[i];
    ctxs[i]->comm_rank = i;
    ctxs[i]->comm_world = n;
  }
  return ALS_OK;
}

int als_pool_select_global(als_ctx* ctx, const int64_t* unlabelled, int64_t M, int64_t selection_size, int64_t shard_lo,
                           int64_t shard_hi, int64_t max_shard, int64_t* out_ids, float* out_unlabelled_conf,
                           int64_t* out_count) {
  if (!ctx) return fail(nullptr, ALS_ERR_INVALID, "context is NULL");
  if (!ctx->comm) {  // world 1
    if (ctx->pool_n >= 0 && (shard_lo != 0 || shard_hi != ctx->pool_n))
      return fail(ctx, ALS_ERR_INVALID, "no communicator: the only rank must own the whole pool [0, %lld)", (long long)ctx->pool_n);
    return als_pool_select(ctx, unlabelled, M, selection_size, out_ids, out_unlabelled_conf, out_count);
  }
  if (!out_count) return fail(ctx, ALS_ERR_INVALID, "out_count is NULL");
  *out_count = 0;
  ALS_TRY(check_global_args(ctx, unlabelled, M, shard_lo, shard_hi, &max_shard));
  ALS_TRY(als::validate_unlabelled(ctx, unlabelled, M));
  const int64_t k = selection_size < 0 ? 0 : (selection_size < M ? selection_size : M);  // :707-708
  if (k > 0 && !out_ids) return fail(ctx, ALS_ERR_INVALID, "out_ids is NULL");
  GlobalArgs a{unlabelled, M, k, max_shard, record_layout(k, max_shard), als::select_block(k, M)};
  ALS_TRY(global_local_phase(ctx, a, shard_lo, shard_hi));
  ALS_TRY(global_gather_phase(ctx, a));
  ALS_TRY(global_merge_phase(ctx, a));
  DeviceGuard g(ctx->device);
  return als::fetch_select_block(ctx, a.blk, k, M, out_ids, out_unlabelled_conf, out_count);
}

int als_pool_select_global_all(als_ctx** ctxs, int n, const int64_t* unlabelled, int64_t M, int64_t selection_size,
                               const int64_t* shard_lo, const int64_t* shard_hi, int64_t max_shard, int64_t* out_ids,
                               float* out_unlabelled_conf, int64_t* out_count) {
  if (!ctxs || n < 1 || !ctxs[0]) return fail(nullptr, ALS_ERR_INVALID, "need at least one context");
  if (!shard_lo || !shard_hi) return fail(ctxs[0], ALS_ERR_INVALID, "shard bounds are NULL");
  if (!out_count) return fail(ctxs[0], ALS_ERR_INVALID, "out_count is NULL");
  *out_count = 0;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return fail(ctxs[0], ALS_ERR_INVALID, "ctxs[%d] is NULL", i);
    if (!ctxs[i]->comm || ctxs[i]->comm_world != n || ctxs[i]->comm_rank != i)
      return fail(ctxs[0], ALS_ERR_STATE, "ctxs[%d] is not rank %d of an als_comm_init_all communicator of %d", i, i, n);
    if (ctxs[i]->pool_n != ctxs[0]->pool_n) return fail(ctxs[0], ALS_ERR_STATE, "the contexts hold pools of different sizes");
  }
  int64_t width = max_shard;
  for (int i = 0; i < n; ++i) {
    int64_t w = max_shard;
    ALS_TRY(check_global_args(ctxs[i], unlabelled, M, shard_lo[i], shard_hi[i], &w));
    width = w;
  }
  ALS_TRY(als::validate_unlabelled(ctxs[0], unlabelled, M));
  const int64_t k = selection_size < 0 ? 0 : (selection_size < M ? selection_size : M);
  if (k > 0 && !out_ids) return fail(ctxs[0], ALS_ERR_INVALID, "out_ids is NULL");
  GlobalArgs a{unlabelled, M, k, width, record_layout(k, width), als::select_block(k, M)};
  for (int i = 0; i < n; ++i) ALS_TRY(global_local_phase(ctxs[i], a, shard_lo[i], shard_hi[i]));
  ALS_NCCL(ctxs[0], nccl().GroupStart());  // one thread drives every rank: the collectives must be issued as a group
  int rc = ALS_OK;
  for (int i = 0; i < n && rc == ALS_OK; ++i) rc = global_gather_phase(ctxs[i], a);
  ALS_NCCL(ctxs[0], nccl().GroupEnd());
  ALS_TRY(rc);
  for (int i = 0; i < n; ++i) ALS_TRY(global_merge_phase(ctxs[i], a));
  for (int i = 1; i < n; ++i) {
    DeviceGuard g(ctxs[i]->device);
    ALS_CUDA(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
  }
  DeviceGuard g(ctxs[0]->device);
  return als::fetch_select_block(ctxs[0], a.blk, k, M, out_ids, out_unlabelled_conf, out_count);
}

}  // extern "C"
