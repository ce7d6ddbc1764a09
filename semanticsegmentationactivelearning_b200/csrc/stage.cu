// Host -> device staging of PAGEABLE memory at PCIe speed.
//
// The reference hands its logits over as fresh NumPy arrays (`sess.run`, /root/reference/active_learning.py:697-698):
// pageable memory.  cudaMemcpyAsync from pageable memory is staged by the driver through ONE thread's worth of copying
// (measured 11.3 GB/s on the GPU box against 55.3 GB/s from pinned memory, profiles/r02_host_staging.txt), which would
// make the drop-in 5x slower end to end than the link allows.  HostStager does what the driver does, in parallel: a few
// worker threads copy the source piece by piece into their own pinned bounce buffers (two each, so one fills while the
// other drains) and queue one cudaMemcpyAsync per piece on the context's copy stream.  Pinned sources bypass it.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "ctx.h"

namespace als {

class HostStager {
 public:
  static constexpr size_t kPiece = 4u << 20;  // bytes per piece: large enough for full-speed DMA, small enough to pipeline

  HostStager(int device, cudaStream_t stream, int threads) : device_(device), stream_(stream) {
    workers_.reserve(threads);
    for (int w = 0; w < threads; ++w) workers_.emplace_back([this, w] { run(w); });
  }

  ~HostStager() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      quit_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
  }

  int threads() const { return static_cast<int>(workers_.size()); }

  // Copy `bytes` from pageable `src` to device `dst`; returns when the source may be reused (every byte is either on
  // the device or in a bounce buffer with its DMA queued on the stream).  cudaSuccess or the first error a worker saw.
  cudaError_t copy(void* dst, const void* src, size_t bytes) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      dst_ = static_cast<unsigned char*>(dst);
      src_ = static_cast<const unsigned char*>(src);
      bytes_ = bytes;
      pieces_ = (bytes + kPiece - 1) / kPiece;
      next_.store(0);
      finished_ = 0;
      error_ = cudaSuccess;
      ++generation_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lock(mu_);
    done_cv_.wait(lock, [this] { return finished_ == static_cast<int>(workers_.size()); });
    return error_;
  }

 private:
  void run(int w) {
    cudaSetDevice(device_);
    unsigned char* slot[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = true;
    for (int b = 0; b < 2 && ok; ++b)
      ok = cudaMallocHost(reinterpret_cast<void**>(&slot[b]), kPiece) == cudaSuccess &&
           cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) == cudaSuccess;
    unsigned long long seen = 0;
    int use = 0;
    while (true) {
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
        if (quit_) break;
      }
      cudaError_t err = ok ? cudaSuccess : cudaErrorMemoryAllocation;
      while (err == cudaSuccess) {
        const size_t i = next_.fetch_add(1);
        if (i >= pieces_) break;
        const size_t off = i * kPiece;
        const size_t n = bytes_ - off < kPiece ? bytes_ - off : kPiece;
        err = cudaEventSynchronize(ev[use]);  // the DMA that last read this bounce buffer has finished
        if (err != cudaSuccess) break;
        memcpy(slot[use], src_ + off, n);
        err = cudaMemcpyAsync(dst_ + off, slot[use], n, cudaMemcpyHostToDevice, stream_);
        if (err == cudaSuccess) err = cudaEventRecord(ev[use], stream_);
        use ^= 1;
      }
      {
        std::lock_guard<std::mutex> lock(mu_);
        if (err != cudaSuccess && error_ == cudaSuccess) error_ = err;
        ++finished_;
      }
      done_cv_.notify_one();
    }
    for (int b = 0; b < 2; ++b) {
      if (ev[b]) {
        cudaEventSynchronize(ev[b]);
        cudaEventDestroy(ev[b]);
      }
      if (slot[b]) cudaFreeHost(slot[b]);
    }
  }

  int device_;
  cudaStream_t stream_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  unsigned long long generation_ = 0;
  bool quit_ = false;
  unsigned char* dst_ = nullptr;
  const unsigned char* src_ = nullptr;
  size_t bytes_ = 0, pieces_ = 0;
  std::atomic<size_t> next_{0};
  int finished_ = 0;
  cudaError_t error_ = cudaSuccess;
};

// ---- the context's side -----------------------------------------------------------------------------
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int stage_copy(als_ctx* ctx, void* dst, const void* src, size_t bytes) {
  // small copies and pinned sources: the plain asynchronous copy is already the fastest way
  if (bytes < 2 * HostStager::kPiece || is_pinned(src)) {
    ALS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    return ALS_OK;
  }
  if (!ctx->stager) {
    int threads = 8;
    if (const char* env = getenv("ALS_STAGE_THREADS")) threads = atoi(env);
    const unsigned hc = std::thread::hardware_concurrency();
    if (hc && threads > static_cast<int>(hc)) threads = static_cast<int>(hc);
    if (threads < 1) {  // ALS_STAGE_THREADS=0: leave pageable memory to the driver
      ALS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
      return ALS_OK;
    }
    ctx->stager = new HostStager(ctx->device, ctx->copy_stream, threads);
  }
  ALS_CUDA(ctx, ctx->stager->copy(dst, src, bytes));
  return ALS_OK;
}

void stage_destroy(als_ctx* ctx) {
  delete ctx->stager;
  ctx->stager = nullptr;
}

}  // namespace als
