// Fused pool-scoring kernels for sm_100a.
//
// One streaming pass over the logits computes, per pixel, softmax -> confidence measure
// (-> Welford mean / variance over T Monte-Carlo samples) and accumulates the per-image
// sum; no probability map ever reaches HBM.  Replaces the ~10 separate TensorFlow ops of
// /root/reference/active_learning.py:239-263 (softmax :239, entropy :240-251, margin
// :252-255, max-prob :256-258, f64 mean :261-263) and, optionally, the per-pixel consumers
// pseudo_label :234-236 and pseudo_mask :265-269.
//
// Data movement: the pool is treated as one flat stream of N*P pixels of C contiguous
// elements.  A tile is a contiguous byte range, so a single 1-D bulk async copy
// (cp.async.bulk, the TMA engine; SASS UBLKCP) per (tile, sample) lands it in shared memory
// and signals an mbarrier with the byte count.  One producer warp keeps `stages` tiles in
// flight per CTA; 8 consumer warps read their pixel's classes out of shared memory (stride
// = classes per lane, bank-conflict free when that is odd or vectorisable), release the
// stage as soon as the values are in registers, and do the math in registers.
#include "score.cuh"

#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "pixel_math.cuh"
#include "tiles.cuh"

namespace als {

// ---- finalize: fixed point -> f64 mean (:261-263), f32 scatter by example index (:700) -------
__device__ __forceinline__ void finalize_image(int i, long long* __restrict__ acc, long long acc_stride,
                                               unsigned int* __restrict__ flags, double inv_scale_p,
                                               double* __restrict__ scores64, float* __restrict__ pool32,
                                               const long long* __restrict__ example_index, long long index_base,
                                               long long num_examples) {
  long long total = 0;
#pragma unroll 8
  for (int r = 0; r < kAccReplicas; ++r) {
    total += __ldcg(acc + r * acc_stride + i);  // L2: the sums were made by REDs of other SMs
    acc[r * acc_stride + i] = 0;
  }
  const double s = __ldcg(flags + i) ? __longlong_as_double(0x7ff8000000000000ll) : static_cast<double>(total) * inv_scale_p;
  flags[i] = 0;
  if (scores64) scores64[i] = s;
  if (pool32) {
    const long long e = example_index ? example_index[i] : index_base + i;
    if (e >= 0 && e < num_examples) pool32[e] = static_cast<float>(s);  // f64 -> f32 round-to-nearest
  }
}

__global__ void finalize_kernel(long long* __restrict__ acc, long long acc_stride, unsigned int* __restrict__ flags,
                                unsigned long long* __restrict__ tile_counter, int n, double inv_scale_p,
                                double* __restrict__ scores64, float* __restrict__ pool32,
                                const long long* __restrict__ example_index, long long index_base, long long num_examples) {
  pdl_launch_dependents();
  pdl_wait();  // the scoring kernel's sums are complete and visible
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *tile_counter = 0ull;  // ready for the next scoring launch on this stream
  if (i >= n) return;
  finalize_image(i, acc, acc_stride, flags, inv_scale_p, scores64, pool32, example_index, index_base, num_examples);
}

// Fused finalize: every CTA counts itself done once all its consumer warps have flushed their sums; the CTA that
// finds itself last (classic "last block" reduction: fence, atomic ticket, fence) does finalize_kernel's work for the
// whole launch.  One launch less per batch: ~3 us of a 50 us batch-of-8 call, and one dependency edge less in front of
// the selection.  Consumer threads only (the producer warp has returned): named barrier 1.
__device__ __forceinline__ void fused_finalize(const ScoreParams& p, int* s_last) {
  if (!last_cta_ticket(p.done_counter, s_last)) return;
  for (int i = threadIdx.x; i < p.fin_n; i += kConsumerThreads)
    finalize_image(i, p.acc, p.acc_stride, p.flags, p.fin_inv_scale_p, p.fin_scores64, p.fin_pool32, p.fin_example_index,
                   p.fin_index_base, p.fin_num_examples);
  if (threadIdx.x == 0) {  // ready for the next scoring launch on this stream
    *p.tile_counter = 0ull;
    *p.done_counter = 0u;
  }
}

template <typename E, int C, int MEASURE>
__global__ void __launch_bounds__(kBlockThreads, Cfg<E, C, (MEASURE == kMulti)>::MINB) score_tiles_kernel(const ScoreParams p) {
  constexpr bool MULTI = (MEASURE == kMulti);
  using K = Cfg<E, C, MULTI>;
  constexpr int CL = K::CL, LPP = K::LPP, PPT = K::PPT, G = K::G, ES = K::ES;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ int s_last;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  TileMeta* meta = reinterpret_cast<TileMeta*>(smem + 128);
  unsigned char* stage_base = smem + kSmemHeader;
  const int nstage = p.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerThreads / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // everything above overlapped the tail of the previous kernel in the stream (programmatic dependent launch);
  // from here on the logits, the accumulators and the tile counter are read
  pdl_launch_dependents();
  pdl_wait();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kConsumerThreads / 32) {
    // ===== producer: one lane claims tiles and feeds the ring with 1-D bulk copies (tiles.cuh) =====
    if (lane == 0) produce_tiles<K, E, C>(p, full, empty, meta, stage_base, nstage);
    return;
  }

  // ===== consumers =====
  const int tid = threadIdx.x;
  const int sub = tid & (LPP - 1);
  const int pl = tid / LPP;
  const int class0 = sub * CL;
  const int nvalid = K::EXACT ? CL : min(CL, C - class0);

  ImageAcc acc;
  int s = 0;
  uint32_t ph = 0;
  while (true) {
    mbar_wait(&full[s], ph);  // sample 0 of the next tile (or the end marker)
    const long long tile_pix0 = meta[s].pix0;
    if (tile_pix0 < 0) break;
    const long long img = meta[s].img;
    const long long off = meta[s].off;
    const int npix = meta[s].npix;
    const int in_img = meta[s].in_img;
    // common case: a full tile inside one image and no per-pixel outputs -> plain per-thread sums
    const bool plain = (in_img == K::TILE_PIX) && !p.any_out;
    if (img != acc.img) {  // CTA-uniform
      acc.flush(p);
      acc.img = img;
    }

    if constexpr (!MULTI) {
      float x[PPT][CL];
      // Pixel slots past the end of the pool (last tile only) hold stale but in-bounds bytes:
      // they are computed like the rest (no divergence around the shuffles) and dropped at emit.
      load_tile_pixels<K, E, C>(stage_base + static_cast<size_t>(s) * K::STAGE_BYTES, pl, class0, nvalid, x);
      // (the tile descriptor sits in the same stage: the fields only the rare emit path reads join the dependency)
      const uint32_t dep = loaded_dep<PPT, CL>(x) | (static_cast<uint32_t>(off) >> 1) | (static_cast<uint32_t>(npix) >> 1);
      warp_release_after_loads(&empty[s], dep, lane, p.never);  // values are in registers: hand the stage back
      if (++s == nstage) { s = 0; ph ^= 1u; }
      float conf[PPT];
      bool bad = false;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        conf[k] = conf_single<CL, LPP, K::EXACT, MEASURE>(x[k], nvalid, p);
        bad |= !(conf[k] == conf[k]);
      }
      if constexpr (MEASURE == kEntropy) {
        if (__any_sync(0xffffffffu, bad)) {  // rare: some pixel of this warp has a -inf / NaN logit
#pragma unroll
          for (int k = 0; k < PPT; ++k) conf[k] = conf_single<CL, LPP, K::EXACT, MEASURE, true>(x[k], nvalid, p);
        }
      }
      if (plain) {
        if (sub == 0) {
          if constexpr (ES == 2) {
            acc.add_q22(conf);  // bf16: XU-bound kernel, conversion moved off the XU pipe
          } else {
#pragma unroll
            for (int k = 0; k < PPT; ++k) acc.add(conf[k], p.fx_scale);
          }
        }
      } else if (in_img == K::TILE_PIX) {  // per-pixel outputs, full tile inside one image
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          int lbl = 0;
          if (p.label) {
            if constexpr (LPP == 1) lbl = argmax_first<CL>(x[k]);
            else lbl = group_argmax<CL, LPP>(x[k], nvalid, class0);
          }
          if (sub == 0) emit_pixel_full(p, acc, conf[k], lbl, tile_pix0 + K::slot(k, pl));
        }
      } else {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          const int l = K::slot(k, pl);
          int lbl = 0;
          if (p.label) lbl = group_argmax<CL, LPP>(x[k], nvalid, class0);
          if (sub == 0 && l < npix) emit_pixel(p, acc, conf[k], lbl, tile_pix0, off, l, in_img);
        }
      }
    } else {
      float mu[PPT][CL];
      float m2s[PPT];
      int lbl[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        m2s[k] = 0.f;
        lbl[k] = 0;
#pragma unroll
        for (int j = 0; j < CL; ++j) mu[k][j] = 0.f;
      }
      for (int t = 0; t < p.T; ++t) {
        float x[PPT][CL];
        if (t > 0) mbar_wait(&full[s], ph);
        load_tile_pixels<K, E, C>(stage_base + static_cast<size_t>(s) * K::STAGE_BYTES, pl, class0, nvalid, x);
        const uint32_t dep = loaded_dep<PPT, CL>(x) | (static_cast<uint32_t>(off) >> 1) | (static_cast<uint32_t>(npix) >> 1);
        warp_release_after_loads(&empty[s], dep, lane, p.never);
        if (++s == nstage) { s = 0; ph ^= 1u; }
        const float inv_t = __frcp_rn(static_cast<float>(t + 1));
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if (t == 0 && p.label) lbl[k] = group_argmax<CL, LPP>(x[k], nvalid, class0);
          welford_update<CL, LPP, K::EXACT>(x[k], nvalid, inv_t, mu[k], m2s[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int l = K::slot(k, pl);
        const float conf = conf_multi<CL, LPP, K::EXACT>(mu[k], m2s[k], nvalid, p);
        if (plain) {
          if (sub == 0) acc.add(conf, p.fx_scale);
        } else if (sub == 0 && l < npix) {
          emit_pixel(p, acc, conf, lbl[k], tile_pix0, off, l, in_img);
        }
      }
    }
  }
  acc.flush(p);
  if (p.fin_n > 0) fused_finalize(p, &s_last);
}

// ---- generic fallback: any C, any alignment (direct global loads, one thread per pixel) -----

template <typename E>
__global__ void __launch_bounds__(kGenericThreads) score_generic_kernel(const ScoreParams p) {
  extern __shared__ float mu_s[];  // [C][kGenericThreads] when T > 1
  pdl_launch_dependents();
  pdl_wait();
  const E* base = static_cast<const E*>(p.logits);
  const int C = p.C;
  long long acc_sum = 0, acc_img = -1;
  unsigned int acc_nan = 0;
  auto flush = [&]() {
    if (acc_img >= 0) {
      if (acc_sum) atomicAdd(reinterpret_cast<unsigned long long*>(p.acc + acc_slot(p) + acc_img), static_cast<unsigned long long>(acc_sum));
      if (acc_nan) atomicOr(p.flags + acc_img, 1u);
    }
    acc_sum = 0;
    acc_nan = 0;
  };
  for (long long g = static_cast<long long>(blockIdx.x) * kGenericThreads + threadIdx.x; g < p.total_pixels;
       g += static_cast<long long>(gridDim.x) * kGenericThreads) {
    const E* px = base + g * C;
    float conf;
    int lbl = 0;
    {
      float bv = ld_elem(px);
      for (int c = 1; c < C; ++c) {
        const float v = ld_elem(px + c);
        if (v > bv) { bv = v; lbl = c; }
      }
    }
    if (p.T == 1) {
      float m1 = ld_elem(px), m2 = -INFINITY;
      for (int c = 1; c < C; ++c) {
        const float v = ld_elem(px + c);
        m2 = fmaxf(m2, fminf(m1, v));
        m1 = fmaxf(m1, v);
      }
      float S = 0.f, A = 0.f;
      for (int c = 0; c < C; ++c) {
        const float d = max_nan(ld_elem(px + c) - m1, -FLT_MAX);
        const float e = ex2_approx(d * kLog2e);
        S += e;
        A = fmaf(e, d, A);
      }
      const float r = rcp_approx(S);
      if (p.measure == kEntropy) conf = fmaf(-fmaf(lg2_approx(S), kLn2, -A * r), p.inv_log2_c * kLog2e, 1.0f);
      else if (p.measure == kMargin) conf = (1.0f - ex2_approx((m2 - m1) * kLog2e)) * r;
      else conf = r;
    } else {
      float m2s = 0.f;
      for (int c = 0; c < C; ++c) mu_s[c * kGenericThreads + threadIdx.x] = 0.f;
      for (int t = 0; t < p.T; ++t) {
        const E* ps = px + t * p.sample_stride;
        float m1 = ld_elem(ps);
        for (int c = 1; c < C; ++c) m1 = fmaxf(m1, ld_elem(ps + c));
        float S = 0.f;
        for (int c = 0; c < C; ++c) S += ex2_approx((ld_elem(ps + c) - m1) * kLog2e);
        const float r = rcp_approx(S);
        const float inv_t = __frcp_rn(static_cast<float>(t + 1));
        for (int c = 0; c < C; ++c) {
          const float pj = __fmul_rn(ex2_approx((ld_elem(ps + c) - m1) * kLog2e), r);  // no contraction: streamed == resident
          float& m = mu_s[c * kGenericThreads + threadIdx.x];
          const float delta = __fsub_rn(pj, m);
          m = fmaf(delta, inv_t, m);
          m2s = fmaf(delta, __fsub_rn(pj, m), m2s);
        }
      }
      if (p.measure == kVariance) {
        conf = fmaf(-m2s, p.inv_T, 1.0f);
      } else if (p.measure == kEntropy) {
        float h = 0.f;
        for (int c = 0; c < C; ++c) {
          const float m = mu_s[c * kGenericThreads + threadIdx.x];
          h = fmaf(-m, lg2_approx(m + kTiny), h);
        }
        conf = fmaf(-h, p.inv_log2_c, 1.0f);
      } else {
        float m1 = mu_s[threadIdx.x], m2 = -INFINITY;
        for (int c = 1; c < C; ++c) {
          const float v = mu_s[c * kGenericThreads + threadIdx.x];
          m2 = fmaxf(m2, fminf(m1, v));
          m1 = fmaxf(m1, v);
        }
        conf = p.measure == kMargin ? m1 - m2 : m1;
      }
    }
    const long long img = g / p.P;
    if (img != acc_img) {
      flush();
      acc_img = img;
    }
    const bool isnan_ = !(conf == conf);
    acc_sum += isnan_ ? 0ll : __float2ll_rn(conf * p.fx_scale);
    acc_nan |= isnan_ ? 1u : 0u;
    if (p.conf_map) p.conf_map[g] = conf;
    if (p.mask) p.mask[g] = (conf < p.threshold) ? 0 : 1;
    if (p.label) p.label[g] = static_cast<uint8_t>(lbl);
  }
  flush();
}

cudaError_t launch_finalize(long long* acc, long long acc_stride, unsigned int* flags, unsigned long long* tile_counter,
                            int n, double inv_scale_p, double* scores64,
                            float* pool32, const long long* example_index, long long index_base, long long num_examples,
                            cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  void* args[] = {&acc, &acc_stride, &flags, &tile_counter, &n, &inv_scale_p, &scores64, &pool32, &example_index, &index_base, &num_examples};
  return launch_pdl((const void*)finalize_kernel, dim3((n + 255) / 256), dim3(256), args, 0, stream);
}

// ---- dispatch ------------------------------------------------------------------------------------
template <typename E, int C>
static bool pick(int measure, int T, LaunchPlan& plan) {
  const bool multi = T > 1;
  const void* f = nullptr;
  const char* name = nullptr;
  if (multi) { f = (const void*)score_tiles_kernel<E, C, kMulti>; name = "score_tiles_kernel<multi>"; }
  else if (measure == kEntropy) { f = (const void*)score_tiles_kernel<E, C, kEntropy>; name = "score_tiles_kernel<entropy>"; }
  else if (measure == kMargin) { f = (const void*)score_tiles_kernel<E, C, kMargin>; name = "score_tiles_kernel<margin>"; }
  else if (measure == kConfidence) { f = (const void*)score_tiles_kernel<E, C, kConfidence>; name = "score_tiles_kernel<confidence>"; }
  else return false;
  plan.func = f;
  plan.name = name;
  if (multi) {
    using K = Cfg<E, C, true>;
    plan.tile_pixels = K::TILE_PIX; plan.lanes_per_pixel = K::LPP; plan.pixels_per_thread = K::PPT;
    plan.smem_bytes = K::STAGE_BYTES;  // per stage for now
    plan.ctas_per_sm = K::MINB;
  } else {
    using K = Cfg<E, C, false>;
    plan.tile_pixels = K::TILE_PIX; plan.lanes_per_pixel = K::LPP; plan.pixels_per_thread = K::PPT;
    plan.smem_bytes = K::STAGE_BYTES;
    plan.ctas_per_sm = K::MINB;
  }
  return true;
}

// Resident CTAs per SM of a tiled kernel at its shared-memory size; the opt-in attribute and the occupancy query are
// driver calls (~10 us together), so they are made once per (kernel, size, device) and remembered -- a batch-of-8 call is
// only ~50 us of GPU time and a single-chunk pool pass pays every host microsecond in front of its one launch.
int resident_ctas(const void* func, int block, int smem_bytes) {
  struct Entry { const void* func; int smem, dev, per_sm; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (const Entry& e : cache)
    if (e.func == func && e.smem == smem_bytes && e.dev == dev) return e.per_sm;
  int per_sm = 0;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, block, smem_bytes) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 1;
  }
  cache.push_back({func, smem_bytes, dev, per_sm});
  return per_sm;
}

LaunchPlan plan_score(int dtype, int C, int measure, int T, long long total_pixels, bool aligned, int num_sms,
                      int max_smem_per_block) {
  LaunchPlan plan{};
  bool ok = false;
  if (aligned) {
    switch (C) {
#define X(c)                                                                            \
  case c:                                                                               \
    ok = (dtype == 0) ? pick<float, c>(measure, T, plan) : pick<__nv_bfloat16, c>(measure, T, plan); \
    break;
      ALS_C_LIST(X)
#undef X
      default: break;
    }
  }
  if (ok) {
    const int stage_bytes = plan.smem_bytes;
    // shared memory per CTA so that `ctas_per_sm` CTAs fit in the SM's 228 KB (1 KB reserved per CTA)
    int per_cta = (228 * 1024) / plan.ctas_per_sm - 1024;
    if (const char* env = getenv("ALS_SMEM_KB")) per_cta = atoi(env) * 1024;  // tuning knob (bench only)
    if (per_cta > max_smem_per_block) per_cta = max_smem_per_block;
    const int budget = per_cta - kSmemHeader - 128;  // 128: static shared memory (fused-finalize flag)
    int stages = budget / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages >= 2) {
      plan.tiled = true;
      plan.stages = stages;
      plan.smem_bytes = kSmemHeader + stages * stage_bytes;
      plan.block = kBlockThreads;
      const long long tiles = (total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
      const int per_sm = resident_ctas(plan.func, plan.block, plan.smem_bytes);
      const long long resident = static_cast<long long>(per_sm) * num_sms;  // persistent: one wave
      plan.grid = static_cast<int>(tiles < resident ? (tiles > 0 ? tiles : 1) : resident);
      return plan;
    }
  }
  plan = LaunchPlan{};
  plan.tiled = false;
  plan.func = nullptr;
  plan.name = "score_generic_kernel";
  plan.block = kGenericThreads;
  plan.smem_bytes = T > 1 ? C * kGenericThreads * 4 : 0;
  plan.stages = 0;
  plan.tile_pixels = kGenericThreads;
  plan.lanes_per_pixel = 1;
  plan.pixels_per_thread = 1;
  const long long blocks = (total_pixels + kGenericThreads - 1) / kGenericThreads;
  const long long cap = 16ll * num_sms;
  plan.grid = static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  return plan;
}

int max_claim(int T, int dtype) {
  static const int env = [] {
    const char* e = getenv("ALS_CLAIM");  // bring-up knob
    return e ? atoi(e) : 0;
  }();
  if (T != 1) return 1;
  if (env > 0) return env;
  return dtype == 0 ? 2 : 16;
}

int claim_shift_for(int grid) {
  static const int mult = [] {
    const char* e = getenv("ALS_CLAIM_TAPER");  // bring-up knob: runs shrink below this many tiles per CTA and run length
    return e && atoi(e) > 0 ? atoi(e) : 2;  // measured 1 / 2 / 4: train8 0.942 / 0.954 / 0.942 of the copy peak, the rest equal
  }();
  int s = 2;
  while ((1ll << s) < static_cast<long long>(mult) * grid) ++s;
  return s;
}

cudaError_t launch_score(const LaunchPlan& plan, int dtype, ScoreParams p, cudaStream_t stream) {
  if (p.total_pixels <= 0) return cudaSuccess;
  p.any_out = (p.conf_map || p.label || p.mask) ? 1 : 0;
  p.claim = max_claim(p.T, dtype);
  p.never = 0xffffffffu;
  p.claim_shift = claim_shift_for(plan.grid);
  cudaError_t err;
  if (plan.tiled) {
    p.stages = plan.stages;
    p.num_tiles = (p.total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
    void* args[] = {&p};
    return launch_pdl(plan.func, dim3(plan.grid), dim3(plan.block), args, plan.smem_bytes, stream);
  }
  if (p.fin_n > 0) return cudaErrorInvalidValue;  // only the tiled kernel folds the finalize in (callers check plan.tiled)
  const void* f = dtype == 0 ? (const void*)score_generic_kernel<float> : (const void*)score_generic_kernel<__nv_bfloat16>;
  if (plan.smem_bytes > 48 * 1024) {
    err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (err != cudaSuccess) return err;
  }
  void* args[] = {&p};
  return launch_pdl(f, dim3(plan.grid), dim3(plan.block), args, plan.smem_bytes, stream);
}

}  // namespace als
