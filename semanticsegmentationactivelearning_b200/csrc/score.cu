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

#include "common.cuh"
#include "pixel_math.cuh"

namespace als {

constexpr int kSmemHeader = 128 + 256;  // mbarriers (2 x 8 x 8 B) + per-stage TileMeta (8 x 32 B)
static_assert(kSmemHeader % 128 == 0 && kMaxStages * 32 <= 256, "stage buffers must stay 128-byte aligned");

// ---- compile-time launch policy -------------------------------------------------------
__host__ __device__ constexpr int lpp_for(int C) { return C <= 36 ? 1 : C <= 72 ? 2 : C <= 144 ? 4 : 8; }
__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int ppt_for(int C, int es, bool multi) {
  const int lpp = lpp_for(C);
  const int cl = cdiv(C, lpp);
  const int row = (kConsumerThreads / lpp) * C * es;  // bytes per pixel-slot row
  int ppt = 1;
  const int reg_cap = multi ? 20 : 40;
  while (ppt < 4 && 2 * ppt * cl <= reg_cap && 2 * ppt * row <= 32 * 1024) ppt *= 2;
  return ppt;
}
__host__ __device__ constexpr int pow2_divisor(int v, int cap) {
  int p = 1;
  while (p < cap && v % (2 * p) == 0) p *= 2;
  return p;
}
__host__ __device__ constexpr int gcd_i(int a, int b) { return b == 0 ? a : gcd_i(b, a % b); }

template <typename E, int C, bool MULTI>
struct Cfg {
  static constexpr int ES = sizeof(E);
  static constexpr int LPP = lpp_for(C);
  static constexpr int CL = cdiv(C, LPP);
  static constexpr bool EXACT = (LPP * CL == C);
  static constexpr int G = kConsumerThreads / LPP;  // pixels per slot row
  static constexpr int PPT = ppt_for(C, ES, MULTI);
  static constexpr int TILE_PIX = G * PPT;
  static constexpr int STAGE_BYTES = ((TILE_PIX * C * ES + 127) / 128) * 128;
  // widest shared-memory access every lane's run start is aligned to
  static constexpr int VB = EXACT ? pow2_divisor(gcd_i(CL * ES, C * ES), 16) : ES;
  // resident CTAs per SM the kernel is compiled for: 3 when the per-thread class registers are few
  // (more warps hide the dependent MUFU/FMA chains), else 2
  static constexpr int MINB = (PPT * CL <= 24) ? 3 : 2;
  static_assert((LPP - 1) * CL < C, "every lane must own at least one class");
};

// ---- shared memory -> registers -------------------------------------------------------
template <typename E, int CL, int VB>
__device__ __forceinline__ void load_run(const unsigned char* __restrict__ src, float (&x)[CL]) {
  if constexpr (sizeof(E) == 4) {
    if constexpr (VB == 16) {
#pragma unroll
      for (int i = 0; i < CL / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(src + 16 * i);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
    } else if constexpr (VB == 8) {
#pragma unroll
      for (int i = 0; i < CL / 2; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(src + 8 * i);
        x[2 * i] = v.x; x[2 * i + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CL; ++j) x[j] = *reinterpret_cast<const float*>(src + 4 * j);
    }
  } else {  // bf16: value = bits << 16
    if constexpr (VB == 16) {
#pragma unroll
      for (int i = 0; i < CL / 8; ++i) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + 16 * i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          x[8 * i + 2 * q] = __uint_as_float(w[q] << 16);
          x[8 * i + 2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
        }
      }
    } else if constexpr (VB == 8) {
#pragma unroll
      for (int i = 0; i < CL / 4; ++i) {
        const uint2 v = *reinterpret_cast<const uint2*>(src + 8 * i);
        x[4 * i] = __uint_as_float(v.x << 16); x[4 * i + 1] = __uint_as_float(v.x & 0xffff0000u);
        x[4 * i + 2] = __uint_as_float(v.y << 16); x[4 * i + 3] = __uint_as_float(v.y & 0xffff0000u);
      }
    } else if constexpr (VB == 4) {
#pragma unroll
      for (int i = 0; i < CL / 2; ++i) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + 4 * i);
        x[2 * i] = __uint_as_float(w << 16); x[2 * i + 1] = __uint_as_float(w & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CL; ++j)
        x[j] = __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(src + 2 * j)) << 16);
    }
  }
}

// Lanes whose class run is only partly inside [0, C) (C not a multiple of LPP).
template <typename E, int CL>
__device__ __forceinline__ void load_run_partial(const unsigned char* __restrict__ src, float (&x)[CL], int nvalid) {
#pragma unroll
  for (int j = 0; j < CL; ++j) {
    float v = -INFINITY;
    if (j < nvalid) {
      if constexpr (sizeof(E) == 4) v = *reinterpret_cast<const float*>(src + 4 * j);
      else v = __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(src + 2 * j)) << 16);
    }
    x[j] = v;
  }
}

// ---- the tiled kernel ------------------------------------------------------------------------
// Per-stage tile descriptor the producer publishes next to the data (sample 0 of a tile only).
struct TileMeta {
  long long pix0;  // first global pixel of the tile, -1 = no more work
  long long img;   // image of that pixel
  long long off;   // its offset inside the image
  int npix;        // pixels in the tile (< TILE_PIX only for the last tile of the pool)
  int in_img;      // how many of them belong to `img` (the rest start the next image(s))
};

template <typename E, int C, int MEASURE>
__global__ void __launch_bounds__(kBlockThreads, Cfg<E, C, (MEASURE == kMulti)>::MINB) score_tiles_kernel(const ScoreParams p) {
  constexpr bool MULTI = (MEASURE == kMulti);
  using K = Cfg<E, C, MULTI>;
  constexpr int CL = K::CL, LPP = K::LPP, PPT = K::PPT, G = K::G, ES = K::ES;

  extern __shared__ __align__(128) unsigned char smem[];
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
    // ===== producer: one lane claims tiles and feeds the ring with 1-D bulk copies =====
    // Tiles are claimed dynamically (first one static, the rest from a global counter) so SMs that
    // see less HBM bandwidth simply take fewer tiles; the integer per-image sums make the result
    // independent of who scored what.
    if (lane == 0) {
      const uint64_t policy = l2_policy_evict_first();
      const E* base = static_cast<const E*>(p.logits);
      int s = 0;
      uint32_t ph = 0;
      long long tile = blockIdx.x;
      while (true) {
        const bool live = tile < p.num_tiles;
        // claim the next tile now; its latency hides behind this tile's copies
        const long long next = live ? static_cast<long long>(gridDim.x) +
                                          static_cast<long long>(atomicAdd(p.tile_counter, 1ull))
                                    : tile;
        if (!live) {
          mbar_wait(&empty[s], ph ^ 1u);
          meta[s].pix0 = -1;
          mbar_arrive_expect_tx(&full[s], 0);
          break;
        }
        const long long pix0 = tile * K::TILE_PIX;
        const long long rem = p.total_pixels - pix0;
        const uint32_t npix = rem < K::TILE_PIX ? static_cast<uint32_t>(rem) : K::TILE_PIX;
        const uint32_t bytes = npix * C * ES;
        const uint32_t bulk = bytes & ~15u;
        const long long img = pix0 / p.P;
        for (int t = 0; t < p.T; ++t) {
          mbar_wait(&empty[s], ph ^ 1u);
          unsigned char* dst = stage_base + static_cast<size_t>(s) * K::STAGE_BYTES;
          const unsigned char* src = reinterpret_cast<const unsigned char*>(base + t * p.sample_stride + pix0 * C);
          if (t == 0) {
            const long long off = pix0 - img * p.P;
            const long long left = p.P - off;  // pixels of image `img` from the tile start on
            meta[s].pix0 = pix0;
            meta[s].img = img;
            meta[s].off = off;
            meta[s].npix = static_cast<int>(npix);
            meta[s].in_img = left < npix ? static_cast<int>(left) : static_cast<int>(npix);
          }
          for (uint32_t b = bulk; b < bytes; ++b) dst[b] = src[b];  // < 16 trailing bytes of the whole pool
          mbar_arrive_expect_tx(&full[s], bulk);                    // release: publishes meta + tail bytes
          if (bulk) bulk_g2s(dst, src, bulk, &full[s], policy);
          if (++s == nstage) { s = 0; ph ^= 1u; }
        }
        tile = next;
      }
    }
    return;
  }

  // ===== consumers =====
  const int tid = threadIdx.x;
  const int sub = tid & (LPP - 1);
  const int pl = tid / LPP;
  const int class0 = sub * CL;
  const int nvalid = K::EXACT ? CL : min(CL, C - class0);
  const unsigned int run_off = (pl * C + class0) * ES;

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
      const unsigned char* st = stage_base + static_cast<size_t>(s) * K::STAGE_BYTES + run_off;
      // Pixel slots past the end of the pool (last tile only) hold stale but in-bounds bytes:
      // they are computed like the rest (no divergence around the shuffles) and dropped at emit.
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if constexpr (K::EXACT) load_run<E, CL, K::VB>(st + k * (G * C * ES), x[k]);
        else load_run_partial<E, CL>(st + k * (G * C * ES), x[k], nvalid);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);  // values are in registers: hand the stage back
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
#pragma unroll
          for (int k = 0; k < PPT; ++k) acc.add(conf[k], p.fx_scale);
        }
      } else {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          const int l = k * G + pl;
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
        const unsigned char* st = stage_base + static_cast<size_t>(s) * K::STAGE_BYTES + run_off;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if constexpr (K::EXACT) load_run<E, CL, K::VB>(st + k * (G * C * ES), x[k]);
          else load_run_partial<E, CL>(st + k * (G * C * ES), x[k], nvalid);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
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
        const int l = k * G + pl;
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
}

// ---- generic fallback: any C, any alignment (direct global loads, one thread per pixel) -----
template <typename E>
__device__ __forceinline__ float ld_elem(const E* p) {
  if constexpr (sizeof(E) == 4) return *reinterpret_cast<const float*>(p);
  else return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(p)) << 16);
}

constexpr int kGenericThreads = 128;

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
          const float pj = ex2_approx((ld_elem(ps + c) - m1) * kLog2e) * r;
          float& m = mu_s[c * kGenericThreads + threadIdx.x];
          const float delta = pj - m;
          m = fmaf(delta, inv_t, m);
          m2s = fmaf(delta, pj - m, m2s);
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

// ---- finalize: fixed point -> f64 mean (:261-263), f32 scatter by example index (:700) -------
__global__ void finalize_kernel(long long* __restrict__ acc, long long acc_stride, unsigned int* __restrict__ flags,
                                unsigned long long* __restrict__ tile_counter, int n, double inv_scale_p,
                                double* __restrict__ scores64, float* __restrict__ pool32,
                                const long long* __restrict__ example_index, long long num_examples) {
  pdl_launch_dependents();
  pdl_wait();  // the scoring kernel's sums are complete and visible
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *tile_counter = 0ull;  // ready for the next scoring launch on this stream
  if (i >= n) return;
  long long total = 0;
#pragma unroll 8
  for (int r = 0; r < kAccReplicas; ++r) {
    total += acc[r * acc_stride + i];
    acc[r * acc_stride + i] = 0;
  }
  const double s = flags[i] ? __longlong_as_double(0x7ff8000000000000ll) : static_cast<double>(total) * inv_scale_p;
  flags[i] = 0;
  if (scores64) scores64[i] = s;
  if (pool32) {
    const long long e = example_index ? example_index[i] : i;
    if (e >= 0 && e < num_examples) pool32[e] = static_cast<float>(s);  // f64 -> f32 round-to-nearest
  }
}

cudaError_t launch_finalize(long long* acc, long long acc_stride, unsigned int* flags, unsigned long long* tile_counter,
                            int n, double inv_scale_p, double* scores64,
                            float* pool32, const long long* example_index, long long num_examples,
                            cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  void* args[] = {&acc, &acc_stride, &flags, &tile_counter, &n, &inv_scale_p, &scores64, &pool32, &example_index, &num_examples};
  return launch_pdl((const void*)finalize_kernel, dim3((n + 255) / 256), dim3(256), args, 0, stream);
}

// ---- dispatch ------------------------------------------------------------------------------------
// Class counts with a specialised tiled kernel; any other C runs the generic kernel.
#define ALS_C_LIST(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(16) X(19) X(20) X(21) \
  X(24) X(32) X(33) X(34) X(37) X(40) X(59) X(60) X(65) X(66) X(91) X(133) X(150) X(151) X(171) X(182)

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
    const int budget = per_cta - kSmemHeader;
    int stages = budget / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages >= 2) {
      plan.tiled = true;
      plan.stages = stages;
      plan.smem_bytes = kSmemHeader + stages * stage_bytes;
      plan.block = kBlockThreads;
      const long long tiles = (total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
      int per_sm = 0;
      if (cudaFuncSetAttribute(plan.func, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan.func, plan.block, plan.smem_bytes) != cudaSuccess ||
          per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
      }
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

cudaError_t launch_score(const LaunchPlan& plan, int dtype, ScoreParams p, cudaStream_t stream) {
  if (p.total_pixels <= 0) return cudaSuccess;
  p.any_out = (p.conf_map || p.label || p.mask) ? 1 : 0;
  cudaError_t err;
  if (plan.tiled) {
    p.stages = plan.stages;
    p.num_tiles = (p.total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
    void* args[] = {&p};
    return launch_pdl(plan.func, dim3(plan.grid), dim3(plan.block), args, plan.smem_bytes, stream);
  }
  const void* f = dtype == 0 ? (const void*)score_generic_kernel<float> : (const void*)score_generic_kernel<__nv_bfloat16>;
  if (plan.smem_bytes > 48 * 1024) {
    err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (err != cudaSuccess) return err;
  }
  void* args[] = {&p};
  return launch_pdl(f, dim3(plan.grid), dim3(plan.block), args, plan.smem_bytes, stream);
}

}  // namespace als
