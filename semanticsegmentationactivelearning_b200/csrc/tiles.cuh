// Launch policy, shared-memory staging and tile descriptors shared by the tiled scoring kernels
// (score.cu: resident [T,N,H,W,C] logits; mc.cu: one Monte-Carlo sample at a time with the Welford state in HBM).
#pragma once
#include <cuda_bf16.h>
#include <float.h>

#include "common.cuh"
#include "pixel_math.cuh"
#include "score.cuh"

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
#ifndef ALS_MULTI_REGCAP   // bring-up knob: register budget (class values per thread) of the T > 1 kernels
#define ALS_MULTI_REGCAP 20
#endif
  const int reg_cap = multi ? ALS_MULTI_REGCAP : 40;
#ifndef ALS_PPT_MAX   // bring-up knob: cap on the pixels per thread
#define ALS_PPT_MAX 4
#endif
  while (ppt < ALS_PPT_MAX && 2 * ppt * cl <= reg_cap && 2 * ppt * row <= 32 * 1024) ppt *= 2;
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
#ifndef ALS_MINB_SMALL  // bring-up knob: resident CTAs per SM the small-footprint kernels are compiled for
#define ALS_MINB_SMALL 3
#endif
  static constexpr int MINB = (PPT * CL <= 24) ? ALS_MINB_SMALL : 2;
  static_assert((LPP - 1) * CL < C, "every lane must own at least one class");
  // bf16 logits with an odd class count: a pixel is C*2 bytes, so only every other pixel starts on a 4-byte boundary
  // and a thread that owns ONE pixel per slot row is left with 2-byte loads (C LDS.U16 per pixel, half of them bank
  // conflicted: ncu counted 47 % extra shared-memory wavefronts on cfg3 bf16).  PAIR: a thread owns two ADJACENT pixels
  // (2*C bf16 = C aligned 32-bit words, lane stride C words: odd, conflict free) -- half the load instructions.
  static constexpr bool PAIR = (ES == 2) && (LPP == 1) && (C % 2 == 1) && (PPT % 2 == 0);
  // pixel slot inside the tile of the k-th pixel of consumer thread `pl`
  __host__ __device__ static constexpr int slot(int k, int pl) {
    return PAIR ? ((k >> 1) * 2 * G + 2 * pl + (k & 1)) : (k * G + pl);
  }
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

// "Last CTA" ticket for the consumer threads of a tiled kernel (the producer warp has returned): every CTA counts itself
// done once all its consumer warps have passed this point; returns true in the CTA that finds itself last -- all other
// CTAs' global writes and atomics are then visible to it (fence, atomic ticket, fence).  Named barrier 1.
__device__ __forceinline__ bool last_cta_ticket(unsigned int* done_counter, int* s_flag) {
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
  if (threadIdx.x == 0) {
    __threadfence();
    *s_flag = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
  if (!*s_flag) return false;
  __threadfence();
  return true;
}

// All PPT pixels of consumer thread (pl, class run class0) of one staged tile -> registers.
template <typename K, typename E, int C>
__device__ __forceinline__ void load_tile_pixels(const unsigned char* __restrict__ stage, int pl, int class0, int nvalid,
                                                 float (&x)[K::PPT][K::CL]) {
  constexpr int CL = K::CL, G = K::G, ES = K::ES, PPT = K::PPT;
  if constexpr (K::PAIR) {
#pragma unroll
    for (int q = 0; q < PPT / 2; ++q) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(stage + (q * 2 * G + 2 * pl) * (C * ES));
      uint32_t w[C];  // 2*C bf16 = C words: pixel 2q is elements 0..C-1, pixel 2q+1 elements C..2C-1
#pragma unroll
      for (int i = 0; i < C; ++i) w[i] = src[i];
#pragma unroll
      for (int j = 0; j < C; ++j) {
        x[2 * q][j] = __uint_as_float((j & 1) ? (w[j >> 1] & 0xffff0000u) : (w[j >> 1] << 16));
        const int e = C + j;
        x[2 * q + 1][j] = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
      }
    }
  } else {
    const unsigned char* st = stage + (pl * C + class0) * ES;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if constexpr (K::EXACT) load_run<E, CL, K::VB>(st + k * (G * C * ES), x[k]);
      else load_run_partial<E, CL>(st + k * (G * C * ES), x[k], nvalid);
    }
  }
}

// A word that depends on every register load_tile_pixels() filled (see mbar_arrive_after_loads): the per-pixel class
// maximum -- the first thing every measure computes anyway, so the chain is shared with it -- OR-ed over the thread's
// pixels, shifted right so that it can never be all ones (= ScoreParams::never, see warp_release_after_loads).
template <int PPT, int CL>
__device__ __forceinline__ uint32_t loaded_dep(const float (&x)[PPT][CL]) {
  uint32_t d = 0;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    float m = x[k][0];
#pragma unroll
    for (int j = 1; j < CL; ++j) m = fmaxf(m, x[k][j]);
    d |= __float_as_uint(m);
  }
  return d >> 1;
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

// ---- generic fallback kernels (any C, any alignment): direct global loads, one thread per pixel ----
template <typename E>
__device__ __forceinline__ float ld_elem(const E* p) {
  if constexpr (sizeof(E) == 4) return *reinterpret_cast<const float*>(p);
  else return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(p)) << 16);
}

constexpr int kGenericThreads = 128;

// ===== producer: one lane claims tiles and feeds the ring with 1-D bulk copies =====
// Tiles are claimed dynamically (first one static, the rest from a global counter) so SMs that
// see less HBM bandwidth simply take fewer tiles; the integer per-image sums make the result
// independent of who scored what.  Every (tile, sample) is one cp.async.bulk into the next ring stage.
template <typename K, typename E, int C>
__device__ __forceinline__ void produce_tiles(const ScoreParams& p, uint64_t* full, uint64_t* empty, TileMeta* meta,
                                              unsigned char* stage_base, const int nstage) {
  constexpr int ES = sizeof(E);
  const uint64_t policy = l2_policy_evict_first();
  const E* base = static_cast<const E*>(p.logits);
  int s = 0;
  uint32_t ph = 0;
  long long tile = blockIdx.x;
  long long run_end = tile + 1;  // the tiles [tile, run_end) are this CTA's current run
  long long next = 0, next_end = 0;
  bool claimed = false;
  long long pix0 = 0, img = 0, off = 0;  // of `tile`: divided out at the start of a run, stepped inside it
  bool stepped = false;
  while (true) {
    const bool live = tile < p.num_tiles;
    // Guided self-scheduling.  At the FIRST tile of a run the run after it is claimed, so the atomic's round trip
    // (~1 us under load, about the time a CTA has per 20 KB tile at 7.5 TB/s) hides behind the whole run; one atomic
    // per tile, consumed one tile later, held T = 1 launches ~5 % (f32) to ~8 % (bf16) under what they reach with
    // runs.  Runs are up to p.claim tiles while there is plenty left (more than two runs per CTA) and shrink to single
    // tiles towards the end, so the tail stays short.
    if (live && !claimed) {
      long long n = (p.num_tiles - run_end) >> p.claim_shift;  // tiles left / (2 * grid, rounded up to a power of two)
      n = n < 1 ? 1 : (n > p.claim ? p.claim : n);
      next = static_cast<long long>(gridDim.x) +
             static_cast<long long>(atomicAdd(p.tile_counter, static_cast<unsigned long long>(n)));
      next_end = next + n;
      claimed = true;
    }
    if (!live) {
      mbar_wait(&empty[s], ph ^ 1u);
      meta[s].pix0 = -1;
      mbar_arrive_expect_tx(&full[s], 0);
      break;
    }
    // This one thread's dependent instruction chain per tile competes with the consumer warps for issue slots (the
    // bf16 kernels are XU- and issue-bound): the 64-bit division is paid once per run, not once per tile.
    if (!stepped) {
      pix0 = tile * K::TILE_PIX;
      img = pix0 / p.P;
      off = pix0 - img * p.P;
    }
    const long long rem = p.total_pixels - pix0;
    const uint32_t npix = rem < K::TILE_PIX ? static_cast<uint32_t>(rem) : K::TILE_PIX;
    const uint32_t bytes = npix * C * ES;
    const uint32_t bulk = bytes & ~15u;
    for (int t = 0; t < p.T; ++t) {
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* dst = stage_base + static_cast<size_t>(s) * K::STAGE_BYTES;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(base + t * p.sample_stride + pix0 * C);
      if (t == 0) {
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
    if (tile + 1 < run_end) {
      ++tile;
      pix0 += K::TILE_PIX;
      off += K::TILE_PIX;
      while (off >= p.P) {
        off -= p.P;
        ++img;
      }
      stepped = true;
    } else {
      tile = next;
      run_end = next_end;
      claimed = false;
      stepped = false;
    }
  }
}

}  // namespace als

// Class counts with a specialised tiled kernel; any other C runs the generic kernel.
#define ALS_C_LIST(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(16) X(19) X(20) X(21) \
  X(24) X(32) X(33) X(34) X(37) X(40) X(59) X(60) X(65) X(66) X(91) X(133) X(150) X(151) X(171) X(182)
