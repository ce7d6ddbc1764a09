// Streamed Monte-Carlo-dropout accumulation for sm_100a (the in-tier half).
//
// The resident path (score.cu, MEASURE = kMulti) needs all T samples of a pixel in HBM at once: [T,N,H,W,C], 948 GB
// for BASELINE config 2.  A producer that runs one stochastic forward pass at a time (the reference only builds
// spatial_dropout under training=True, /root/reference/models/util/extra_ops.py:137-151 and
// /root/reference/models/enet/enet_modules.py:591-594) can instead hand over each sample as it is made:
//
//   als_mc_begin -> als_mc_add_sample x T -> als_mc_finish
//
// Per sample, one streaming pass reads the sample's logits through the same bulk-copy ring as score.cu, reads the
// pixel's Welford state (negated running mean per class + summed M2), folds the sample in with the SAME
// welford_update() the resident kernel runs in registers, and writes the state back.  f32 loads / stores are exact, so
// the final confidences and per-image scores are bit-identical to the resident path.
//
// The state is read and written with .cg accesses (L2 only): tiles are claimed dynamically, so the SM that reads a
// tile's state in launch t+1 is usually not the one that wrote it in launch t, and with programmatic dependent launch
// the next grid's CTAs become resident while the previous grid still runs -- an L1 line left from an earlier launch
// must never satisfy a state load (found by the full-geometry parity test: bit-identical on small pools, not on large).
//
// State layout (tiled kernels): one block of PPT * (CL + 1) * 256 floats per tile, indexed
// [pixel slot k][class j | M2][consumer thread], so every load / store of a warp is one coalesced 128-byte line.
// Algorithmic bytes per pixel and sample: C*sizeof(E) logits read + (C+1)*4 state read (not for sample 0)
// + (C+1)*4 state written; finish: (C+1)*4 read.  Roofline: HBM copy bandwidth (reads and writes).
#include "mc.cuh"

#include <stdlib.h>

#include "tiles.cuh"

namespace als {

template <typename E, int C>
struct McLayout {
  using K = Cfg<E, C, true>;
  static constexpr int ROW = K::CL + 1;  // floats of state per lane and pixel slot
  static constexpr long long TILE_FLOATS = static_cast<long long>(K::PPT) * ROW * kConsumerThreads;
};

// ---- one sample: logits tile (bulk-copy ring) + state tile (direct coalesced loads / stores) ----
template <typename E, int C, bool FIRST>
__global__ void __launch_bounds__(kBlockThreads, Cfg<E, C, true>::MINB) mc_update_kernel(const ScoreParams p,
                                                                                       float* __restrict__ state,
                                                                                       const int t) {
  using K = Cfg<E, C, true>;
  using L = McLayout<E, C>;
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
  pdl_launch_dependents();
  pdl_wait();  // the previous sample's state writes (and the tile counter reset) are complete

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == kConsumerThreads / 32) {
    if (lane == 0) produce_tiles<K, E, C>(p, full, empty, meta, stage_base, nstage);
    return;
  }

  const int tid = threadIdx.x;
  const int sub = tid & (LPP - 1);
  const int pl = tid / LPP;
  const int class0 = sub * CL;
  const int nvalid = K::EXACT ? CL : min(CL, C - class0);
  const float inv_t = __frcp_rn(static_cast<float>(t + 1));  // as in score_tiles_kernel<kMulti>

  int s = 0;
  uint32_t ph = 0;
  while (true) {
    mbar_wait(&full[s], ph);
    const long long tile_pix0 = meta[s].pix0;
    if (tile_pix0 < 0) break;
    const int npix = meta[s].npix;
    float* __restrict__ stp = state + (tile_pix0 / K::TILE_PIX) * L::TILE_FLOATS + tid;

    // logits first, and the stage goes back as soon as they are in registers; only then the state loads: queued in
    // front of the shared-memory reads they would hold those back for a global-memory latency
    float x[PPT][CL];
    load_tile_pixels<K, E, C>(stage_base + static_cast<size_t>(s) * K::STAGE_BYTES, pl, class0, nvalid, x);
    const uint32_t dep = loaded_dep<PPT, CL>(x) | (static_cast<uint32_t>(npix) >> 1);  // npix: same stage, read late
    warp_release_after_loads(&empty[s], dep, lane, p.never);
    if (++s == nstage) { s = 0; ph ^= 1u; }

    float nmu[PPT][CL];
    float m2s[PPT];
    if constexpr (!FIRST) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
#pragma unroll
        for (int j = 0; j < CL; ++j) nmu[k][j] = __ldcg(stp + (k * L::ROW + j) * kConsumerThreads);
        m2s[k] = __ldcg(stp + (k * L::ROW + CL) * kConsumerThreads);
      }
    } else {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        m2s[k] = 0.f;
#pragma unroll
        for (int j = 0; j < CL; ++j) nmu[k][j] = 0.f;
      }
    }

#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if constexpr (FIRST) {
        if (p.label) {  // pseudo_label = argmax of sample 0 (active_learning.py:234-236)
          const int lbl = group_argmax<CL, LPP>(x[k], nvalid, class0);
          const int l = K::slot(k, pl);
          if (sub == 0 && l < npix) p.label[tile_pix0 + l] = static_cast<uint8_t>(lbl);
        }
      }
      welford_update<CL, LPP, K::EXACT>(x[k], nvalid, inv_t, nmu[k], m2s[k]);
#pragma unroll
      for (int j = 0; j < CL; ++j) __stcg(stp + (k * L::ROW + j) * kConsumerThreads, nmu[k][j]);
      __stcg(stp + (k * L::ROW + CL) * kConsumerThreads, m2s[k]);
    }
  }
  // the last CTA re-arms the dynamic tile scheduler for the next launch on this stream (kernel -> kernel ordering only:
  // a stream-ordered memset between two programmatically dependent launches is not something to rely on)
  if (last_cta_ticket(p.done_counter, &s_last) && threadIdx.x == 0) {
    *p.tile_counter = 0ull;
    *p.done_counter = 0u;
  }
}

// ---- finish: state -> measure of the predictive mean / summed variance -> per-image fixed-point sums ----
template <typename E, int C>
__global__ void __launch_bounds__(kConsumerThreads, 3) mc_finish_kernel(const ScoreParams p, const float* __restrict__ state) {
  using K = Cfg<E, C, true>;
  using L = McLayout<E, C>;
  constexpr int CL = K::CL, LPP = K::LPP, PPT = K::PPT, G = K::G;
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int sub = tid & (LPP - 1);
  const int pl = tid / LPP;
  const int nvalid = K::EXACT ? CL : min(CL, C - sub * CL);
  ImageAcc acc;
  for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const long long pix0 = tile * K::TILE_PIX;
    const long long rem = p.total_pixels - pix0;
    const int npix = rem < K::TILE_PIX ? static_cast<int>(rem) : K::TILE_PIX;
    const long long img = pix0 / p.P;
    const long long off = pix0 - img * p.P;
    const long long left = p.P - off;
    const int in_img = left < npix ? static_cast<int>(left) : npix;
    const bool plain = (in_img == K::TILE_PIX) && !p.any_out;
    if (img != acc.img) {  // CTA-uniform
      acc.flush(p);
      acc.img = img;
    }
    const float* __restrict__ stp = state + tile * L::TILE_FLOATS + tid;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      float nmu[CL];
#pragma unroll
      for (int j = 0; j < CL; ++j) nmu[j] = __ldcg(stp + (k * L::ROW + j) * kConsumerThreads);
      const float m2s = __ldcg(stp + (k * L::ROW + CL) * kConsumerThreads);
      const float conf = conf_multi<CL, LPP, K::EXACT>(nmu, m2s, nvalid, p);
      const int l = K::slot(k, pl);
      if (plain) {
        if (sub == 0) acc.add(conf, p.fx_scale);
      } else if (sub == 0 && l < npix) {
        emit_pixel(p, acc, conf, 0, pix0, off, l, in_img);
      }
    }
  }
  acc.flush(p);
}

// ---- generic fallback: any C, any alignment; planar state [class | M2][pixel], positive mean ----
// mode 0: first sample, 1: later sample, 2: finish.  Same expressions as score_generic_kernel's T > 1 branch.
template <typename E>
__global__ void __launch_bounds__(kGenericThreads) mc_generic_kernel(const ScoreParams p, float* __restrict__ state,
                                                                     const int t, const int mode) {
  pdl_launch_dependents();
  pdl_wait();
  const E* base = static_cast<const E*>(p.logits);
  const int C = p.C;
  const long long total = p.total_pixels;
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
  for (long long g = static_cast<long long>(blockIdx.x) * kGenericThreads + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * kGenericThreads) {
    if (mode < 2) {
      const E* ps = base + g * C;
      float m1 = ld_elem(ps);
      for (int c = 1; c < C; ++c) m1 = fmaxf(m1, ld_elem(ps + c));
      if (mode == 0 && p.label) {  // pseudo_label = argmax of sample 0, first maximum wins (:234-236)
        float bv = ld_elem(ps);
        int lbl = 0;
        for (int c = 1; c < C; ++c) {
          const float v = ld_elem(ps + c);
          if (v > bv) { bv = v; lbl = c; }
        }
        p.label[g] = static_cast<uint8_t>(lbl);
      }
      float S = 0.f;
      for (int c = 0; c < C; ++c) S += ex2_approx((ld_elem(ps + c) - m1) * kLog2e);
      const float r = rcp_approx(S);
      const float inv_t = __frcp_rn(static_cast<float>(t + 1));
      float m2s = mode == 0 ? 0.f : state[static_cast<long long>(C) * total + g];
      for (int c = 0; c < C; ++c) {
        const float pj = __fmul_rn(ex2_approx((ld_elem(ps + c) - m1) * kLog2e), r);  // no contraction: streamed == resident
        float m = mode == 0 ? 0.f : state[c * total + g];
        const float delta = __fsub_rn(pj, m);
        m = fmaf(delta, inv_t, m);
        m2s = fmaf(delta, __fsub_rn(pj, m), m2s);
        state[c * total + g] = m;
      }
      state[static_cast<long long>(C) * total + g] = m2s;
      continue;
    }
    float conf;
    if (p.measure == kVariance) {
      conf = fmaf(-state[static_cast<long long>(C) * total + g], p.inv_T, 1.0f);
    } else if (p.measure == kEntropy) {
      float h = 0.f;
      for (int c = 0; c < C; ++c) {
        const float m = state[c * total + g];
        h = fmaf(-m, lg2_approx(m + kTiny), h);
      }
      conf = fmaf(-h, p.inv_log2_c, 1.0f);
    } else {
      float m1 = state[g], m2 = -INFINITY;
      for (int c = 1; c < C; ++c) {
        const float v = state[c * total + g];
        m2 = fmaxf(m2, fminf(m1, v));
        m1 = fmaxf(m1, v);
      }
      conf = p.measure == kMargin ? m1 - m2 : m1;
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
  }
  flush();
}

// ---- dispatch ------------------------------------------------------------------------------------
template <typename E, int C>
static void pick_mc(McPlan& plan) {
  using K = Cfg<E, C, true>;
  using L = McLayout<E, C>;
  plan.first = (const void*)mc_update_kernel<E, C, true>;
  plan.update = (const void*)mc_update_kernel<E, C, false>;
  plan.finish = (const void*)mc_finish_kernel<E, C>;
  plan.name = "mc_update_kernel";
  plan.tile_pixels = K::TILE_PIX;
  plan.smem_bytes = K::STAGE_BYTES;  // per stage for now
  plan.grid = K::MINB;               // CTAs per SM for now
  plan.state_floats = L::TILE_FLOATS;  // per tile for now
}

McPlan plan_mc(int dtype, int C, long long total_pixels, bool aligned, int num_sms, int max_smem_per_block) {
  McPlan plan{};
  bool ok = false;
  if (aligned) {
    switch (C) {
#define X(c)                                                                   \
  case c:                                                                      \
    if (dtype == 0) pick_mc<float, c>(plan); else pick_mc<__nv_bfloat16, c>(plan); \
    ok = true;                                                                 \
    break;
      ALS_C_LIST(X)
#undef X
      default: break;
    }
  }
  if (ok) {
    const int stage_bytes = plan.smem_bytes;
    const int ctas_per_sm = plan.grid;
    int per_cta = (228 * 1024) / ctas_per_sm - 1024;
    if (per_cta > max_smem_per_block) per_cta = max_smem_per_block;
    int stages = (per_cta - kSmemHeader - 128) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages >= 2) {
      plan.tiled = true;
      plan.stages = stages;
      plan.smem_bytes = kSmemHeader + stages * stage_bytes;
      plan.block = kBlockThreads;
      const long long tiles = (total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
      int per_sm = 0;
      for (const void* f : {plan.first, plan.update}) {
        const int n = resident_ctas(f, plan.block, plan.smem_bytes);
        per_sm = per_sm == 0 ? n : (n < per_sm ? n : per_sm);
      }
      const long long resident = static_cast<long long>(per_sm) * num_sms;
      plan.grid = static_cast<int>(tiles < resident ? (tiles > 0 ? tiles : 1) : resident);
      plan.state_floats *= (tiles > 0 ? tiles : 1);
      plan.finish_block = kConsumerThreads;
      const long long fin = 6ll * num_sms;
      plan.finish_grid = static_cast<int>(tiles < fin ? (tiles > 0 ? tiles : 1) : fin);
      return plan;
    }
  }
  plan = McPlan{};
  plan.tiled = false;
  plan.name = "mc_generic_kernel";
  plan.block = plan.finish_block = kGenericThreads;
  plan.tile_pixels = kGenericThreads;
  const long long blocks = (total_pixels + kGenericThreads - 1) / kGenericThreads;
  const long long cap = 16ll * num_sms;
  plan.grid = plan.finish_grid = static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  plan.state_floats = (static_cast<long long>(C) + 1) * (total_pixels > 0 ? total_pixels : 1);
  return plan;
}

cudaError_t launch_mc_update(const McPlan& plan, int dtype, ScoreParams p, float* state, int t, cudaStream_t stream) {
  if (p.total_pixels <= 0) return cudaSuccess;
  p.T = 1;
  p.sample_stride = 0;
  p.any_out = 0;
  p.never = 0xffffffffu;
  p.claim_shift = claim_shift_for(plan.grid);
  p.claim = 1;  // runs of tiles were measured on the streamed update too: 1.016 -> 1.007 (2) -> 0.996 (4) of the copy peak
  if (plan.tiled) {
    p.stages = plan.stages;
    p.num_tiles = (p.total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
    void* args[] = {&p, &state, &t};
    return launch_pdl(t == 0 ? plan.first : plan.update, dim3(plan.grid), dim3(plan.block), args, plan.smem_bytes, stream);
  }
  int mode = t == 0 ? 0 : 1;
  const void* f = dtype == 0 ? (const void*)mc_generic_kernel<float> : (const void*)mc_generic_kernel<__nv_bfloat16>;
  void* args[] = {&p, &state, &t, &mode};
  return launch_pdl(f, dim3(plan.grid), dim3(plan.block), args, 0, stream);
}

cudaError_t launch_mc_finish(const McPlan& plan, int dtype, ScoreParams p, const float* state, cudaStream_t stream) {
  if (p.total_pixels <= 0) return cudaSuccess;
  p.any_out = (p.conf_map || p.mask) ? 1 : 0;
  p.label = nullptr;
  if (plan.tiled) {
    p.num_tiles = (p.total_pixels + plan.tile_pixels - 1) / plan.tile_pixels;
    void* args[] = {&p, &state};
    return launch_pdl(plan.finish, dim3(plan.finish_grid), dim3(plan.finish_block), args, 0, stream);
  }
  int t = 0, mode = 2;
  float* st = const_cast<float*>(state);
  const void* f = dtype == 0 ? (const void*)mc_generic_kernel<float> : (const void*)mc_generic_kernel<__nv_bfloat16>;
  void* args[] = {&p, &st, &t, &mode};
  return launch_pdl(f, dim3(plan.finish_grid), dim3(plan.finish_block), args, 0, stream);
}

}  // namespace als
