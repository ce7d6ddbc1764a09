// Per-pixel confidence math and per-image accumulation shared by the scoring kernels (score.cu reads
// logits from HBM, head.cu takes them from tensor memory after the fused classifier head).
// Reference: /root/reference/active_learning.py:234-263.
#pragma once
#include <float.h>

#include "common.cuh"
#include "score.cuh"

namespace als {

enum : int { kEntropy = 0, kMargin = 1, kConfidence = 2, kVariance = 3, kMulti = 4 };

// Bring-up knob: the first ALS_POLY_PAIRS class pairs of a pixel take their exp2 from ex2_poly2 (FMA pipe) instead of
// MUFU.EX2 (XU pipe).  0 in the shipped library -- see DESIGN.md for what the split measured.
#ifndef ALS_POLY_PAIRS
#define ALS_POLY_PAIRS 0
#endif
__device__ __forceinline__ f32x2 ex2_pair(f32x2 t2, int pair_index) {
  if (pair_index < ALS_POLY_PAIRS) return ex2_poly2(t2);
  float t0, t1;
  unpack2(t2, t0, t1);
  return pack2(ex2_approx(t0), ex2_approx(t1));
}

// ---- per-pixel math ---------------------------------------------------------------------
// top-2 merge across the LPP lanes of a pixel
template <int LPP>
__device__ __forceinline__ void group_top2(float& m1, float& m2) {
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) {
    const float o1 = __shfl_xor_sync(0xffffffffu, m1, o);
    const float o2 = __shfl_xor_sync(0xffffffffu, m2, o);
    m2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
    m1 = fmaxf(m1, o1);
  }
}

// pseudo_label = argmax_c logits, first maximum wins (active_learning.py:234-236)
template <int CL, int LPP>
__device__ __forceinline__ int group_argmax(const float (&x)[CL], int nvalid, int class0) {
  float bv = x[0];
  int bi = class0;
#pragma unroll
  for (int j = 1; j < CL; ++j)
    if (j < nvalid && x[j] > bv) { bv = x[j]; bi = class0 + j; }
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  return bi;
}

// Same result for a pixel held by ONE thread (LPP == 1), two instructions per class instead of three: the maximum by
// the same fmaxf chain conf_single() runs (the compiler shares it), then the FIRST class that equals it.
template <int CL>
__device__ __forceinline__ int argmax_first(const float (&x)[CL]) {
  float m = x[0];
#pragma unroll
  for (int j = 1; j < CL; ++j) m = fmaxf(m, x[j]);
  int bi = 0;
#pragma unroll
  for (int j = CL - 1; j >= 0; --j) bi = (x[j] == m) ? j : bi;
  return (x[0] != x[0]) ? 0 : bi;  // a NaN in class 0 wins like in group_argmax (and np.argmax); the score is NaN anyway
}

// T == 1: confidence of one pixel straight from its logits.
//   softmax (:239): e = exp(x - max), p = e / S       -- never materialised
//   entropy (:243-251): -sum p log p = log S - (sum e*(x-max)) / S   (one log per pixel, not C)
//   margin  (:254-255): p(1) - p(2) = (1 - exp(x(2) - max)) / S
//   max-prob (:258):    1 / S
// Entropy works in log2 units with t = fma(x, log2e, -max*log2e): the rounding of max*log2e
// shifts every t by the same epsilon, which cancels between log2 S and (sum e*t)/S.  A -inf
// logit gives e*t = 0*(-inf) = NaN; instead of clamping every class, a NaN result re-runs the
// pixel's warp with the clamp (CLAMP = true), which also tells a real NaN from that artefact.
template <int CL, int LPP, bool EXACT, bool CLAMP>
__device__ __forceinline__ float entropy_conf(const float (&x)[CL], int nvalid, float m1, const ScoreParams& p) {
  const float ml = m1 * kLog2e;
  float S = 0.f, A = 0.f;
  if constexpr (EXACT && !CLAMP) {
    // two classes per issue slot: t = x*log2e - ml (FFMA2), S += e (FADD2), A += e*t (FFMA2)
    const f32x2 l2 = pack2(kLog2e, kLog2e), nml2 = pack2(-ml, -ml);
    f32x2 S2 = pack2(0.f, 0.f), A2 = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j + 1 < CL; j += 2) {
      const f32x2 t2 = fma2(pack2(x[j], x[j + 1]), l2, nml2);
      const f32x2 e2 = ex2_pair(t2, j / 2);
      S2 = add2(S2, e2);
      A2 = fma2(e2, t2, A2);
    }
    S = hsum2(S2);
    A = hsum2(A2);
    if constexpr (CL & 1) {
      const float t = fmaf(x[CL - 1], kLog2e, -ml);
      const float e = ex2_approx(t);
      S += e;
      A = fmaf(e, t, A);
    }
  } else {
#pragma unroll
    for (int j = 0; j < CL; ++j) {
      if (EXACT || j < nvalid) {
        float t = fmaf(x[j], kLog2e, -ml);
        if constexpr (CLAMP) t = max_nan(t, -FLT_MAX);  // -inf logits: p = 0 and 0*log(tiny) = 0 (:243)
        const float e = ex2_approx(t);
        S += e;
        A = fmaf(e, t, A);
      }
    }
  }
  if constexpr (LPP > 1) {
    S = group_sum<LPP>(S);
    A = group_sum<LPP>(A);
  }
  const float h2 = fmaf(-A, rcp_approx(S), lg2_approx(S));  // entropy in bits
  return fmaf(-h2, p.inv_log2_c, 1.0f);                      // 1 - H / log(C)
}

// sum_c exp2(x_c*log2e - m1*log2e), two classes per issue slot
template <int CL, bool EXACT>
__device__ __forceinline__ float exp_sum(const float (&x)[CL], int nvalid, float m1) {
  const float ml = m1 * kLog2e;
  float S = 0.f;
  if constexpr (EXACT) {
    const f32x2 l2 = pack2(kLog2e, kLog2e), nml2 = pack2(-ml, -ml);
    f32x2 S2 = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j + 1 < CL; j += 2) {
      S2 = add2(S2, ex2_pair(fma2(pack2(x[j], x[j + 1]), l2, nml2), j / 2));
    }
    S = hsum2(S2);
    if constexpr (CL & 1) S += ex2_approx(fmaf(x[CL - 1], kLog2e, -ml));
  } else {
#pragma unroll
    for (int j = 0; j < CL; ++j)
      if (j < nvalid) S += ex2_approx(fmaf(x[j], kLog2e, -ml));
  }
  return S;
}

// CLAMP (entropy only): the slower variant that survives -inf logits; the caller re-runs a whole warp
// with it when the fast variant produced a NaN (once per tile, not per pixel).
template <int CL, int LPP, bool EXACT, int MEASURE, bool CLAMP = false>
__device__ __forceinline__ float conf_single(const float (&x)[CL], int nvalid, const ScoreParams& p) {
  float m1 = x[0], m2 = -INFINITY;
#pragma unroll
  for (int j = 1; j < CL; ++j) {
    if (EXACT || j < nvalid) {
      if constexpr (MEASURE == kMargin) m2 = fmaxf(m2, fminf(m1, x[j]));
      m1 = fmaxf(m1, x[j]);
    }
  }
  if constexpr (LPP > 1) {
    if constexpr (MEASURE == kMargin) group_top2<LPP>(m1, m2);
    else m1 = group_max<LPP>(m1);
  }
  if constexpr (MEASURE == kEntropy) {
    if constexpr (CLAMP) return entropy_conf<CL, LPP, EXACT, true>(x, nvalid, m1, p);
    else return entropy_conf<CL, LPP, EXACT, false>(x, nvalid, m1, p);
  } else {
    float S = exp_sum<CL, EXACT>(x, nvalid, m1);
    if constexpr (LPP > 1) S = group_sum<LPP>(S);
    // the rounding of m1*log2e shifts every exponent by the same epsilon: take the top terms through the
    // same expression so it cancels in e / S
    const float ml = m1 * kLog2e;
    const float r = rcp_approx(S);
    const float e1 = ex2_approx(fmaf(m1, kLog2e, -ml));
    if constexpr (MEASURE == kMargin) return (e1 - ex2_approx(fmaf(m2, kLog2e, -ml))) * r;
    else return e1 * r;
  }
}

// T > 1: fold one sample's softmax into the running per-class mean and the summed M2.
// Welford with delta = p - mu_old:  mu += delta / n,  M2 += delta * (p - mu_new) = delta^2 * (1 - 1/n),
// so the per-class work is three FFMAs; p = e / S is formed inside the first one.  The exponent
// argument is one FFMA as well: the shared rounding error of max*log2e cancels in e / S.
template <int CL, int LPP, bool EXACT>
__device__ __forceinline__ void welford_update(float (&x)[CL], int nvalid, float inv_t, float (&nmu)[CL], float& m2s) {
  // nmu holds the NEGATED running mean so that delta = e*r - mu is a single (packed) FMA
  float m1 = x[0];
#pragma unroll
  for (int j = 1; j < CL; ++j)
    if (EXACT || j < nvalid) m1 = fmaxf(m1, x[j]);
  if constexpr (LPP > 1) m1 = group_max<LPP>(m1);
  const float ml = m1 * kLog2e;
  float S = 0.f, q = 0.f;
  if constexpr (EXACT) {
    const f32x2 l2 = pack2(kLog2e, kLog2e), nml2 = pack2(-ml, -ml);
    f32x2 e2[CL / 2];
    f32x2 S2 = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j + 1 < CL; j += 2) {
      float t0, t1;
      unpack2(fma2(pack2(x[j], x[j + 1]), l2, nml2), t0, t1);
      e2[j / 2] = pack2(ex2_approx(t0), ex2_approx(t1));
      S2 = add2(S2, e2[j / 2]);
    }
    S = hsum2(S2);
    float elast = 0.f;
    if constexpr (CL & 1) {
      elast = ex2_approx(fmaf(x[CL - 1], kLog2e, -ml));
      S += elast;
    }
    if constexpr (LPP > 1) S = group_sum<LPP>(S);
    const float r = rcp_approx(S);
    const f32x2 r2 = pack2(r, r), nit2 = pack2(-inv_t, -inv_t);
    f32x2 q2 = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j + 1 < CL; j += 2) {
      f32x2 n2 = pack2(nmu[j], nmu[j + 1]);
      const f32x2 d2 = fma2(e2[j / 2], r2, n2);  // delta = p - mu
      n2 = fma2(d2, nit2, n2);                  // -mu -= delta / n
      q2 = fma2(d2, d2, q2);
      unpack2(n2, nmu[j], nmu[j + 1]);
    }
    q = hsum2(q2);
    if constexpr (CL & 1) {
      const float delta = fmaf(elast, r, nmu[CL - 1]);
      nmu[CL - 1] = fmaf(delta, -inv_t, nmu[CL - 1]);
      q = fmaf(delta, delta, q);
    }
  } else {
#pragma unroll
    for (int j = 0; j < CL; ++j) {
      if (j < nvalid) {
        x[j] = ex2_approx(fmaf(x[j], kLog2e, -ml));
        S += x[j];
      }
    }
    if constexpr (LPP > 1) S = group_sum<LPP>(S);
    const float r = rcp_approx(S);
#pragma unroll
    for (int j = 0; j < CL; ++j) {
      if (j < nvalid) {
        const float delta = fmaf(x[j], r, nmu[j]);
        nmu[j] = fmaf(delta, -inv_t, nmu[j]);
        q = fmaf(delta, delta, q);
      }
    }
  }
  m2s = fmaf(q, 1.0f - inv_t, m2s);
}

// T > 1: measure of the predictive mean (or the summed population variance).
template <int CL, int LPP, bool EXACT>
__device__ __forceinline__ float conf_multi(const float (&nmu)[CL], float m2s, int nvalid, const ScoreParams& p) {
  if (p.measure == kVariance) {
    if constexpr (LPP > 1) m2s = group_sum<LPP>(m2s);
    return fmaf(-m2s, p.inv_T, 1.0f);
  }
  if (p.measure == kEntropy) {
    float h = 0.f;
#pragma unroll
    for (int j = 0; j < CL; ++j)
      if (EXACT || j < nvalid) h = fmaf(nmu[j], lg2_approx(kTiny - nmu[j]), h);  // -sum mu*log2(mu + tiny), bits
    if constexpr (LPP > 1) h = group_sum<LPP>(h);
    return fmaf(-h, p.inv_log2_c, 1.0f);
  }
  float m1 = -nmu[0], m2 = -INFINITY;
#pragma unroll
  for (int j = 1; j < CL; ++j) {
    if (EXACT || j < nvalid) {
      m2 = fmaxf(m2, fminf(m1, -nmu[j]));
      m1 = fmaxf(m1, -nmu[j]);
    }
  }
  if constexpr (LPP > 1) group_top2<LPP>(m1, m2);
  return p.measure == kMargin ? m1 - m2 : m1;
}

// ---- per-image accumulation ---------------------------------------------------------------
// f64 mean of the f32 map (:261-263) done as an exact integer sum of round(conf * 2^shift):
// integer adds commute, so the result is independent of CTA scheduling (run-to-run identical).
// All CTAs walk the images in step, so every warp flushes to the same image at about the same
// time; kAccReplicas copies of the accumulator vector (picked by warp and CTA, laid out
// [replica][acc_stride] so replicas sit in different L2 slices) keep those REDs from serialising
// on one address.  finalize_kernel adds the replicas up.
__device__ __forceinline__ long long acc_slot(const ScoreParams& p) {
  const int replica = ((threadIdx.x >> 5) & 7) | ((blockIdx.x & (kAccReplicas / 8 - 1)) << 3);
  return static_cast<long long>(replica) * p.acc_stride;
}

struct ImageAcc {
  long long sum = 0;
  unsigned int nan = 0;
  long long img = -1;

  // a pixel known to belong to image `img`; cvt.rni.s64.f32 turns NaN into 0, the flag records it
  __device__ __forceinline__ void add(float conf, float fx_scale) {
    nan |= (conf != conf) ? 1u : 0u;
    sum += __float2ll_rn(conf * fx_scale);
  }

  // bf16 logits (fixed-point scale 2^22, capi.cu: fx_shift_for): the same integer as add() gives -- round-to-nearest-even
  // of conf * 2^22 -- without the F2I.S64 on the XU pipe: conf + 3 lies in [2, 4] for conf in [-1, 1], where a float's
  // mantissa IS (conf + 1) * 2^22 on a grid of 2^-22 (3 * 2^22 is even, so the tie rule is the same); one FADD and one
  // integer add on the FMA / ALU pipes per pixel.  NaN: garbage here, reported through the flag as in add().
  __device__ __forceinline__ void add_q22(float conf) {
    nan |= (conf != conf) ? 1u : 0u;
    sum += static_cast<int>(__float_as_uint(conf + 3.0f) - 0x40400000u);
  }
  // the same for the N pixels of a thread at once: the raw mantissa words are summed in 32 bits (each is below 2^23 in
  // magnitude once the bias is taken off, N <= 4) and widened once
  template <int N>
  __device__ __forceinline__ void add_q22(const float (&conf)[N]) {
    static_assert(N <= 64, "32-bit partial sum");
    unsigned int q = 0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q += __float_as_uint(conf[k] + 3.0f);
      bad |= conf[k] != conf[k];
    }
    nan |= bad ? 1u : 0u;
    sum += static_cast<int>(q - static_cast<unsigned int>(N) * 0x40400000u);
  }

  __device__ __forceinline__ void flush(const ScoreParams& p) {  // warp-collective
    const long long s = warp_sum_ll(sum);
    const unsigned int n = __any_sync(0xffffffffu, nan != 0);
    if ((threadIdx.x & 31) == 0 && img >= 0) {
      if (s != 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.acc + acc_slot(p) + img), static_cast<unsigned long long>(s));
      if (n) atomicOr(p.flags + img, 1u);
    }
    sum = 0;
    nan = 0;
  }
};

// Per-pixel outputs of a pixel of a FULL tile that lies inside one image (every tile but ~1 per image): no range or
// image-boundary checks, the sum goes through ImageAcc::add like on the ranked path.
__device__ __forceinline__ void emit_pixel_full(const ScoreParams& p, ImageAcc& acc, float conf, int lbl, long long g) {
  acc.add(conf, p.fx_scale);
  if (p.conf_map) p.conf_map[g] = conf;
  if (p.mask) p.mask[g] = (conf < p.threshold) ? 0 : 1;  // :265-269
  if (p.label) p.label[g] = static_cast<uint8_t>(lbl);
}

// l: pixel slot in the tile; in_img: slots below it belong to the tile's first image (acc.img).
__device__ __forceinline__ void emit_pixel(const ScoreParams& p, ImageAcc& acc, float conf, int lbl,
                                           long long tile_pix0, long long off, int l, int in_img) {
  const bool isnan_ = !(conf == conf);
  const long long fx = isnan_ ? 0ll : __float2ll_rn(conf * p.fx_scale);
  const long long g = tile_pix0 + l;
  if (l < in_img) {
    acc.sum += fx;
    acc.nan |= isnan_ ? 1u : 0u;
  } else {  // tile straddles an image boundary: rare, go straight to the image's accumulator
    const long long q = acc.img + (off + l) / p.P;
    atomicAdd(reinterpret_cast<unsigned long long*>(p.acc + acc_slot(p) + q), static_cast<unsigned long long>(fx));
    if (isnan_) atomicOr(p.flags + q, 1u);
  }
  if (p.conf_map) p.conf_map[g] = conf;
  if (p.mask) p.mask[g] = (conf < p.threshold) ? 0 : 1;  // :265-269
  if (p.label) p.label[g] = static_cast<uint8_t>(lbl);
}

}  // namespace als
