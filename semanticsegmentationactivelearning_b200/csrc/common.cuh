// Shared device helpers for the alscore kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "alscore kernels are written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace als {

constexpr int kConsumerThreads = 256;              // 8 compute warps
constexpr int kBlockThreads = kConsumerThreads + 32;  // + 1 bulk-copy producer warp
constexpr int kMaxStages = 8;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// active_learning.py:40  EPSILON = np.finfo(np.float32).tiny
constexpr float kTiny = 1.17549435e-38f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// Stage release that cannot overtake the loads of the stage (WAR on a TMA-filled buffer).  `mbarrier.arrive` does not
// wait for the arriving thread's in-flight ld.shared: the producer may see the slot free and the bulk copy may overwrite
// bytes that have not been read yet.  The arrival is therefore made DATA dependent on the loaded registers: it is
// predicated on `dep != never`, where `dep` is a word computed from every loaded value and `never` is a launch parameter
// (always 0xffffffff; `dep` has its top bit clear, so the test is always true) -- a RUN-TIME value on purpose: against a
// literal, ptxas proves the test from the shift that clears the top bit, folds it and drops the dependency (round 2
// shipped that for a while: the arrival then sits wherever the scheduler puts it, usually late enough, and a single
// class value of a single pixel was read stale about once in 300 streamed runs, profiles/r02_mc_single_pixel.txt).
__device__ __forceinline__ void mbar_arrive_after_loads(uint64_t* bar, uint32_t dep, uint32_t never) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %1, %2;\n\t"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(dep), "r"(never)
      : "memory");
}

// Warp-level form: ONE arrival per warp, issued by lane 0, that waits for the loads of EVERY lane -- the OR-reduction
// (REDUX) needs each lane's word, so it holds the arrival back for all of them however the lanes were scheduled.
__device__ __forceinline__ void warp_release_after_loads(uint64_t* bar, uint32_t dep, int lane, uint32_t never) {
  const uint32_t all = __reduce_or_sync(0xffffffffu, dep);
  if (lane == 0) mbar_arrive_after_loads(bar, all, never);
}

// For waits that are expected to be long: back off between polls so that the spinning warp does not take
// issue slots from the warps doing the work (the SM's arbiter favours higher warp ids).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(40);
}

// ---- 1-D bulk async copy (TMA engine, SASS UBLKCP) -----------------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared, completion counted in bytes on `bar`.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- programmatic dependent launch ------------------------------------------------------------------------
// Consecutive kernels of a pool pass are tiny next to their launch latency when batches are small (8 images =
// ~50 us of HBM time).  A kernel launched with launch_pdl() may start while its predecessor in the stream is still
// running; it must call pdl_wait() before it touches anything an earlier kernel wrote (then it sees all of it).
// pdl_launch_dependents() lets the NEXT such kernel become resident early.  Both are no-ops without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, func, args);
}

// ---- math -----------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32 pairs (sm_100a: FFMA2 / FADD2 / FMUL2, two fp32 lanes per issue slot) ---------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float hsum2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}

// ---- exp2 on the FMA pipe (bring-up option ALS_POLY_PAIRS, measured in DESIGN.md; the shipped build does not use it) ----
// Two exp2 per call without touching the MUFU (XU) pipe: round-to-nearest split t = n + f by the 1.5*2^23 magic add,
// degree-6 polynomial for 2^f on [-0.5, 0.5] (relative error 2.4e-7 in fp32, the same as ex2.approx), exponent inserted with an integer shift-add.
// t is clamped at -126 (2^t flushes to ~1e-38 there, like ex2.approx.ftz returns 0 a little further down).
// Cost: 3 + 6 packed FMA-pipe instructions + 2 FMNMX + 2 LEA for two results, against 2 MUFU.EX2 issue slots.
__device__ __forceinline__ f32x2 ex2_poly2(f32x2 t2) {
  float t0, t1;
  unpack2(t2, t0, t1);
  t2 = pack2(fmaxf(t0, -126.0f), fmaxf(t1, -126.0f));
  const f32x2 magic = pack2(12582912.0f, 12582912.0f), nmagic = pack2(-12582912.0f, -12582912.0f);
  const f32x2 one = pack2(1.0f, 1.0f), mone = pack2(-1.0f, -1.0f);
  const f32x2 r2 = add2(t2, magic);          // low mantissa bits = n = round(t)
  const f32x2 n2 = add2(r2, nmagic);         // n as a float (exact)
  const f32x2 f2 = fma2(n2, mone, t2);       // f = t - n in [-0.5, 0.5]
  f32x2 p = pack2(1.5403530e-4f, 1.5403530e-4f);
  p = fma2(p, f2, pack2(1.3333558e-3f, 1.3333558e-3f));
  p = fma2(p, f2, pack2(9.6181291e-3f, 9.6181291e-3f));
  p = fma2(p, f2, pack2(5.5504109e-2f, 5.5504109e-2f));
  p = fma2(p, f2, pack2(2.4022651e-1f, 2.4022651e-1f));
  p = fma2(p, f2, pack2(6.9314718e-1f, 6.9314718e-1f));
  p = fma2(p, f2, one);
  float p0, p1, r0, r1;
  unpack2(p, p0, p1);
  unpack2(r2, r0, r1);
  p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
  return pack2(p0, p1);
}

// max that propagates NaN (fmaxf drops it): keeps a NaN logit visible after the -inf clamp
__device__ __forceinline__ float max_nan(float a, float b) {
  float y;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b));
  return y;
}

// xor-shuffle reductions inside groups of G adjacent lanes (G power of two)
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace als
