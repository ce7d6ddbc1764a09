// Fused ENet classifier head + pool scoring for sm_100a: the logits never reach HBM.
//
// Reference: `Final.call` (/root/reference/models/enet/enet_modules.py:1359-1381) is a 3x3 stride-2
// transposed convolution 16 -> C without bias that turns the last feature map [N,h,w,16] into the logits
// [N,2h,2w,C] scored by active_learning.py:239-263.  Reading 16 B of features per OUTPUT pixel instead of
// 4*C B of logits removes the HBM bound of score.cu; the contraction (2.25*16*C MAC per output pixel) runs
// on the 5th-generation tensor cores, the softmax / confidence math on the CUDA cores straight out of
// tensor memory.
//
// GEMM view.  Input pixel (i,j) owns the output quad (2i+dy, 2j+dx).  With SAME padding (extra row /
// column at the bottom / right) the quad depends on the four input pixels (i-oy, j-ox), oy,ox in {0,1}:
//   out(0,0) = Y[i,j]K00 + Y[i-1,j]K20 + Y[i,j-1]K02 + Y[i-1,j-1]K22      out(0,1) = Y[i,j]K01 + Y[i-1,j]K21
//   out(1,0) = Y[i,j]K10 + Y[i,j-1]K12                                     out(1,1) = Y[i,j]K11
// A tile is 128 consecutive input pixels of one row (UMMA M = 128); its accumulator has one block of
// CB = round_up(C,4) columns per quad pixel, in the order (0,1) (0,0) (1,0) (1,1) so that every source
// pixel contributes to ONE contiguous column range: four narrow MMAs (N = 4CB, 2CB, 2CB, CB rounded up to
// 16) instead of one 64 x 4C GEMM that is 44 % zeros.
//
// Precision.  kind::tf32 keeps 11 significand bits, the parity bar is 1e-5 on confidences, so both
// operands are split x = hi + lo (hi = x rounded to tf32, lo = x - hi) and three products are accumulated in
// fp32: hi*lo + lo*hi + hi*hi (the dropped lo*lo term and the tf32 rounding of the lo parts are ~2^-24 relative each).
//
// Data movement.  The feature map is described to the TMA engine as a 4-D tensor [T*N][h][w][16 floats]; a feature
// row segment of 129 pixels (one halo pixel on the left) arrives as ONE tiled copy with a box of 16 floats x 129
// pixels and the 64-byte swizzle (cp.async.bulk.tensor, SASS UTMALDG): the engine stores the 16-byte chunk c of pixel
// px at chunk position c ^ ((px >> 1) & 3) inside the pixel's 64 bytes, so a thread can read ITS pixel with four
// 16-byte loads without bank conflicts (unswizzled, a pixel is half a bank row and eight lanes collide four ways),
// and it zero-fills everything outside the image -- the row above the first, the pixel left of column 0, the columns
// right of the last -- which is exactly TF's SAME padding.  (First version: 1-D bulk copies + a transpose through
// shared memory with two bar.sync per row.  Second: four unswizzled boxes of 4 floats x 129 pixels = chunk planes;
// 516 sixteen-byte box rows per feature row kept the TMA engine busy for ~12 % of the kernel.)  "Loader" warps read
// their pixel and its left neighbour,
// split every value into hi/lo in registers and write the A operands into TENSOR MEMORY with
// tcgen05.st (lane = pixel, 16 columns = channels): own pixel and left neighbour, hi and lo = 64 columns per
// feature row, a ring of four rows.  The MMAs take A from tensor memory and only the (small) weight
// operand from shared memory -- with A in shared memory every one of the 24 MMAs of a tile re-read a 4 KB
// A tile and the kernel was shared-memory bound at ~1000 clk / tile (profiles/r01_head_*).  A CTA walks
// down a strip of 128 columns: each feature row is loaded once and used by two consecutive tiles.
//
// Roles (576 threads, 1 CTA / SM, persistent, static round-robin over units of R rows x 128 columns):
//   warps 0-11  epilogue: 3 accumulator stages x 4 lane quarters (2 stages above 20 classes)
//   warps 12-15 loaders: swizzled rows -> hi/lo A rows in TMEM (T > 1: two groups of four warps, alternate rows)
//   warp 16     MMA issuer (one elected lane), TMEM owner      warp 17  producer: TMA tile copies
//
// Monte-Carlo samples (T > 1, features [T,N,h,w,16]).  The Welford state of a pixel (running mean per class + summed
// M2) has to stay in registers over the T samples, so the samples of ONE tile run back to back: accumulator q =
// (tile, sample t), each with its own previous and current feature row (a row is therefore loaded and split twice,
// the second time out of L2), and the epilogue is organised by BLOCK instead of by stage: 8 warps (lane quarter x
// quad-pixel pair) each take two pixels per thread of EVERY accumulator, their two update chains interleaved.  Measured
// (profiles/head_trace.py, per-role clock64 timeline): the three roles -- loaders (2 rows per accumulator), the MMA
// issuer (24 tcgen05.mma at ~25 clk of issue each + ~100 clk per mbarrier round trip) and the lock-stepped epilogue
// warps (MUFU phase, then FMA phase) -- each need 1000-1300 clk per accumulator and couple through the 4-row A ring
// and the 3 accumulator stages to ~1750 clk; knocking any single role out gains only 10-16 %.
//
// Issue rate of the tensor pipe (measured): a tcgen05.mma of this shape (M = 128, K = 8, N <= 80) occupies the pipe
// for >= ~21-25 clk however narrow N is, and a second issuer warp does not change that (two issuers taking alternate
// tiles: +3 % at C = 6, -4 % at C = 19) -- the 24 MMAs of a tile are a ~550-600 clk floor at every class count,
// which is what bounds the kernel below C ~ 12 (770-850 clk per tile at C = 2...6).
#include "head.cuh"

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "pixel_math.cuh"
#include "tc05.cuh"

namespace als {

#ifdef ALS_HEAD_TRACE
// Bring-up instrumentation (never built into the shipped library): SM-clock timestamps of CTA 0's first tiles.
//   [tile][0..1] splitter start/end of the tile's row, [2..3] MMA issue start/end, [4] epilogue sees the accumulator,
//   [5] epilogue released it, [6] epilogue done, [7] producer issued the row's copy;  T > 1 (tile = accumulator =
//   (tile, sample)): [0..1] is the CURRENT row, [8..9] the previous row's splitter start/end, [10] / [11] the moment
//   the previous / current row's loader got its tensor-memory slot back from the MMAs
constexpr int kTraceTiles = 1024;
constexpr int kTraceSlots = 16;  // T > 1: [12] [13] [14] = MMA warp past its wait for the previous row / current row / accumulator stage
__device__ long long g_head_trace[kTraceTiles][kTraceSlots];
#define ALS_TRACE(tile, slot)                                                         \
  do {                                                                                \
    if (blockIdx.x == 0 && (tile) < kTraceTiles) g_head_trace[(tile)][(slot)] = clock64(); \
  } while (0)
#else
#define ALS_TRACE(tile, slot) \
  do {                        \
  } while (0)
#endif

namespace {

constexpr int kRawStages = 6;
constexpr int kARing = 4;                             // feature rows resident in tensor memory (2 in use by the MMAs, 2 being written)
constexpr int kARowCols = 64;                         // [own hi 16 | own lo 16 | left hi 16 | left lo 16]
constexpr int kTmemCols = 512;
constexpr int kSlotPx = kHeadTileQuads + 1;           // 129 pixels: one halo pixel on the left
constexpr int kRowCopyBytes = kSlotPx * 64;           // 8256: one TMA box = 129 pixels x 16 channels
constexpr int kRawSlotBytes = 8704;                   // ... in a slot that keeps the 512-byte phase of the 64-byte swizzle
constexpr int kRawAlign = 512;
static_assert(kRawSlotBytes % kRawAlign == 0 && kRawSlotBytes >= kRowCopyBytes, "raw ring slots");
constexpr int kMaxLoaderGroups = 2;                   // groups of four warps taking alternate feature rows
constexpr int kLoaderThreads = 128;                   // per group
constexpr int kHeaderBytes = 512;

struct Unit {
  int n, i0, rows, j0, valid;
};

__device__ __forceinline__ Unit decode_unit(const HeadParams& p, long long u) {
  Unit un;
  const int strip = static_cast<int>(u % p.n_strips);
  const long long v = u / p.n_strips;
  const int rb = static_cast<int>(v % p.n_rowblocks);
  un.n = static_cast<int>(v / p.n_rowblocks);
  un.i0 = rb * p.rows_per_unit;
  un.rows = min(p.rows_per_unit, p.h - un.i0);
  un.j0 = strip * kHeadTileQuads;
  un.valid = min(kHeadTileQuads, p.w - un.j0);
  return un;
}

__host__ __device__ constexpr int cround(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// Compile-time twin of head_geometry(C) (the host packs the weights with the run-time one; a static_assert-free
// consistency check runs in plan_head).
//
// EPB = accumulator blocks (quad pixels) one epilogue warp owns.  4: single sample -- a warp serves ONE accumulator
// stage and all four blocks of its lane quarter.  2 or 1: Monte-Carlo samples (T > 1) -- the per-class Welford state
// of four pixels does not fit one thread, so the blocks are split over 16 / EPB warps that all work on the SAME
// accumulator and keep the state of their EPB pixels in registers across the T samples of a tile.
template <int C, int EPB = 4, bool MULTI_ = false>
struct Geom {
  static_assert(EPB == 4 || EPB == 2 || EPB == 1, "blocks per epilogue warp");
  static constexpr bool MULTI = MULTI_;
  static constexpr int CB = cround(C, 4);
  static constexpr int N1 = cround(2 * CB, 16), N2 = cround(2 * CB, 16), N3 = cround(CB, 16);
  static constexpr int N0 = cround(cmax(4 * CB, cmax(N1, cmax(CB + N2, CB + N3))), 16);  // also initialises every column
  static constexpr int ROWS = N0 + N1 + N2 + N3;
  // accumulator stages that fit beside the A-row ring in the 512 tensor-memory columns (3 up to C = 20, else 2)
  static constexpr int ACC_STAGES = (cround(3 * N0, 32) + kARing * kARowCols <= kTmemCols) ? 3 : 2;
  static constexpr int A_COL0 = cround(ACC_STAGES * N0, 32);  // first column of the A-row ring
  static_assert(A_COL0 + kARing * kARowCols <= kTmemCols, "tensor memory budget");
  // Warp roles by warp id.  The SM's issue arbiter favours HIGHER warp ids, so the roles on the critical path
  // (producer, MMA issuer, loaders) sit above the epilogue warps, which mostly wait.
  static constexpr int EPI_WARPS = MULTI ? 16 / EPB : 4 * ACC_STAGES;
  static constexpr int ACC_ARRIVALS = MULTI ? EPI_WARPS : 4;  // epilogue warps that release one accumulator stage
  static constexpr int FIRST_LOADER_WARP = EPI_WARPS;  // warps below: epilogue, lane quarter = warp & 3 (stage or block group = warp >> 2)
  // Loader groups of four warps taking alternate feature rows.  A single group keeps up now that the TMA engine lays
  // the rows out (the group only splits and stores them), and the warps saved buy registers for the epilogue:
  // T = 1: 576 threads / 96 registers, +2.5 % at C = 19 and +8 % at C = 6 against two groups;  T > 1 (two rows per
  // accumulator): 448 threads / 128 registers with two pixels per epilogue thread, +8 % against 832 threads.
#ifdef ALS_HEAD_LOADER_GROUPS_MC  // bring-up: loader groups of the T > 1 kernel
  static constexpr int LOADER_GROUPS = MULTI ? ALS_HEAD_LOADER_GROUPS_MC : 1;
#else
  static constexpr int LOADER_GROUPS = 1;
#endif
  static_assert(LOADER_GROUPS <= kMaxLoaderGroups, "loader groups");
  static constexpr int MMA_WARP = FIRST_LOADER_WARP + 4 * LOADER_GROUPS;
  static constexpr int PRODUCER_WARP = MMA_WARP + 1;
#ifdef ALS_HEAD_MC_SETMAXNREG
  // bring-up: T > 1 with register re-allocation between the roles (setmaxnreg works on warpgroups of 4 warps: the
  // MMA / producer warps get two idle companions so that every role is a whole number of warpgroups)
  static constexpr bool REALLOC = MULTI;
  static constexpr int THREADS = MULTI ? 32 * (MMA_WARP + 4) : 32 * (PRODUCER_WARP + 1);
#else
  static constexpr bool REALLOC = false;
  static constexpr int THREADS = 32 * (PRODUCER_WARP + 1);   // 576 (3 stages) or 448 (2 stages); T > 1: 448 (EPB = 2)
#endif
};

#ifndef ALS_HEAD_REGS_EPI
#define ALS_HEAD_REGS_EPI 128
#define ALS_HEAD_REGS_LOADER 72
#define ALS_HEAD_REGS_MMA 56
#endif
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// TMEM -> registers, CNT consecutive columns (CNT multiple of 4, <= 32)
template <int CNT>
__device__ __forceinline__ void ld_cols(uint32_t taddr, float (&v)[CNT]) {
  static_assert(CNT % 4 == 0 && CNT >= 4 && CNT <= 32, "column count");
  uint32_t r[CNT];
  int done = 0;
  if constexpr (CNT == 32) {
    float t[32];
    tc05::ld32(taddr, t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = t[i];
    return;
  }
  if constexpr (CNT & 16) {
    float t[16];
    tc05::ld16(taddr + done, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[done + i] = t[i];
    done += 16;
  }
  if constexpr (CNT & 8) {
    float t[8];
    tc05::ld8(taddr + done, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[done + i] = t[i];
    done += 8;
  }
  if constexpr (CNT & 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr + done)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[done + i] = __uint_as_float(r[i]);
  }
}

}  // namespace

// one pixel (16 channels as 4 float4) -> hi and lo parts (2 x 16 tensor-memory columns of this thread's lane).
// The tensor core TRUNCATES its fp32 inputs to tf32 (probes/umma_probe.cu), so hi must be rounded here:
// Veltkamp's split with 2^13 + 1 gives hi = x rounded to nearest on 11 significand bits and lo = x - hi exactly,
// two packed FMA-pipe instructions per element pair each.  lo (|lo| <= 2^-11 |x|) is left to the hardware's
// truncation: an error of at most 2^-21 |x|, the same order as the dropped lo*lo term.
__device__ __forceinline__ void split_hi_lo(const float4 (&v)[4], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#ifdef ALS_HEAD_TRUNC_SPLIT
  // bring-up: let the tensor core's own truncation make hi (pass x itself), lo = x - trunc13(x): one LOP + one FADD
  // per element instead of Veltkamp's four FMAs per pair; |lo| < 2^-10 |x| instead of <= 2^-11 |x|
  const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t u = __float_as_uint(f[i]);
    hi[i] = u;
    lo[i] = __float_as_uint(f[i] - __uint_as_float(u & 0xffffe000u));
  }
  return;
#endif
  const f32x2 k2 = pack2(8193.0f, 8193.0f), m1 = pack2(-1.0f, -1.0f);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const f32x2 x2[2] = {pack2(v[c].x, v[c].y), pack2(v[c].z, v[c].w)};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const f32x2 cc = fma2(x2[j], k2, pack2(0.f, 0.f));  // c = x * (2^13 + 1)
      const f32x2 t = fma2(x2[j], m1, cc);                 // c - x
      const f32x2 h = fma2(t, m1, cc);                     // hi = c - (c - x)
      const f32x2 l = fma2(h, m1, x2[j]);                  // lo = x - hi
      float a0, a1;
      unpack2(h, a0, a1);
      hi[4 * c + 2 * j] = __float_as_uint(a0);
      hi[4 * c + 2 * j + 1] = __float_as_uint(a1);
      unpack2(l, a0, a1);
      lo[4 * c + 2 * j] = __float_as_uint(a0);
      lo[4 * c + 2 * j + 1] = __float_as_uint(a1);
    }
  }
}

template <int C, int MEASURE, int EPB>
__global__ void __launch_bounds__(Geom<C, EPB, (MEASURE == kMulti)>::THREADS, 1)
score_head_kernel(const HeadParams p, const __grid_constant__ CUtensorMap feat_map) {
  using G = Geom<C, EPB, (MEASURE == kMulti)>;
  constexpr bool MULTI = G::MULTI;  // T > 1: every tile is computed once per Monte-Carlo sample
  static_assert(MULTI == (MEASURE == kMulti), "T > 1 runs the kMulti instantiation (measure is a run-time field)");
  constexpr int CB = G::CB;
  constexpr int kAccStages = G::ACC_STAGES;
  constexpr int kAccStride = G::N0;   // accumulator columns per stage
  constexpr int kACol0 = G::A_COL0;   // first column of the A-row ring
  constexpr int kFirstLoaderWarp = G::FIRST_LOADER_WARP, kMmaWarp = G::MMA_WARP, kProducerWarp = G::PRODUCER_WARP;
  constexpr int kFirstEpilogueWarp = 0;
  constexpr int wpart_bytes = 4 * G::ROWS * 16;  // one precision part of the packed weights

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_raw = full_raw + kRawStages;
  uint64_t* full_a = empty_raw + kRawStages;
  uint64_t* empty_a = full_a + kARing;
  uint64_t* full_acc = empty_a + kARing;
  uint64_t* empty_acc = full_acc + kAccStages;
  uint64_t* wbar = empty_acc + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  // the raw ring starts on a 512-byte boundary of the shared-memory WINDOW (the swizzle is a function of the address)
  unsigned char* raw_base = smem + kHeaderBytes + ((kRawAlign - ((smem_u32(smem) + kHeaderBytes) & (kRawAlign - 1))) & (kRawAlign - 1));
  unsigned char* w_base = raw_base + kRawStages * kRawSlotBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRawStages; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&empty_raw[s], kLoaderThreads / 32);  // the four warps (lane quarters) of the group that takes the row
    }
    for (int s = 0; s < kARing; ++s) {
      mbar_init(&full_a[s], kLoaderThreads / 32);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&full_acc[s], 1);
      mbar_init(&empty_acc[s], G::ACC_ARRIVALS);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tc05::tmem_alloc<kTmemCols>(tmem_slot);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_launch_dependents();  // the finalize kernel behind this one may become resident now (it waits for our sums)

  if (warp >= kMmaWarp) {
  // (REALLOC builds: this warpgroup = MMA issuer, producer and two idle companions; it gives registers back)
  if constexpr (G::REALLOC) setmaxnreg_dec<ALS_HEAD_REGS_MMA>();
  if (warp == kProducerWarp) {
    // ===== producer =====
    if (lane == 0) {
      // evict_first although the four plane copies of a row touch the same lines and T > 1 re-reads a row T samples
      // later: measured 1.3 % faster than evict_normal on cfg1h and cfg2h (the re-reads still hit: the lines live
      // long enough in the 126 MB L2)
      const uint64_t policy = l2_policy_evict_first();
      mbar_arrive_expect_tx(wbar, 2u * wpart_bytes);
      bulk_g2s(w_base, p.weights, 2u * wpart_bytes, wbar, policy);
      int s = 0;
      uint32_t ph = 0;
      long long ptile = 0;
      for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const Unit un = decode_unit(p, u);
        // feature row `row` (-1 = the padding row above the image) of sample t -> next raw slot, as four chunk planes.
        // Coordinates outside the tensor (row -1, pixel -1, pixels >= w) are zero-filled by the TMA engine and
        // still count towards the transaction bytes.
        auto issue_row = [&](int row, int t) {
          mbar_wait_relaxed(&empty_raw[s], ph ^ 1u);
          unsigned char* dst = raw_base + s * kRawSlotBytes;
          const int img = t * p.n_images + un.n;
          mbar_arrive_expect_tx(&full_raw[s], kRowCopyBytes);
          tc05::tma_load_4d(dst, &feat_map, 0, un.j0 - 1, row, img, &full_raw[s], policy);
          if (++s == kRawStages) { s = 0; ph ^= 1u; }
        };
        if constexpr (!MULTI) {
          // one sample: a row is loaded once and serves the tile below it as "previous" and its own as "current"
          for (int r = -1; r < un.rows; ++r) {
            if (r >= 0) { ALS_TRACE(ptile, 7); ++ptile; }
            issue_row(un.i0 + r, 0);
          }
        } else {
          // T samples: the Welford state of a tile lives in the epilogue's registers, so the samples of ONE tile
          // run back to back and every (tile, sample) brings its own previous and current row (the re-read of the
          // previous row comes out of L2: it was the current row of the tile above, T samples ago)
          for (int k = 0; k < un.rows; ++k)
            for (int t = 0; t < p.T; ++t) {
              issue_row(un.i0 + k - 1, t);
              ALS_TRACE(ptile, 7);
              ++ptile;
              issue_row(un.i0 + k, t);
            }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer: the whole warp walks the (warp-uniform) schedule, one elected lane issues =====
    const bool leader = tc05::elect_one();
    constexpr uint32_t kIdesc[4] = {tc05::idesc_tf32(kHeadTileQuads, G::N0), tc05::idesc_tf32(kHeadTileQuads, G::N1),
                                    tc05::idesc_tf32(kHeadTileQuads, G::N2), tc05::idesc_tf32(kHeadTileQuads, G::N3)};
    constexpr uint32_t kCol0[4] = {0, 0, G::CB, G::CB};
    constexpr uint32_t kRow0[4] = {0, G::N0, G::N0 + G::N1, G::N0 + G::N1 + G::N2};
    constexpr uint32_t kLboB = G::ROWS * 16;         // bytes between chunk planes of the packed weights
    constexpr uint32_t kWPart = 4 * kLboB;           // one precision part of the weights
    const uint64_t b_base = tc05::smem_desc(smem_u32(w_base), kLboB, 128);
    mbar_wait(wbar, 0);
    static_assert((kARing & (kARing - 1)) == 0, "ring positions are masked");
    uint32_t seq = 0;   // feature rows consumed so far: ring slot = seq & (kARing - 1), phase = (seq / kARing) & 1
    uint32_t tile = 0;  // (trace only)
    int a = 0;          // accumulator stage of the next tile and its phase
    uint32_t acc_ph = 0;
    // the 24 MMAs of one accumulator: current row in ring slot rb_cur, the row above it in rb_prev
    auto issue_tile = [&](int rb_prev, int rb_cur, bool release_cur) {
      ALS_TRACE(tile, 2);
      const uint32_t d = tmem + a * kAccStride;
      const uint32_t a_row[2] = {tmem + kACol0 + rb_cur * kARowCols, tmem + kACol0 + rb_prev * kARowCols};
      // The small cross terms go first and the hi*hi products last: the tensor core truncates when it adds
      // into the accumulator, so every MMA issued after the accumulator has reached full scale costs
      // ~2^-24 of it -- 8 such steps this way round instead of 24.
#pragma unroll
      for (int pc = 0; pc < 3; ++pc) {  // hi*lo, lo*hi, hi*hi
        const uint32_t a_part = (pc == 1) ? 16 : 0;        // columns: hi 0..15, lo 16..31
        const uint32_t b_part = (pc == 0) ? kWPart : 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a_col = ((o < 2) ? 0u : 32u) + a_part + ks * 8;  // own pixel for sources (.,j), left for (.,j-1)
            const uint32_t b_off = (b_part + kRow0[o] * 16 + ks * 2 * kLboB) >> 4;
            tc05::mma_tf32_ts(d + kCol0[o], a_row[o & 1] + a_col, b_base + b_off, kIdesc[o],
                              !(pc == 0 && o == 0 && ks == 0));
          }
        }
      }
      tc05::commit(&full_acc[a]);
      tc05::commit(&empty_a[rb_prev]);
      if (release_cur) tc05::commit(&empty_a[rb_cur]);
      ALS_TRACE(tile, 3);
    };
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit un = decode_unit(p, u);
      if constexpr (!MULTI) {
        // halo row of the unit
        mbar_wait(&full_a[seq & (kARing - 1)], (seq / kARing) & 1u);
        for (int k = 0; k < un.rows; ++k) {
          const uint32_t sp = seq + k, sc = seq + k + 1;
          const int rb_prev = sp & (kARing - 1), rb_cur = sc & (kARing - 1);
          mbar_wait(&full_a[rb_cur], (sc / kARing) & 1u);
          mbar_wait(&empty_acc[a], acc_ph ^ 1u);
          tc05::fence_after_sync();
          if (leader) issue_tile(rb_prev, rb_cur, k == un.rows - 1);
          __syncwarp();
          ++tile;
          if (++a == kAccStages) { a = 0; acc_ph ^= 1u; }
        }
        seq += un.rows + 1;
      } else {
        // T > 1: one accumulator per (tile, sample), each with its own pair of rows
        const int n_acc = un.rows * p.T;
        for (int q = 0; q < n_acc; ++q, seq += 2, ++tile) {
          const int rb_prev = seq & (kARing - 1), rb_cur = (seq + 1) & (kARing - 1);
          if (leader) ALS_TRACE(tile, 15);
          mbar_wait(&full_a[rb_prev], (seq / kARing) & 1u);
          if (leader) ALS_TRACE(tile, 12);
          mbar_wait(&full_a[rb_cur], ((seq + 1) / kARing) & 1u);
          if (leader) ALS_TRACE(tile, 13);
          mbar_wait(&empty_acc[a], acc_ph ^ 1u);
          if (leader) ALS_TRACE(tile, 14);
          tc05::fence_after_sync();
          if (leader) issue_tile(rb_prev, rb_cur, true);
          __syncwarp();
          if (++a == kAccStages) { a = 0; acc_ph ^= 1u; }
        }
      }
    }
  }
  } else if (warp >= kFirstLoaderWarp) {
    // ===== loaders: raw NHWC row -> chunk planes (shared) -> hi / lo A rows in tensor memory =====
    if constexpr (G::REALLOC) setmaxnreg_dec<ALS_HEAD_REGS_LOADER>();
    const int grp = (warp - kFirstLoaderWarp) >> 2;      // rows alternate between the two groups
    const int lt = threadIdx.x - 32 * kFirstLoaderWarp - grp * kLoaderThreads;  // 0..127 inside the group
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may write
    const int m = quarter * 32 + lane;                   // A row = accumulator row = quad column inside the strip
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + kACol0;
    uint32_t rs = 0;      // feature rows seen so far: A-ring slot = rs & (kARing - 1)
    int s = 0;            // raw-ring slot of row rs and its phase
    uint32_t phs = 0;
    uint32_t stile = 0;   // (trace only)
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit un = decode_unit(p, u);
      const int n_rows = MULTI ? 2 * un.rows * p.T : un.rows + 1;  // feature rows the producer sends for this unit
      for (int vr = 0; vr < n_rows; ++vr, ++rs) {
        const int r = MULTI ? 0 : vr - 1;  // (trace only; single sample: row inside the unit)
        if (r >= 0 && !MULTI) ++stile;
        if (MULTI) stile = (rs >> 1) + 1;
        const int tr0 = (MULTI && !(rs & 1)) ? 8 : 0;  // (trace only) slot pair: previous row of a T > 1 accumulator -> 8, 9
        const int s_row = s;
        const uint32_t phs_row = phs;
        if (++s == kRawStages) { s = 0; phs ^= 1u; }
        if (G::LOADER_GROUPS > 1 && (rs & (G::LOADER_GROUPS - 1)) != static_cast<uint32_t>(grp)) continue;
        const int rb = rs & (kARing - 1);
        mbar_wait(&full_raw[s_row], phs_row);
        if (r >= 0 && lt == 0) ALS_TRACE(stile - 1, tr0);
        // this thread's pixel (slot pixel m + 1) and its left neighbour (slot pixel m): 16-byte reads from the four
        // chunk planes, consecutive lanes = consecutive 16-byte words
        // 64-byte swizzle: 16-byte chunk c of slot pixel px sits at chunk position c ^ ((px >> 1) & 3) of its 64 bytes,
        // so the eight lanes of a quarter-warp (consecutive pixels, same c) hit eight different 4-bank groups
        const unsigned char* rowp = raw_base + s_row * kRawSlotBytes;
        const unsigned char* own_p = rowp + (m + 1) * 64;
        const unsigned char* left_p = rowp + m * 64;
        const int own_x = ((m + 1) >> 1) & 3, left_x = (m >> 1) & 3;
        float4 own[4], left[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          own[c] = *reinterpret_cast<const float4*>(own_p + ((c ^ own_x) << 4));
          left[c] = *reinterpret_cast<const float4*>(left_p + ((c ^ left_x) << 4));
        }
        // release the raw slot only once the loads have READ it (common.cuh: mbar_arrive_after_loads)
        uint32_t dep = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) dep |= __float_as_uint(own[c].x) | __float_as_uint(left[c].x);
        warp_release_after_loads(&empty_raw[s_row], dep >> 1, lane, p.sp.never);
        // split in registers while the tensor-memory slot may still be in use, then only the stores wait for it
        uint32_t hi_o[16], lo_o[16], hi_l[16], lo_l[16];
        split_hi_lo(own, hi_o, lo_o);
        split_hi_lo(left, hi_l, lo_l);
        mbar_wait(&empty_a[rb], ((rs / kARing) & 1u) ^ 1u);  // the MMAs that read this slot completed
        tc05::fence_after_sync();
        if (MULTI && lt == 0) ALS_TRACE(stile - 1, tr0 ? 10 : 11);
        const uint32_t t_row = t_lane + rb * kARowCols;
        tc05::st16(t_row, hi_o);
        tc05::st16(t_row + 16, lo_o);
        tc05::st16(t_row + 32, hi_l);
        tc05::st16(t_row + 48, lo_l);
        tc05::st_wait();
        tc05::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[rb]);
        if (r >= 0 && lt == 0) ALS_TRACE(stile - 1, tr0 + 1);
      }
    }
  } else if constexpr (!MULTI) {
    // ===== epilogue, one sample: tensor memory -> confidences -> per-image sums =====
    const int e = warp - kFirstEpilogueWarp;
    const int a = e >> 2;          // accumulator stage this warp serves
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int m = quarter * 32 + lane;
    const uint32_t tbase = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + a * kAccStride;
    const ScoreParams& sp = p.sp;
    const int W = 2 * p.w;
    ImageAcc acc;
    uint32_t tile = 0;   // tiles of this CTA so far; this warp serves tiles a, a + kAccStages, ...
    uint32_t mine = a;
    uint32_t ph = 0;
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit un = decode_unit(p, u);
      if (un.n != acc.img) {
        acc.flush(sp);
        acc.img = un.n;
      }
      const bool valid = m < un.valid;
      for (int k = 0; k < un.rows; ++k, ++tile) {
        if (tile != mine) continue;
        mine += kAccStages;
        mbar_wait_relaxed(&full_acc[a], ph);
        ph ^= 1u;
        tc05::fence_after_sync();
        if (quarter == 0 && lane == 0) ALS_TRACE(tile, 4);
        float conf[4];
        int lbl[4];
        // two blocks at a time: the two pixels' dependent chains (max -> ex2 -> sums -> lg2 / rcp) interleave in the
        // instruction stream (+1.2 % at C = 19 on top of the single loader group that pays for the registers)
#pragma unroll
        for (int bp = 0; bp < 2; ++bp) {
          float v0[CB], v1[CB];
          ld_cols<CB>(tbase + (2 * bp) * CB, v0);
          ld_cols<CB>(tbase + (2 * bp + 1) * CB, v1);
          tc05::ld_wait();
          float x[2][C];
#pragma unroll
          for (int j = 0; j < C; ++j) { x[0][j] = v0[j]; x[1][j] = v1[j]; }
          if (bp == 1) {  // everything is in registers: hand the accumulator back to the MMA warp
            tc05::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_acc[a]);
            if (quarter == 0 && lane == 0) ALS_TRACE(tile, 5);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) lbl[2 * bp + q] = sp.label ? group_argmax<C, 1>(x[q], C, 0) : 0;
          conf[2 * bp] = conf_single<C, 1, true, MEASURE>(x[0], C, sp);
          conf[2 * bp + 1] = conf_single<C, 1, true, MEASURE>(x[1], C, sp);
          if constexpr (MEASURE == kEntropy) {
            if (__any_sync(0xffffffffu, !(conf[2 * bp] == conf[2 * bp]) || !(conf[2 * bp + 1] == conf[2 * bp + 1]))) {
              conf[2 * bp] = conf_single<C, 1, true, MEASURE, true>(x[0], C, sp);       // rare: -inf / NaN logits
              conf[2 * bp + 1] = conf_single<C, 1, true, MEASURE, true>(x[1], C, sp);
            }
          }
        }
        if (quarter == 0 && lane == 0) ALS_TRACE(tile, 6);
        if (valid) {
#pragma unroll
          for (int b = 0; b < 4; ++b) acc.add(conf[b], sp.fx_scale);
          if (sp.any_out) {
            const int i = un.i0 + k;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int dy = b >> 1, dx = (b == 0 || b == 3) ? 1 : 0;  // blocks: (0,1) (0,0) (1,0) (1,1)
              const long long g = (static_cast<long long>(un.n) * (2 * p.h) + (2 * i + dy)) * W + 2 * (un.j0 + m) + dx;
              if (sp.conf_map) sp.conf_map[g] = conf[b];
              if (sp.mask) sp.mask[g] = (conf[b] < sp.threshold) ? 0 : 1;  // active_learning.py:265-269
              if (sp.label) sp.label[g] = static_cast<uint8_t>(lbl[b]);
            }
          }
        }
      }
    }
    acc.flush(sp);
  } else {
    // ===== epilogue, T samples: every warp takes its EPB blocks of EVERY accumulator; the running mean of the
    // softmax (per class) and the summed M2 of its pixels stay in registers over the T samples of a tile =====
    if constexpr (G::REALLOC) setmaxnreg_inc<ALS_HEAD_REGS_EPI>();
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int blk0 = (warp >> 2) * EPB;       // first accumulator block (quad pixel) of this warp
    const int m = quarter * 32 + lane;
    const uint32_t tbase = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + blk0 * CB;
    const ScoreParams& sp = p.sp;
    const int W = 2 * p.w;
    ImageAcc acc;
    int a = 0;           // accumulator stage of the next (tile, sample) and its phase
    uint32_t ph = 0;
    uint32_t tq = 0;     // (trace only) accumulators so far
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit un = decode_unit(p, u);
      if (un.n != acc.img) {
        acc.flush(sp);
        acc.img = un.n;
      }
      const bool valid = m < un.valid;
      for (int k = 0; k < un.rows; ++k) {
        float nmu[EPB][C];  // negated running mean of the class probabilities (pixel_math.cuh: welford_update)
        float m2s[EPB];
        int lbl[EPB];
#pragma unroll
        for (int b = 0; b < EPB; ++b) {
          m2s[b] = 0.f;
          lbl[b] = 0;
#pragma unroll
          for (int j = 0; j < C; ++j) nmu[b][j] = 0.f;
        }
        for (int t = 0; t < p.T; ++t) {
          mbar_wait_relaxed(&full_acc[a], ph);
          tc05::fence_after_sync();
          if (warp == 0 && lane == 0) ALS_TRACE(tq, 4);
          const float inv_t = __frcp_rn(static_cast<float>(t + 1));
          const uint32_t tacc = tbase + a * kAccStride;
          if constexpr (EPB % 2 == 0) {
            // a pair of pixels at a time: both accumulator blocks first, then the two updates back to back (two
            // independent dependent chains in the instruction stream)
#pragma unroll
            for (int bp = 0; bp < EPB / 2; ++bp) {
              float v0[CB], v1[CB];
              ld_cols<CB>(tacc + (2 * bp) * CB, v0);
              ld_cols<CB>(tacc + (2 * bp + 1) * CB, v1);
              tc05::ld_wait();
              if (bp == EPB / 2 - 1) {  // this warp's share is in registers
                tc05::fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_acc[a]);
                if (warp == 0 && lane == 0) ALS_TRACE(tq, 5);
              }
              float x[2][C];
#pragma unroll
              for (int j = 0; j < C; ++j) { x[0][j] = v0[j]; x[1][j] = v1[j]; }
              if (t == 0 && sp.label) {
                lbl[2 * bp] = group_argmax<C, 1>(x[0], C, 0);
                lbl[2 * bp + 1] = group_argmax<C, 1>(x[1], C, 0);
              }
              welford_update<C, 1, true>(x[0], C, inv_t, nmu[2 * bp], m2s[2 * bp]);
              welford_update<C, 1, true>(x[1], C, inv_t, nmu[2 * bp + 1], m2s[2 * bp + 1]);
            }
          } else {
#pragma unroll
          for (int b = 0; b < EPB; ++b) {
            float v[CB];
            ld_cols<CB>(tacc + b * CB, v);
            tc05::ld_wait();
            if (b == EPB - 1) {  // this warp's share is in registers
              tc05::fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty_acc[a]);
              if (warp == 0 && lane == 0) ALS_TRACE(tq, 5);
            }
            float x[C];
#pragma unroll
            for (int j = 0; j < C; ++j) x[j] = v[j];
            if (t == 0 && sp.label) lbl[b] = group_argmax<C, 1>(x, C, 0);  // pseudo_label of sample 0, as in score.cu
            welford_update<C, 1, true>(x, C, inv_t, nmu[b], m2s[b]);
          }
          }
          if (warp == 0 && lane == 0) ALS_TRACE(tq, 6);
          ++tq;
          if (++a == kAccStages) { a = 0; ph ^= 1u; }
        }
        if (valid) {
          const int i = un.i0 + k;
#pragma unroll
          for (int b = 0; b < EPB; ++b) {
            const float conf = conf_multi<C, 1, true>(nmu[b], m2s[b], C, sp);
            acc.add(conf, sp.fx_scale);
            if (sp.any_out) {
              const int blk = blk0 + b;
              const int dy = blk >> 1, dx = (blk == 0 || blk == 3) ? 1 : 0;  // blocks: (0,1) (0,0) (1,0) (1,1)
              const long long g = (static_cast<long long>(un.n) * (2 * p.h) + (2 * i + dy)) * W + 2 * (un.j0 + m) + dx;
              if (sp.conf_map) sp.conf_map[g] = conf;
              if (sp.mask) sp.mask[g] = (conf < sp.threshold) ? 0 : 1;  // active_learning.py:265-269
              if (sp.label) sp.label[g] = static_cast<uint8_t>(lbl[b]);
            }
          }
        }
      }
    }
    acc.flush(sp);
  }

  tc05::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc05::tmem_dealloc<kTmemCols>(tmem);
}

// ---- host side -------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

HeadGeom head_geometry(int C) {
  HeadGeom g{};
  g.C = C;
  g.CB = round_up(C, 4);
  const int CB = g.CB;
  g.n[1] = round_up(2 * CB, 16);  g.col0[1] = 0;    // source (i-1, j)   -> blocks 0,1
  g.n[2] = round_up(2 * CB, 16);  g.col0[2] = CB;   // source (i, j-1)   -> blocks 1,2
  g.n[3] = round_up(CB, 16);      g.col0[3] = CB;   // source (i-1, j-1) -> block 1
  int n0 = round_up(4 * CB, 16);                    // source (i, j)     -> blocks 0..3; also initialises every column
  for (int o = 1; o < 4; ++o)
    if (g.col0[o] + g.n[o] > n0) n0 = round_up(g.col0[o] + g.n[o], 16);
  g.n[0] = n0;
  g.col0[0] = 0;
  int r = 0;
  for (int o = 0; o < 4; ++o) {
    g.row0[o] = r;
    r += g.n[o];
  }
  g.rows = r;
  return g;
}

static float host_round_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;  // inf / nan unchanged
  u = (u + 0x1000u) & 0xffffe000u;                 // round to nearest, ties away (cvt.rna.tf32.f32)
  memcpy(&x, &u, 4);
  return x;
}

size_t pack_head_weights(const float* kernel, int C, float* out) {
  const HeadGeom g = head_geometry(C);
  // tap (ky,kx) of source o for block b, -1 = none.  Filter layout [ky][kx][c][ch]  (enet_modules.py:1341)
  static const int tap[4][4][2] = {
      {{0, 1}, {0, 0}, {1, 0}, {1, 1}},      // Y[i, j]
      {{2, 1}, {2, 0}, {-1, -1}, {-1, -1}},  // Y[i-1, j]
      {{-1, -1}, {0, 2}, {1, 2}, {-1, -1}},  // Y[i, j-1]
      {{-1, -1}, {2, 2}, {-1, -1}, {-1, -1}},  // Y[i-1, j-1]
  };
  const size_t part = static_cast<size_t>(4) * g.rows * 4;  // floats per precision part
  for (size_t i = 0; i < 2 * part; ++i) out[i] = 0.f;
  for (int o = 0; o < 4; ++o)
    for (int nn = 0; nn < g.n[o]; ++nn) {
      const int col = g.col0[o] + nn;
      const int b = col / g.CB, c = col % g.CB;
      if (b >= 4 || c >= C || tap[o][b][0] < 0) continue;
      const float* w = kernel + ((static_cast<size_t>(tap[o][b][0]) * 3 + tap[o][b][1]) * C + c) * kHeadChannels;
      const int row = g.row0[o] + nn;
      for (int k = 0; k < kHeadChannels; ++k) {
        const float hi = host_round_tf32(w[k]);
        const float lo = host_round_tf32(w[k] - hi);
        const size_t at = (static_cast<size_t>(k / 4) * g.rows + row) * 4 + (k % 4);
        out[at] = hi;
        out[part + at] = lo;
      }
    }
  return 2 * part;
}

// Monte-Carlo variant (T > 1): blocks per epilogue warp.  Two pixels per thread (8 epilogue warps; with the single
// loader group 448 threads and up to 128 registers) hold the Welford state of up to kHeadMaxClassesMC classes and let
// the two pixels' update chains interleave; one pixel per thread (16 warps) measured 8 % slower, four pixels per
// thread (4 warps, 166 registers) 12 % slower.  ALS_HEAD_EPB
// overrides the choice in bring-up builds.
#ifdef ALS_HEAD_EPB
__host__ __device__ constexpr int epb_multi(int) { return ALS_HEAD_EPB; }
#else
__host__ __device__ constexpr int epb_multi(int) { return 2; }
#endif

template <int C>
static const void* pick_head(int measure, int T, const char** name, int* block) {
  if (T > 1) {
    if constexpr (C <= kHeadMaxClassesMC) {
      *name = "score_head_kernel<multi>";
      *block = Geom<C, epb_multi(C), true>::THREADS;
      return (const void*)score_head_kernel<C, kMulti, epb_multi(C)>;
    } else {
      return nullptr;
    }
  }
  *block = Geom<C>::THREADS;
  switch (measure) {
    case kEntropy: *name = "score_head_kernel<entropy>"; return (const void*)score_head_kernel<C, kEntropy, 4>;
    case kMargin: *name = "score_head_kernel<margin>"; return (const void*)score_head_kernel<C, kMargin, 4>;
    case kConfidence: *name = "score_head_kernel<confidence>"; return (const void*)score_head_kernel<C, kConfidence, 4>;
    default: return nullptr;
  }
}

// One instantiation per class count 2..32 (the reference's datasets use 19, 6 and 19: datasets/*.py num_classes).
#ifdef ALS_HEAD_ONLY_C  // bring-up builds: a single class count compiles in seconds
#define ALS_HEAD_C_LIST(X) X(ALS_HEAD_ONLY_C)
#else
#define ALS_HEAD_C_LIST(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) \
  X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)
#endif

template <int C>
static bool geometry_agrees(const HeadGeom& g) {
  using T = Geom<C>;
  return g.CB == T::CB && g.n[0] == T::N0 && g.n[1] == T::N1 && g.n[2] == T::N2 && g.n[3] == T::N3 && g.rows == T::ROWS &&
         g.col0[0] == 0 && g.col0[1] == 0 && g.col0[2] == T::CB && g.col0[3] == T::CB && g.row0[1] == T::N0 &&
         g.row0[2] == T::N0 + T::N1 && g.row0[3] == T::N0 + T::N1 + T::N2;
}

HeadPlan plan_head(int C, int measure, int T, int num_sms) {
  HeadPlan plan{};
  const HeadGeom g = head_geometry(C);
  bool same = false;
  if (T < 1 || measure < kEntropy || measure > kVariance || (measure == kVariance && T < 2)) return plan;
  switch (C) {
#define X(c)                                                       \
  case c:                                                          \
    plan.func = pick_head<c>(measure, T, &plan.name, &plan.block); \
    same = geometry_agrees<c>(g);                                  \
    break;
    ALS_HEAD_C_LIST(X)
#undef X
    default: plan.func = nullptr; break;
  }
  if (!plan.func) return plan;
  if (!same) {  // host packing and device geometry disagree: refuse rather than mis-compute
    plan.func = nullptr;
    return plan;
  }
  plan.grid = num_sms;
  plan.smem_bytes = kHeaderBytes + kRawAlign + kRawStages * kRawSlotBytes + 2 * 4 * g.rows * 16;
  return plan;
}

#ifdef ALS_HEAD_TRACE
extern "C" __attribute__((visibility("default"))) int als_debug_head_trace(long long* out, int tiles) {
  if (tiles > kTraceTiles) tiles = kTraceTiles;
  return cudaMemcpyFromSymbol(out, g_head_trace, sizeof(long long) * kTraceSlots * tiles) == cudaSuccess ? tiles : -1;
}
#endif

// The feature map as a 4-D tensor for the TMA engine: [T*N images][h rows][w pixels][16 floats], box = 4 floats x 129
// pixels of one row (one 16-byte chunk plane of a row segment); out-of-range coordinates read as zero.
static cudaError_t make_feature_map(const HeadParams& p, CUtensorMap* map) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) return cudaErrorNotSupported;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(kHeadChannels), static_cast<cuuint64_t>(p.w), static_cast<cuuint64_t>(p.h),
                              static_cast<cuuint64_t>(p.T) * static_cast<cuuint64_t>(p.n_images)};
  const cuuint64_t strides[3] = {kHeadChannels * sizeof(float), static_cast<cuuint64_t>(p.w) * kHeadChannels * sizeof(float),
                                 static_cast<cuuint64_t>(p.h) * p.w * kHeadChannels * sizeof(float)};
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(kHeadChannels), static_cast<cuuint32_t>(kSlotPx), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.features), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_head(const HeadPlan& plan, HeadParams p, cudaStream_t stream) {
  if (p.n_units <= 0) return cudaSuccess;
  p.sp.any_out = (p.sp.conf_map || p.sp.label || p.sp.mask) ? 1 : 0;
  p.sp.never = 0xffffffffu;
  cudaError_t err = cudaFuncSetAttribute(plan.func, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
  if (err != cudaSuccess) return err;
  const long long grid = p.n_units < plan.grid ? p.n_units : plan.grid;
  alignas(64) CUtensorMap feat_map;
  err = make_feature_map(p, &feat_map);
  if (err != cudaSuccess) return err;
  void* args[] = {&p, &feat_map};
  return cudaLaunchKernel(plan.func, dim3(static_cast<unsigned>(grid)), dim3(plan.block), args, plan.smem_bytes, stream);
}

}  // namespace als
