// Host-visible description of one scoring launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace als {

constexpr int kAccReplicas = 32;  // copies of the per-image accumulator vector (see score.cu)

struct ScoreParams {
  const void* logits;       // [T][N][P][C], class innermost (active_learning.py:231)
  long long sample_stride;  // elements between MC samples = N*P*C
  long long total_pixels;   // N*P
  long long P;              // pixels per image = H*W
  long long num_tiles;
  int T;
  int C;
  int measure;
  int stages;
  int any_out;  // conf_map || label || mask (set by launch_score)
  unsigned int never;  // 0xffffffff: the value a stage-release dependency word never takes (common.cuh; set by the launchers)
  int claim;    // most tiles a producer takes from the dynamic scheduler per atomic (set by the launchers; >= 1)
  int claim_shift;  // ceil(log2(2 * grid)): runs shrink once fewer than claim << claim_shift tiles are left (launchers)
  float inv_log2_c;  // 1 / log2(C): entropy in bits -> H / log(float32(C))   (active_learning.py:248-249)
  float threshold;  // alparams["threshold"]      (active_learning.py:265)
  float inv_T;
  float fx_scale;        // 2^fx_shift: per-pixel confidences are summed in Q(fx_shift) fixed point
  long long* acc;        // [kAccReplicas][acc_stride] fixed-point per-image sums (zero on entry; finalize re-zeroes)
  long long acc_stride;  // elements between replicas (>= N)
  unsigned int* flags;   // [N] bit0: a NaN confidence was seen
  unsigned long long* tile_counter;  // dynamic tile scheduler (zero on entry; finalize re-zeroes)
  float* conf_map;       // optional [N*P]
  uint8_t* label;        // optional [N*P]
  uint8_t* mask;         // optional [N*P]
  // Fused finalize (tiled kernel only, fin_n > 0): the last CTA to finish turns the fixed-point sums into scores
  // itself instead of a second launch -- see fused_finalize() in score.cu.  Same fields as launch_finalize.
  int fin_n;                      // images of this launch, 0 = a separate finalize_kernel follows
  unsigned int* done_counter;     // CTAs finished so far (zero on entry; the last CTA re-zeroes it)
  double fin_inv_scale_p;
  double* fin_scores64;
  float* fin_pool32;
  const long long* fin_example_index;
  long long fin_index_base;
  long long fin_num_examples;
};

// Largest launch (images) whose finalize is folded into the scoring kernel's last CTA; above it the separate,
// fully parallel finalize_kernel is cheaper than one CTA walking 32 accumulator replicas per image.
constexpr int kFusedFinalizeMaxImages = 2048;

// Longest run of consecutive tiles a CTA takes from the scheduler at once (produce_tiles, tiles.cuh).  A T > 1 tile
// is T stages already: 1.  T = 1, one ~20 KB stage per tile: f32 2 (reaches the same DRAM rate as T = 8 launches;
// longer runs only lengthen the tail of a batch-of-8 call), bf16 16 (XU- and issue-bound: the producer thread's
// per-tile instruction chain is what runs amortise; cfg3 went 0.865 -> 0.93 of the copy peak with runs of 16).
int max_claim(int T, int dtype);
int claim_shift_for(int grid);

struct LaunchPlan {
  const void* func;   // nullptr -> generic fallback
  const char* name;
  int grid, block, smem_bytes, stages, tile_pixels, lanes_per_pixel, pixels_per_thread, ctas_per_sm;
  bool tiled;
};

// Chooses the kernel for (dtype, C, measure, T); never fails (falls back to the generic kernel).
LaunchPlan plan_score(int dtype, int C, int measure, int T, long long total_pixels, bool aligned,
                      int num_sms, int max_smem_per_block);

cudaError_t launch_score(const LaunchPlan& plan, int dtype, ScoreParams p, cudaStream_t stream);

// Resident CTAs per SM of `func` with `smem_bytes` of dynamic shared memory (sets the opt-in attribute); cached.
int resident_ctas(const void* func, int block, int smem_bytes);

// scores64[i] = flags ? NaN : acc * 2^-shift / P ; optional f32 scatter pool32[example_index[i]] (example_index == nullptr:
// pool32[index_base + i]); re-zeroes acc/flags.
cudaError_t launch_finalize(long long* acc, long long acc_stride, unsigned int* flags, unsigned long long* tile_counter,
                            int n, double inv_scale_p,
                            double* scores64, float* pool32, const long long* example_index, long long index_base,
                            long long num_examples, cudaStream_t stream);

}  // namespace als
