// Streamed Monte-Carlo accumulation (mc.cu): one dropout sample per launch, Welford state resident in HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "score.cuh"

namespace als {

struct McPlan {
  bool tiled;             // false: generic fallback (any C, any alignment)
  const void* first;      // sample 0: no state read
  const void* update;     // sample t > 0: read state, fold the sample in, write state
  const void* finish;     // state -> confidence -> per-image sums (+ optional map / mask)
  const char* name;
  int grid, block, smem_bytes, stages, tile_pixels;
  int finish_grid, finish_block;
  long long state_floats;  // size of the state buffer for this (dtype, C, pixel count)
};

// Never fails: class counts without a specialised kernel, or a misaligned sample, run the generic kernels.
McPlan plan_mc(int dtype, int C, long long total_pixels, bool aligned, int num_sms, int max_smem_per_block);

// p: logits = this sample ([N,P,C]), T/sample_stride ignored; label (optional) is written by sample 0.
// p.tile_counter / p.done_counter must be zero on entry; the kernel's last CTA re-zeroes them.
cudaError_t launch_mc_update(const McPlan& plan, int dtype, ScoreParams p, float* state, int t, cudaStream_t stream);
// p: measure / inv_T / conf_map / mask / acc as for launch_score; the caller runs launch_finalize afterwards.
cudaError_t launch_mc_finish(const McPlan& plan, int dtype, ScoreParams p, const float* state, cudaStream_t stream);

}  // namespace als
