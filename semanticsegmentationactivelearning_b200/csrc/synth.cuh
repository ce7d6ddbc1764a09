// Synthetic logits generator launcher (see synth.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace als {
cudaError_t launch_synth(void* out, int dtype, long long T, long long n0, long long n_imgs, long long P, int C,
                         uint64_t seed, int mc, cudaStream_t stream);
cudaError_t launch_fill(void* buf, size_t bytes, cudaStream_t stream);
}  // namespace als
