// Selection launchers (see select.cu).
#pragma once
#include <cuda_runtime.h>

namespace als {

// out[i] = scores[ids[i]]
cudaError_t launch_gather(const float* scores, const long long* ids, long long M, float* out, cudaStream_t stream);

// k smallest (key, id) pairs of M, ascending, into out_keys/out_ids[min(k, M)].
// tmp_keys/tmp_ids: device scratch of min(k, M) entries.
cudaError_t launch_select(const float* keys, const long long* ids, long long M, long long k, float* tmp_keys,
                          long long* tmp_ids, float* out_keys, long long* out_ids, cudaStream_t stream);

}  // namespace als
