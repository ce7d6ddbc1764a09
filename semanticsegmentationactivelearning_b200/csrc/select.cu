// Block-radix selection of the k smallest (key, id) pairs -- the device replacement for
// np.argpartition in /root/reference/active_learning.py:705-714.
//
// Total order (the reference leaves ties/NaN to NumPy's introselect): keys compare as floats
// with -0.0 == +0.0, NaN after +inf; equal keys are ordered by the smaller id.  Because ids are
// unique the k-smallest set is unique, so the result does not depend on thread scheduling.
//
// Pools are small (<= a few 1e5 images), so one 1024-thread CTA does an MSB-first radix select
// over the 96-bit virtual key (orderable(key):32 | biased id:64), 8 bits per pass, histogram in
// shared memory; a second, multi-CTA kernel puts the k survivors in ascending order by ranking.
// The select stops as soon as a digit bin holds exactly the elements still wanted (with distinct scores that is
// after the four key passes at the latest: the eight id passes only run when the k-th boundary falls inside a
// group of equal keys).
#include "select.cuh"

#include "common.cuh"

namespace als {

__device__ __forceinline__ uint32_t orderable(float f) {
  if (f != f) return 0xffffffffu;  // NaN last (NumPy sorts NaN to the end)
  const uint32_t u = __float_as_uint(f + 0.0f);  // -0.0 + 0.0 == +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long biased(long long id) {
  return static_cast<unsigned long long>(id) ^ 0x8000000000000000ull;
}
__device__ __forceinline__ bool pair_less(uint32_t ao, unsigned long long ai, uint32_t bo, unsigned long long bi) {
  return ao < bo || (ao == bo && ai < bi);
}

constexpr int kSelThreads = 1024;

// keys[i] (or scores[ids[i]] when gather_from != nullptr) / ids[i], i < M.
__global__ void __launch_bounds__(kSelThreads) select_threshold_kernel(const float* __restrict__ keys,
                                                                       const long long* __restrict__ ids, long long M,
                                                                       long long k, float* __restrict__ tmp_keys,
                                                                       long long* __restrict__ tmp_ids) {
  __shared__ unsigned int hist[256];
  __shared__ uint32_t s_ord;
  __shared__ unsigned long long s_id;
  __shared__ long long s_remaining;
  __shared__ unsigned int s_count;
  __shared__ int s_done;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_ord = 0;
    s_id = 0;
    s_remaining = k;
    s_count = 0;
    s_done = 0;
  }
  __syncthreads();
  if (k >= M) {
    if (tid == 0) { s_ord = 0xffffffffu; s_id = ~0ull; }
  } else {
    for (int pass = 0; pass < 12; ++pass) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const uint32_t pord = s_ord;
      const unsigned long long pid = s_id;
      for (long long i = tid; i < M; i += kSelThreads) {
        const uint32_t o = orderable(keys[i]);
        unsigned int digit;
        bool match;
        if (pass < 4) {
          const int sh = 32 - 8 * pass;  // bits already fixed: the top 8*pass
          match = (pass == 0) || ((o >> sh) == (pord >> sh));
          digit = (o >> (24 - 8 * pass)) & 0xffu;
        } else {
          const unsigned long long b = biased(ids[i]);
          const int q = pass - 4;
          const int sh = 64 - 8 * q;
          match = (o == pord) && (q == 0 || ((b >> sh) == (pid >> sh)));
          digit = static_cast<unsigned int>((b >> (56 - 8 * q)) & 0xffull);
        }
        if (match) atomicAdd(&hist[digit], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        long long rem = s_remaining;
        unsigned int d = 0;
        for (; d < 255; ++d) {
          if (rem <= static_cast<long long>(hist[d])) break;
          rem -= hist[d];
        }
        s_remaining = rem;
        if (pass < 4) s_ord |= d << (24 - 8 * pass);
        else s_id |= static_cast<unsigned long long>(d) << (56 - 8 * (pass - 4));
        if (rem == static_cast<long long>(hist[d])) {
          // every element with this prefix is wanted: the threshold is the largest virtual key with the prefix
          if (pass < 4) {
            s_ord |= (pass == 3) ? 0u : (0xffffffffu >> (8 * (pass + 1)));
            s_id = ~0ull;
          } else {
            s_id |= (pass == 11) ? 0ull : (~0ull >> (8 * (pass - 3)));
          }
          s_done = 1;
        }
      }
      __syncthreads();
      if (s_done) break;
    }
  }
  __syncthreads();
  const uint32_t tord = s_ord;
  const unsigned long long tidb = s_id;
  for (long long i = tid; i < M; i += kSelThreads) {
    const float key = keys[i];
    const long long id = ids[i];
    const uint32_t o = orderable(key);
    const unsigned long long b = biased(id);
    if (!pair_less(tord, tidb, o, b)) {  // (o, b) <= threshold
      const unsigned int pos = atomicAdd(&s_count, 1u);
      tmp_keys[pos] = key;
      tmp_ids[pos] = id;
    }
  }
}

// out[rank(j)] = tmp[j], rank = number of survivors strictly smaller.
__global__ void __launch_bounds__(256) rank_scatter_kernel(const float* __restrict__ tmp_keys,
                                                           const long long* __restrict__ tmp_ids, long long n,
                                                           float* __restrict__ out_keys, long long* __restrict__ out_ids) {
  __shared__ uint32_t so[256];
  __shared__ unsigned long long sb[256];
  const long long j = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  float key = 0.f;
  long long id = 0;
  uint32_t o = 0;
  unsigned long long b = 0;
  if (j < n) {
    key = tmp_keys[j];
    id = tmp_ids[j];
    o = orderable(key);
    b = biased(id);
  }
  long long rank = 0;
  for (long long base = 0; base < n; base += 256) {
    const long long i = base + threadIdx.x;
    if (i < n) {
      so[threadIdx.x] = orderable(tmp_keys[i]);
      sb[threadIdx.x] = biased(tmp_ids[i]);
    }
    __syncthreads();
    const int lim = static_cast<int>((n - base) < 256 ? (n - base) : 256);
    for (int q = 0; q < lim; ++q) rank += pair_less(so[q], sb[q], o, b) ? 1 : 0;
    __syncthreads();
  }
  if (j < n) {
    out_keys[rank] = key;
    out_ids[rank] = id;
  }
}

__global__ void gather_scores_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, long long M,
                                     float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < M) out[i] = scores[ids[i]];  // unlabelled_confidence = confidence[unlabelled]  (:705)
}

cudaError_t launch_gather(const float* scores, const long long* ids, long long M, float* out, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  gather_scores_kernel<<<static_cast<unsigned int>((M + 255) / 256), 256, 0, stream>>>(scores, ids, M, out);
  return cudaGetLastError();
}

cudaError_t launch_select(const float* keys, const long long* ids, long long M, long long k, float* tmp_keys,
                          long long* tmp_ids, float* out_keys, long long* out_ids, cudaStream_t stream) {
  const long long kk = k < M ? k : M;
  if (kk <= 0) return cudaSuccess;
  select_threshold_kernel<<<1, kSelThreads, 0, stream>>>(keys, ids, M, kk, tmp_keys, tmp_ids);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  rank_scatter_kernel<<<static_cast<unsigned int>((kk + 255) / 256), 256, 0, stream>>>(tmp_keys, tmp_ids, kk, out_keys,
                                                                                      out_ids);
  return cudaGetLastError();
}

}  // namespace als
