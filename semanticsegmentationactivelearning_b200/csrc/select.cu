// Block-radix selection of the k smallest (key, id) pairs -- the device replacement for
// np.argpartition in /root/reference/active_learning.py:705-714.
//
// Total order (the reference leaves ties/NaN to NumPy's introselect): keys compare as floats
// with -0.0 == +0.0, NaN after +inf; equal keys are ordered by the smaller id.  Because ids are
// unique the k-smallest set is unique, so the result does not depend on thread scheduling.
//
// Pools are small (<= a few 1e5 images), so ONE 1024-thread CTA does the whole of :705-714 in one launch:
//   gather   key = confidence[unlabelled[i]]                       (:705, also written out as unlabelled_confidence)
//   select   MSB-first radix select over the 96-bit virtual key (orderable(key):32 | biased id:64), 8 bits per pass,
//            shared-memory histogram, the 256 bins scanned by one warp (8 bins per lane + a shuffle scan);
//            it stops as soon as a digit bin holds exactly the elements still wanted (with distinct scores that is
//            after the four key passes at the latest: the eight id passes only run when the k-th boundary falls
//            inside a group of equal keys)
//   order    the <= 1024 survivors are ranked against each other in shared memory and written in ascending order
// (more than 1024 survivors -- "select everything" calls -- go through a second, multi-CTA ranking kernel).
// The same kernel serves the multi-GPU path: restricted to the ids a rank owns it produces that rank's candidates
// (and exports its score slice into the exchange record); fed the all-gathered records it scatters the other ranks'
// scores into the pool vector and merges the candidates.
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
constexpr int kSelItems = 4;  // pairs per thread kept in registers across the radix passes

// pair i of the source; false = not part of this selection (not owned by this rank / padding)
__device__ __forceinline__ bool load_pair(const SelectSrc& s, long long i, float& key, long long& id) {
  if (s.mode == 0) {
    key = s.keys[i];
    id = s.ids[i];
    return true;
  }
  if (s.mode == 1) {
    id = s.ids[i];
    if (id < s.lo || id >= s.hi) return false;
    key = s.pool[id];  // unlabelled_confidence = confidence[unlabelled]  (:705)
    return true;
  }
  const long long r = i / s.kc, j = i - r * s.kc;
  const unsigned char* rec = s.rec + r * s.rec_stride;
  id = reinterpret_cast<const long long*>(rec + s.ids_off)[j];
  if (id >= kPadId) return false;
  key = reinterpret_cast<const float*>(rec + s.keys_off)[j];
  return true;
}

// sort_here: order the survivors in shared memory and write them to out.keys / out.ids (needs min(k, M) <= 1024);
// otherwise they go to tmp_keys / tmp_ids unordered and rank_scatter_kernel finishes the job.
__global__ void __launch_bounds__(kSelThreads) select_kernel(const SelectSrc s, const long long M, const long long k,
                                                             const ScatterDesc sc, const ExportDesc ex, const SelectOut out,
                                                             float* __restrict__ tmp_keys, long long* __restrict__ tmp_ids,
                                                             const int sort_here) {
  __shared__ unsigned int hist[256];
  __shared__ uint32_t s_ord;
  __shared__ unsigned long long s_id;
  __shared__ long long s_remaining;
  __shared__ unsigned int s_count;
  __shared__ int s_done;
  __shared__ long long s_warp[kSelThreads / 32];
  __shared__ long long s_valid;
  __shared__ uint32_t so[kSelFusedMaxK];
  __shared__ unsigned long long sb[kSelFusedMaxK];
  __shared__ float sk[kSelFusedMaxK];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  if (tid == 0) {
    s_ord = 0;
    s_id = 0;
    s_count = 0;
    s_done = 0;
  }
  pdl_launch_dependents();
  pdl_wait();  // the pool vector is written by the finalize launches before us in the stream

  // ---- multi-GPU, merge side: the other ranks' score slices complete this rank's pool vector ----
  if (sc.rec) {
    for (int r = 0; r < sc.world; ++r) {
      if (r == sc.self) continue;
      const unsigned char* rec = sc.rec + r * sc.rec_stride;
      const long long lo = reinterpret_cast<const long long*>(rec)[0];
      const long long n = reinterpret_cast<const long long*>(rec)[1];
      if (lo < 0 || n < 0 || lo + n > sc.pool_n) continue;  // reported through the status word below
      const float* src = reinterpret_cast<const float*>(rec + sc.scores_off);
      for (long long i = tid; i < n; i += kSelThreads) sc.pool[lo + i] = src[i];
    }
    if (tid == 0) {  // the shards must tile [0, pool_n): every example has exactly one owner
      long long total = 0;
      int bad = 0;
      for (int r = 0; r < sc.world; ++r) {
        const long long* h = reinterpret_cast<const long long*>(sc.rec + r * sc.rec_stride);
        if (h[0] < 0 || h[1] < 0 || h[0] + h[1] > sc.pool_n) bad = 1;
        total += h[1];
        for (int q = 0; q < r; ++q) {
          const long long* g = reinterpret_cast<const long long*>(sc.rec + q * sc.rec_stride);
          if (h[1] > 0 && g[1] > 0 && h[0] < g[0] + g[1] && g[0] < h[0] + h[1]) bad = 1;
        }
      }
      if (total != sc.pool_n) bad = 1;
      out.count[1] = bad;
    }
  } else if (tid == 0 && out.count) {
    out.count[1] = 0;
  }
  // ---- multi-GPU, local side: export this rank's score slice next to its candidates ----
  if (ex.rec) {
    if (tid == 0) {
      reinterpret_cast<long long*>(ex.rec)[0] = ex.lo;
      reinterpret_cast<long long*>(ex.rec)[1] = ex.n;
    }
    float* dst = reinterpret_cast<float*>(ex.rec + ex.scores_off);
    for (long long i = tid; i < ex.width; i += kSelThreads) dst[i] = i < ex.n ? ex.pool[ex.lo + i] : 0.f;
  }
  __syncthreads();  // scattered scores are visible to the whole CTA from here on

  // ---- unlabelled_confidence = confidence[unlabelled]  (:705) ----
  if (out.uconf)
    for (long long i = tid; i < out.uconf_M; i += kSelThreads) out.uconf[i] = out.uconf_pool[out.uconf_ids[i]];

  // ---- the pairs: up to kSelItems per thread stay in registers for all passes (pools up to 4096 images -- every
  //      BASELINE configuration); larger inputs are re-read from global memory (L2 resident) in every pass ----
  const bool cached = M <= static_cast<long long>(kSelItems) * kSelThreads;
  uint32_t co[kSelItems];
  unsigned long long cb[kSelItems];
  float ck[kSelItems];
  bool cv[kSelItems];
  long long mine = 0;
  if (cached) {
#pragma unroll
    for (int q = 0; q < kSelItems; ++q) {
      const long long i = tid + static_cast<long long>(q) * kSelThreads;
      float key = 0.f;
      long long id = 0;
      cv[q] = i < M && load_pair(s, i, key, id);
      ck[q] = key;
      co[q] = orderable(key);
      cb[q] = biased(id);
      mine += cv[q] ? 1 : 0;
    }
  } else {
    float key;
    long long id;
    for (long long i = tid; i < M; i += kSelThreads) mine += load_pair(s, i, key, id) ? 1 : 0;
  }
  // fn(orderable key, biased id, key) for every pair of this thread that takes part
  auto visit = [&](auto&& fn) {
    if (cached) {
#pragma unroll
      for (int q = 0; q < kSelItems; ++q)
        if (cv[q]) fn(co[q], cb[q], ck[q]);
    } else {
      for (long long i = tid; i < M; i += kSelThreads) {
        float key;
        long long id;
        if (load_pair(s, i, key, id)) fn(orderable(key), biased(id), key);
      }
    }
  };
  mine = warp_sum_ll(mine);
  if (lane == 0) s_warp[tid >> 5] = mine;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) t += s_warp[w];
    s_valid = t;
    s_remaining = k < t ? k : t;
  }
  __syncthreads();
  const long long nvalid = s_valid;
  const long long kk = k < nvalid ? k : nvalid;  // :707-708  min(len(unlabelled), selection_size)

  if (kk >= nvalid || kk == 0) {
    if (tid == 0) { s_ord = 0xffffffffu; s_id = ~0ull; }  // everything valid is wanted (or nothing: skipped below)
  } else {
    for (int pass = 0; pass < 12; ++pass) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const uint32_t pord = s_ord;
      const unsigned long long pid = s_id;
      visit([&](uint32_t o, unsigned long long b, float) {
        unsigned int digit;
        bool match;
        if (pass < 4) {
          const int sh = 32 - 8 * pass;  // bits already fixed: the top 8*pass
          match = (pass == 0) || ((o >> sh) == (pord >> sh));
          digit = (o >> (24 - 8 * pass)) & 0xffu;
        } else {
          const int q = pass - 4;
          const int sh = 64 - 8 * q;
          match = (o == pord) && (q == 0 || ((b >> sh) == (pid >> sh)));
          digit = static_cast<unsigned int>((b >> (56 - 8 * q)) & 0xffull);
        }
        if (match) atomicAdd(&hist[digit], 1u);
      });
      __syncthreads();
      if (tid < 32) {
        // bin scan: lane l owns bins 8l .. 8l+7; the bin where the running count reaches `remaining` is the digit
        unsigned int c[8];
        unsigned int sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          c[q] = hist[8 * lane + q];
          sum += c[q];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const long long rem = s_remaining;
        const long long excl = static_cast<long long>(incl) - sum;
        if (rem > excl && rem <= static_cast<long long>(incl)) {  // exactly one lane (the matching set holds >= rem pairs)
          long long r = rem - excl;
          unsigned int d = 8 * lane;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (r <= static_cast<long long>(c[q])) { d = 8 * lane + q; break; }
            r -= c[q];
          }
          const unsigned int cd = hist[d];
          s_remaining = r;
          if (pass < 4) s_ord = pord | (d << (24 - 8 * pass));
          else s_id = pid | (static_cast<unsigned long long>(d) << (56 - 8 * (pass - 4)));
          if (r == static_cast<long long>(cd)) {
            // every pair with this prefix is wanted: the threshold is the largest virtual key with the prefix
            if (pass < 4) {
              s_ord |= (pass == 3) ? 0u : (0xffffffffu >> (8 * (pass + 1)));
              s_id = ~0ull;
            } else {
              s_id |= (pass == 11) ? 0ull : (~0ull >> (8 * (pass - 3)));
            }
            s_done = 1;
          }
        }
      }
      __syncthreads();
      if (s_done) break;
    }
  }
  __syncthreads();
  const uint32_t tord = s_ord;
  const unsigned long long tidb = s_id;
  const unsigned int cap = static_cast<unsigned int>(kk);  // unique ids give exactly kk survivors; duplicates must not overflow
  if (kk > 0)
    visit([&](uint32_t o, unsigned long long b, float key) {
      if (!pair_less(tord, tidb, o, b)) {  // (o, b) <= threshold
        const unsigned int pos = atomicAdd(&s_count, 1u);
        if (pos < cap) {
          if (sort_here) {
            so[pos] = o;
            sb[pos] = b;
            sk[pos] = key;
          } else {
            tmp_keys[pos] = key;
            tmp_ids[pos] = static_cast<long long>(b ^ 0x8000000000000000ull);
          }
        }
      }
    });
  __syncthreads();
  const long long n = s_count < cap ? s_count : cap;
  if (tid == 0 && out.count) out.count[0] = n;
  if (sort_here) {
    // out[rank(j)] = survivor j, rank = number of survivors strictly smaller
    if (tid < n) {
      const uint32_t o = so[tid];
      const unsigned long long b = sb[tid];
      int rank = 0;
      for (int q = 0; q < n; ++q) rank += pair_less(so[q], sb[q], o, b) ? 1 : 0;
      out.keys[rank] = sk[tid];
      out.ids[rank] = static_cast<long long>(b ^ 0x8000000000000000ull);
    }
  }
  if (out.pad_base >= 0)
    for (long long j = n + tid; j < k; j += kSelThreads) {
      out.keys[j] = __int_as_float(0x7fc00000);
      out.ids[j] = out.pad_base + j;
    }
}

// out[rank(j)] = tmp[j], rank = number of survivors strictly smaller; *count survivors (large selections only).
__global__ void __launch_bounds__(256) rank_scatter_kernel(const float* __restrict__ tmp_keys,
                                                           const long long* __restrict__ tmp_ids,
                                                           const long long* __restrict__ count,
                                                           float* __restrict__ out_keys, long long* __restrict__ out_ids) {
  __shared__ uint32_t so[256];
  __shared__ unsigned long long sb[256];
  const long long n = *count;
  const long long j = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (static_cast<long long>(blockIdx.x) * 256 >= n) return;
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

cudaError_t launch_select(const SelectSrc& src, long long M, long long k, const ScatterDesc& sc, const ExportDesc& ex,
                          const SelectOut& out, float* tmp_keys, long long* tmp_ids, cudaStream_t stream, int* launches) {
  if (launches) *launches = 0;
  if (M < 0 || k < 0) return cudaErrorInvalidValue;
  const long long kmax = k < M ? k : M;
  int sort_here = kmax <= kSelFusedMaxK ? 1 : 0;
  SelectSrc s = src;
  long long m = M, kk = k;
  ScatterDesc scd = sc;
  ExportDesc exd = ex;
  SelectOut o = out;
  void* args[] = {&s, &m, &kk, &scd, &exd, &o, &tmp_keys, &tmp_ids, &sort_here};
  cudaError_t err = launch_pdl((const void*)select_kernel, dim3(1), dim3(kSelThreads), args, 0, stream);
  if (err != cudaSuccess) return err;
  if (launches) *launches = 1;
  if (!sort_here) {
    rank_scatter_kernel<<<static_cast<unsigned int>((kmax + 255) / 256), 256, 0, stream>>>(tmp_keys, tmp_ids, out.count,
                                                                                        out.keys, out.ids);
    if (launches) *launches = 2;
    return cudaGetLastError();
  }
  return cudaSuccess;
}

}  // namespace als
