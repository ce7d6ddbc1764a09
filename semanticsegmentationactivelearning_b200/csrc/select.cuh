// Selection launchers (see select.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace als {

// ids at or above this value are padding (candidate slots of a rank that owns fewer than k unlabelled images)
constexpr long long kPadId = 1ll << 62;
// up to this many survivors are ordered inside the select kernel (one launch); above it a second, multi-CTA kernel ranks them
constexpr int kSelFusedMaxK = 1024;

// Where the (key, id) pairs of a selection come from.
struct SelectSrc {
  int mode;                  // 0: arrays   1: gather from the pool vector   2: gathered candidate records
  const float* keys;         // mode 0: keys[M]
  const long long* ids;      // mode 0: ids[M];  mode 1: unlabelled[M] (key = pool[id])
  const float* pool;         // mode 1: confidence vector (active_learning.py:685)
  long long lo, hi;          // mode 1: only ids in [lo, hi) take part (the images this rank owns)
  const unsigned char* rec;  // mode 2: pair i lives in record i / kc at slot i % kc
  long long rec_stride, keys_off, ids_off, kc;
};

// Score exchange folded into the merge launch (multi-GPU): record r = {int64 lo, int64 n, ...} + f32 scores[n] at
// scores_off; scores of every other rank are written into this rank's pool vector before the selection runs.
struct ScatterDesc {
  const unsigned char* rec;  // nullptr: nothing to scatter
  long long rec_stride, scores_off;
  int world, self;
  float* pool;
  long long pool_n;
};

struct SelectOut {
  long long* count;          // [2]: pairs written (= min(k, valid pairs)), status (0 ok, 1 shards do not partition the pool)
  float* keys;               // [k]  ascending in (key, id)
  long long* ids;            // [k]
  long long pad_base;        // >= 0: slots [count, k) become (NaN, pad_base + slot)
  float* uconf;              // optional: uconf[i] = pool[uconf_ids[i]], i < uconf_M   (unlabelled_confidence, :705)
  const long long* uconf_ids;
  long long uconf_M;
  const float* uconf_pool;
};

// This rank's half of the score exchange, folded into the local candidate selection: write the record header
// {int64 lo, int64 n} and the score slice pool[lo, lo + n) (zero padded to `width` floats) at scores_off.
struct ExportDesc {
  unsigned char* rec;  // nullptr: nothing to export
  long long scores_off, lo, n, width;
  const float* pool;
};

// k smallest valid (key, id) pairs of the M the source describes.  tmp_keys / tmp_ids: device scratch of min(k, M)
// entries, only touched when min(k, M) > kSelFusedMaxK.  Returns the number of kernels launched in *launches.
cudaError_t launch_select(const SelectSrc& src, long long M, long long k, const ScatterDesc& sc, const ExportDesc& ex,
                          const SelectOut& out, float* tmp_keys, long long* tmp_ids, cudaStream_t stream, int* launches);

}  // namespace als
