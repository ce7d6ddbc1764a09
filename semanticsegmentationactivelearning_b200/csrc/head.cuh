// Fused ENet classifier head + pool scoring (head.cu): host-visible launch description.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "score.cuh"

namespace als {

constexpr int kHeadChannels = 16;   // input channels of `Final` (models/enet/enet_modules.py:1341)
constexpr int kHeadTileQuads = 128; // input pixels per MMA tile (UMMA M)
constexpr int kHeadMaxClasses = 32;   // fused kernels exist for 2 <= C <= 32 ...
constexpr int kHeadMaxClassesMC = 24; // ... and, with T > 1 Monte-Carlo samples, for 2 <= C <= 24

// Column layout of one accumulator tile: 4 blocks of CB = round_up(C, 4) columns, one per output pixel of
// the 2x2 quad an input pixel produces:  block 0 -> (dy,dx) = (0,1), 1 -> (0,0), 2 -> (1,0), 3 -> (1,1).
// Operand o multiplies source pixel (i - oy, j - ox): o = 0 -> (0,0), 1 -> (1,0), 2 -> (0,1), 3 -> (1,1).
struct HeadGeom {
  int C, CB;
  int n[4];      // UMMA N of operand o (multiple of 16)
  int col0[4];   // first accumulator column operand o adds into
  int row0[4];   // first row of operand o inside the packed B image
  int rows;      // total rows of the packed B image
};
HeadGeom head_geometry(int C);

// Packs the transposed-convolution kernel [3][3][C][16] (TF filter layout, enet_modules.py:1341) into the
// canonical K-major UMMA operand image, split into tf32 hi and lo parts:
//   out[part][chunk plane c (4)][row (geom.rows)][4 floats], part 0 = hi, 1 = lo.
// Returns the number of floats written (2 * 4 * rows * 4).
size_t pack_head_weights(const float* kernel, int C, float* out);

struct HeadParams {
  ScoreParams sp;          // P = H*W of the OUTPUT (2h x 2w); acc / flags / outputs / fx_scale as in score.cu
  const float* features;   // [T][N][h][w][16] fp32, 16-byte aligned (read through a TMA tensor map: sample t, image n = tensor row t*N + n)
  int T;                   // samples (1 = the reference's single forward pass)
  const float* weights;    // packed B image (pack_head_weights), device
  int h, w;                // input (feature) height / width; output is 2h x 2w
  int n_images;
  int n_strips;            // ceil(w / 128)
  int rows_per_unit;       // R
  int n_rowblocks;         // ceil(h / R)
  long long n_units;       // N * n_rowblocks * n_strips
  HeadGeom g;
};

struct HeadPlan {
  const void* func;
  const char* name;
  int grid, block, smem_bytes;
};

// nullptr func when (C, measure, T) has no fused-head instantiation.
HeadPlan plan_head(int C, int measure, int T, int num_sms);
cudaError_t launch_head(const HeadPlan& plan, HeadParams p, cudaStream_t stream);

}  // namespace als
