// Bring-up probe for the tcgen05 pieces the fused classifier-head kernel relies on (run on a B200):
//   1. canonical K-major no-swizzle operands written by ordinary threads (+ fence.proxy.async),
//   2. an A operand "shifted by one row" = the same buffer with start address + 16 bytes,
//   3. accumulating a narrower MMA into a column sub-range of an existing accumulator (D column offsets),
//   4. how kind::tf32 treats the 13 low mantissa bits of its fp32 inputs (truncate vs round).
// Prints one line per case; exit code 0 iff the layout cases are bit-exact against the host model.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../tc05.cuh"

using namespace als;

constexpr int kRowsA = 129;            // 128 MMA rows + 1 halo row
constexpr int kLboA = 130 * 16;        // bytes between 16-byte chunk planes of A
constexpr int kLboB = 129 * 16;        // same for B (N <= 128 rows)
constexpr int kK = 16;                 // channels = 4 chunk planes = 2 UMMA k-steps

struct ProbeArgs {
  const float* A;   // [129][16]
  const float* B0;  // [128][16]
  const float* B1;  // [n1][16]
  float* D;         // [128][128]
  int shift;        // A row shift of the second MMA (0 or 1)
  int dcol;         // first accumulator column of the second MMA
  int n1;           // N of the second MMA (multiple of 16, <= 128)
  int ts;           // 1: the second MMA takes A from tensor memory (written with tcgen05.st), not shared memory
};

__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs a) {
  __shared__ __align__(128) unsigned char sA[4 * kLboA];
  __shared__ __align__(128) unsigned char sB0[4 * kLboB];
  __shared__ __align__(128) unsigned char sB1[4 * kLboB];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kRowsA * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<float*>(sA + (k / 4) * kLboA + r * 16 + (k % 4) * 4) = a.A[i];
  }
  for (int i = tid; i < 128 * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<float*>(sB0 + (k / 4) * kLboB + r * 16 + (k % 4) * 4) = a.B0[i];
  }
  for (int i = tid; i < a.n1 * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<float*>(sB1 + (k / 4) * kLboB + r * 16 + (k % 4) * 4) = a.B1[i];
  }
  tc05::fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tc05::tmem_alloc<256>(&tslot);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t taddr = tslot;

  if (a.ts) {
    // A operand of the second MMA in tensor memory: lane = row, 16 consecutive 32-bit columns = K (cols 128..143)
    uint32_t r[16];
    for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(a.A[(tid + a.shift) * kK + k]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr + (static_cast<uint32_t>(32 * warp) << 16) + 128), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
        "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
  }

  if (tid == 0) {
    for (int ks = 0; ks < 2; ++ks)
      tc05::mma_tf32(taddr, tc05::smem_desc(smem_u32(sA) + 16 + ks * 2 * kLboA, kLboA, 128),
                     tc05::smem_desc(smem_u32(sB0) + ks * 2 * kLboB, kLboB, 128), tc05::idesc_tf32(128, 128), ks > 0);
    for (int ks = 0; ks < 2; ++ks) {
      if (a.ts)
        tc05::mma_tf32_ts(taddr + a.dcol, taddr + 128 + ks * 8, tc05::smem_desc(smem_u32(sB1) + ks * 2 * kLboB, kLboB, 128),
                          tc05::idesc_tf32(128, a.n1), true);
      else
        tc05::mma_tf32(taddr + a.dcol, tc05::smem_desc(smem_u32(sA) + 16 * a.shift + ks * 2 * kLboA, kLboA, 128),
                       tc05::smem_desc(smem_u32(sB1) + ks * 2 * kLboB, kLboB, 128), tc05::idesc_tf32(128, a.n1), true);
    }
    tc05::commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc05::fence_after_sync();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tc05::ld32(taddr + (static_cast<uint32_t>(32 * warp) << 16) + c0, v);
    tc05::ld_wait();
    for (int i = 0; i < 32; ++i) a.D[(32 * warp + lane) * 128 + c0 + i] = v[i];
  }
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc<256>(taddr);
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}
static float rna_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                  \
    }                                                                           \
  } while (0)

int main() {
  std::vector<float> A(kRowsA * kK), B0(128 * kK), B1(128 * kK), D(128 * 128);
  float *dA, *dB0, *dB1, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB0, B0.size() * 4));
  CK(cudaMalloc(&dB1, B1.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  int bad_layout = 0;

  auto run = [&](int shift, int dcol, int n1, int mode, double* err_out, int ts = 0) {
    // mode 0: values exact in tf32 (small integers / 8); 1: full 24-bit mantissas
    srand(1234 + shift * 7 + dcol * 13 + n1);
    auto rnd = [&]() {
      if (mode == 0) return static_cast<float>((rand() % 33) - 16) / 8.0f;
      return static_cast<float>(rand()) / RAND_MAX * 2.0f - 1.0f;
    };
    for (auto& v : A) v = rnd();
    for (auto& v : B0) v = rnd();
    for (auto& v : B1) v = rnd();
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB0, B0.data(), B0.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4));
    ProbeArgs a{dA, dB0, dB1, dD, shift, dcol, n1, ts};
    probe_kernel<<<1, 128>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    // host models: exact (double), truncated inputs, rounded inputs
    double err[3] = {0, 0, 0};
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double ref[3] = {0, 0, 0};
        for (int k = 0; k < kK; ++k) {
          const float x = A[(m + 1) * kK + k], w = B0[n * kK + k];
          ref[0] += static_cast<double>(x) * w;
          ref[1] += static_cast<double>(trunc_tf32(x)) * trunc_tf32(w);
          ref[2] += static_cast<double>(rna_tf32(x)) * rna_tf32(w);
        }
        if (n >= dcol && n < dcol + n1)
          for (int k = 0; k < kK; ++k) {
            const float x = A[(m + shift) * kK + k], w = B1[(n - dcol) * kK + k];
            ref[0] += static_cast<double>(x) * w;
            ref[1] += static_cast<double>(trunc_tf32(x)) * trunc_tf32(w);
            ref[2] += static_cast<double>(rna_tf32(x)) * rna_tf32(w);
          }
        for (int q = 0; q < 3; ++q) err[q] = fmax(err[q], fabs(ref[q] - D[m * 128 + n]));
      }
    for (int q = 0; q < 3; ++q) err_out[q] = err[q];
  };

  // (accumulator column offsets that are not a multiple of 4 fault with "misaligned address": dcol = 19 did)
  const int cases[][3] = {{1, 0, 128}, {0, 0, 128}, {0, 0, 64}, {1, 32, 32}, {0, 32, 64}, {0, 64, 64}, {0, 16, 32},
                          {0, 8, 32},  {0, 24, 48}, {0, 20, 48}, {1, 20, 32}, {0, 4, 16}};
  {
    double e[3];
    run(1, 32, 32, 1, e);
    printf("rounding  err vs exact=%.3e  vs truncated inputs=%.3e  vs rna-rounded inputs=%.3e  -> hardware %s\n", e[0], e[1],
           e[2], e[1] < e[2] ? "TRUNCATES" : "ROUNDS");
  }
  for (auto& c : cases) {
    double e[3];
    run(c[0], c[1], c[2], 0, e);
    const bool ok = e[0] == 0.0;
    printf("layout  shift=%d dcol=%3d n1=%3d  max|err|=%.3e  %s\n", c[0], c[1], c[2], e[0], ok ? "EXACT" : "MISMATCH");
    if (!ok && c[1] % 32 == 0) bad_layout = 1;
  }
  for (auto& c : cases) {
    double e[3];
    run(c[0], c[1], c[2], 0, e, 1);
    const bool ok = e[0] == 0.0;
    printf("A-in-TMEM  shift=%d dcol=%3d n1=%3d  max|err|=%.3e  %s\n", c[0], c[1], c[2], e[0], ok ? "EXACT" : "MISMATCH");
    if (!ok && c[1] % 32 == 0) bad_layout = 1;
  }
  printf(bad_layout ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad_layout;
}
