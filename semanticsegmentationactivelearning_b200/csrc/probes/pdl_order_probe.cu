// Bring-up probe (not part of the library): does a NON-KERNEL stream operation (memset / memcpy / event wait) that sits
// between two kernels keep its place when the second kernel is launched with programmatic stream serialization (PDL)
// and the first one calls griddepcontrol.launch_dependents early?
//
//   A (long, triggers its dependents at once)  ->  op writes `word`  ->  B (PDL attribute; griddepcontrol.wait, then reads `word`)
//
// Stream order says B must see the value the op wrote.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_probe pdl_order_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void kernel_a(unsigned int* word, unsigned int value, long long spin) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *word = value;  // A leaves its own mark; the op overwrites it
}

__global__ void kernel_b(const unsigned int* word, unsigned int* seen, int i, int do_wait) {
  if (do_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) seen[i] = *reinterpret_cast<const volatile unsigned int*>(word);
}

static void launch_b(cudaStream_t s, const unsigned int* word, unsigned int* seen, int i, int pdl, int do_wait) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(32);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  void* args[] = {(void*)&word, (void*)&seen, (void*)&i, (void*)&do_wait};
  cudaLaunchKernelExC(&cfg, (const void*)kernel_b, args);
}

int main() {
  const int N = 400;
  unsigned int *word, *seen, *host_src;
  cudaMalloc(&word, 64);
  cudaMalloc(&seen, N * sizeof(unsigned int));
  cudaMallocHost(&host_src, 64);
  cudaStream_t s, s2;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t ev;
  cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  static unsigned int out[N];
  const char* names[] = {"memset", "memcpy H2D (pinned)", "event wait on another stream's memset", "nothing (A's own write)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int pdl = 0; pdl < 2; ++pdl) {
      cudaMemset(seen, 0, N * sizeof(unsigned int));
      cudaDeviceSynchronize();
      for (int i = 0; i < N; ++i) {
        const unsigned int a_mark = 0x10000000u + i, op_mark = mode == 0 ? 0x5a5a5a5au : 0x20000000u + i;
        kernel_a<<<1, 32, 0, s>>>(word, a_mark, 60000);
        if (mode == 0) cudaMemsetAsync(word, 0x5a, 4, s);
        if (mode == 1) { *host_src = op_mark; cudaMemcpyAsync(word, host_src, 4, cudaMemcpyHostToDevice, s); }
        if (mode == 2) {
          // the other stream waits for A, writes, and our stream waits for that write
          cudaEventRecord(ev, s);
          cudaStreamWaitEvent(s2, ev, 0);
          *host_src = op_mark;
          cudaMemcpyAsync(word, host_src, 4, cudaMemcpyHostToDevice, s2);
          cudaEventRecord(ev, s2);
          cudaStreamWaitEvent(s, ev, 0);
        }
        launch_b(s, word, seen, i, pdl, 1);
        if (mode == 1 || mode == 2) cudaStreamSynchronize(s);  // host_src is reused
      }
      cudaDeviceSynchronize();
      cudaMemcpy(out, seen, N * sizeof(unsigned int), cudaMemcpyDeviceToHost);
      int wrong = 0;
      for (int i = 0; i < N; ++i) {
        const unsigned int want = mode == 0 ? 0x5a5a5a5au : (mode == 3 ? 0x10000000u + i : 0x20000000u + i);
        wrong += out[i] != want;
      }
      printf("op between A and B: %-42s B launched %s: %d of %d reads saw a stale value%s\n", names[mode],
             pdl ? "with PDL   " : "normally   ", wrong, N, cudaGetLastError() == cudaSuccess ? "" : "  (CUDA error)");
    }
  return 0;
}
