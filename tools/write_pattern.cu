// Write-pattern microbenchmark: 1.47 GB written as [rows][6144 B]; every CTA owns 128 consecutive rows and
// sweeps them in steps of S contiguous bytes per row (S = 512 is the packers' 4 lines per step).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_pattern tools/write_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr(float4* dst, long long stripes, int S16 /* 16-byte units per row per step */, int ROW16) {
  for (long long st = blockIdx.x; st < stripes; st += gridDim.x) {
    float4* base = dst + st * 128 * ROW16;
    for (int off = 0; off < ROW16; off += S16) {
      // 128 rows x S16 units; consecutive threads -> consecutive 16-byte units of a row
      for (int u = threadIdx.x; u < 128 * S16; u += blockDim.x) {
        const int r = u / S16, c = u - r * S16;
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(base + (long long)r * ROW16 + off + c), "f"((float)u) : "memory");
      }
    }
  }
}
int main() {
  const int ROW16 = 6144 / 16;
  const long long rows = 239616, stripes = rows / 128;
  float4* d; cudaMalloc(&d, rows * 6144);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int Ss[] = {128, 256, 512, 1024, 2048, 6144};
  int threads[] = {128, 256, 384};
  for (int th : threads) for (int S : Ss) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0); wr<<<148 * (th == 128 ? 3 : 1), th>>>(d, stripes, S / 16, ROW16); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("threads %d  S=%4d B contiguous per row per step: %.1f us  %.0f GB/s\n", th, S, best * 1e3, rows * 6144.0 / best / 1e6);
  }
  return 0;
}
