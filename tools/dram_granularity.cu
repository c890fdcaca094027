// Microbenchmark: DRAM bytes fetched per random aligned read of S bytes (S = 16..256) from a 4 GiB
// buffer.  Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum`.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int S, int MODE>  // MODE 0: ld.global.nc.L1::no_allocate  1: plain ld.global  2: ld.global.cg
__global__ void rnd_read(const float* __restrict__ buf, size_t n_units, float* out, int per_thread) {
  uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  float acc = 0.f;
  constexpr int V = S / 16;  // float4 loads per unit (S >= 16)
  for (int it = 0; it < per_thread; ++it) {
    r ^= r << 13; r ^= r >> 7; r ^= r << 17;
    const size_t unit = r % n_units;
    const float* p = buf + unit * (S / 4);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float4 q;
      if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p + 4 * v));
      else if (MODE == 1) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p + 4 * v));
      else asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p + 4 * v));
      acc += q.x + q.y + q.z + q.w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int S, int MODE>
void run(const float* buf, size_t bytes, float* out) {
  const int blocks = 148 * 8, threads = 256, per = 64;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  rnd_read<S, MODE><<<blocks, threads>>>(buf, bytes / S, out, per);
  cudaEventRecord(a);
  rnd_read<S, MODE><<<blocks, threads>>>(buf, bytes / S, out, per);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double useful = (double)blocks * threads * per * S;
  printf("S=%3d mode=%d: %8.1f us  useful %7.1f GB/s  (%.0f M accesses)\n", S, MODE, ms * 1e3, useful / ms / 1e6, blocks * threads * per / 1e6);
}

int main() {
  const size_t bytes = 4ull << 30;
  float *buf, *out;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
  cudaMemset(buf, 0, bytes);
  run<16, 0>(buf, bytes, out);
  run<32, 0>(buf, bytes, out);
  run<64, 0>(buf, bytes, out);
  run<128, 0>(buf, bytes, out);
  run<256, 0>(buf, bytes, out);
  run<32, 1>(buf, bytes, out);
  run<32, 2>(buf, bytes, out);
  run<64, 1>(buf, bytes, out);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
