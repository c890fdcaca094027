// What a persistent CTA per SM can push through the TMA store path - the output path of corr_pack_tf32 on its own:
// 239 616 rows x 49 lines x 128 B = 1.50 GB written as in csrc/corr_pack_tcgen05.cu (stripes of 128 rows, 13 steps of
// four [128 rows][128 B] tiles, one TMA store per tile through a map {32 floats, line, row}).
//   mode 0  TMA stores only (tiles never rewritten): issue + wait_group.read per step, no barrier
//   mode 1  + the epilogue's staging: every thread rewrites its row of the four tiles (32 x st.shared.v4), fence,
//           two CTA barriers per step, single-buffered (the shipped structure)
//   mode 2  the same with two sets of four tiles (128 KB): the stores of step k drain while step k+1 is staged
//   mode 3  mode 1 with plain coalesced st.global.v4 from the staging tiles instead of TMA
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_store_rate tools/tma_store_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128) wr(const __grid_constant__ CUtensorMap map, float* dst, int stripes, int nblk, int steps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  uint32_t it = 0;
  for (int st = blockIdx.x; st < stripes; st += gridDim.x) {
    const int row0 = st * 128;
    for (int cc = 0; cc < steps; ++cc, ++it) {
      uint8_t* buf = smem + (MODE == 2 ? (it & 1) * 65536 : 0);
      if (MODE != 0) {
        if (tid == 0) {
          if (MODE == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else if (MODE == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const float v = (float)(cc + tid);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(buf + j * 16384 + tid * 128 + ((k ^ (tid & 7)) << 4)) = make_float4(v, v + 1, v + 2, v + 3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (MODE == 3) {
        // 4 tiles x 128 rows x 8 chunks: thread t copies chunk (t & 7) of rows (t >> 3) + 16 n
        for (int j = 0; j < 4; ++j) {
          if (4 * cc + j >= nblk) break;
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const int r = (tid >> 3) + 16 * n, k = tid & 7;
            const float4 x = *reinterpret_cast<const float4*>(buf + j * 16384 + r * 128 + ((k ^ (r & 7)) << 4));
            float* p = dst + ((size_t)(row0 + r) * nblk + 4 * cc + j) * 32 + k * 4;
            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
          }
        }
      } else if (tid == 0) {
        for (int j = 0; j < 4; ++j)
          if (4 * cc + j < nblk) tma_store_3d(&map, buf + j * 16384, 0, 4 * cc + j, row0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (MODE == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
static void run(const CUtensorMap& map, float* d, int stripes, int nblk, int steps, int ctas_per_sm, double bytes, const char* what) {
  const int smem = MODE == 2 ? 131072 : 65536;
  cudaFuncSetAttribute(wr<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0); wr<MODE><<<148 * ctas_per_sm, 128, smem>>>(map, d, stripes, nblk, steps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-58s %d CTA/SM: %7.1f us  %5.0f GB/s  %s\n", what, ctas_per_sm, best * 1e3, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int nblk = 49, steps = 13;
  const long long rows = 239616; const int stripes = (int)(rows / 128);
  const double bytes = (double)rows * nblk * 128;
  float* d; cudaMalloc(&d, (size_t)bytes);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)p;
  CUtensorMap map;
  cuuint64_t dims[3] = {32, (cuuint64_t)nblk, (cuuint64_t)rows};
  cuuint64_t str[2] = {128, (cuuint64_t)nblk * 128};
  cuuint32_t box[3] = {32, 1, 128}, ones[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  for (int c = 1; c <= 3; ++c) run<0>(map, d, stripes, nblk, steps, c, bytes, "mode 0: TMA stores only");
  for (int c = 1; c <= 3; ++c) run<1>(map, d, stripes, nblk, steps, c, bytes, "mode 1: staging + TMA stores, single-buffered (shipped)");
  run<2>(map, d, stripes, nblk, steps, 1, bytes, "mode 2: staging + TMA stores, double-buffered");
  for (int c = 1; c <= 3; ++c) run<3>(map, d, stripes, nblk, steps, c, bytes, "mode 3: staging + coalesced st.global.v4");
  return 0;
}
