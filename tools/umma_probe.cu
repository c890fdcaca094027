// Bring-up probe for tcgen05.mma.kind::tf32 descriptor semantics (not product code).
// One CTA computes D[128x128] = A[128x32] * B[128x32]^T with operands placed in shared memory by
// plain stores in three canonical layouts:
//   variant 0: K-major,  SWIZZLE_NONE   variant 1: MN-major, SWIZZLE_NONE   variant 2: MN-major, SWIZZLE_128B
// and prints max |D - ref| for each.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __host__ inline uint32_t elem_off(int variant, int mn, int k) {
  if (variant == 0) return ((mn / 8) * (K / 4) + (k / 4)) * 128 + (mn % 8) * 16 + (k % 4) * 4;
  if (variant == 1) return (mn % 4) * 4 + (k % 8) * 16 + (mn / 4) * 128 + (k / 8) * 4096;
  if (variant == 2) {
    uint32_t off = (mn / 32) * 4096 + (k / 8) * 1024 + (k % 8) * 128 + (mn % 32) * 4;
    return off ^ (((off >> 7) & 7) << 4);
  }
  uint32_t off = (mn / 32) * 4096 + (k / 4) * 512 + (k % 4) * 128 + (mn % 32) * 4;  // variant 3: 128B swizzle, 32B atoms
  return off ^ (((off >> 7) & 3) << 5);
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int variant, int skip_mma) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + 16384;
  uint64_t* bar = (uint64_t*)(base + 32768);
  uint32_t* slot = (uint32_t*)(base + 32768 + 8);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int e = tid; e < M * K; e += 128) {
    const int mn = e / K, k = e % K;
    *(float*)(sA + elem_off(variant, mn, k)) = A[e];
    *(float*)(sB + elem_off(variant, mn, k)) = B[e];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy smem writes must be visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;

  if (skip_mma) {  // TMEM store/load path only
    uint32_t val = __float_as_uint((float)(tid * 1000));
    for (int c = 0; c < N; ++c) {
      uint32_t v = __float_as_uint((float)(tid * 1000 + c));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    (void)val;
  } else if (warp == 1 && (tid & 31) == 0) {
    const uint32_t major = variant == 0 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (major << 15) | (major << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t lbo, sbo, step, lay;
    if (variant == 0) { lbo = 128; sbo = (K / 4) * 128; step = 256; lay = 0; }
    else if (variant == 1) { lbo = 4096; sbo = 128; step = 4096; lay = 0; }
    else if (variant == 2) { lbo = 4096; sbo = 1024; step = 1024; lay = 2; }
    else { lbo = 4096; sbo = 512; step = 1024; lay = 1; }
    for (int k = 0; k < K / 8; ++k) {
      auto mk = [&](uint32_t addr) {
        uint64_t d = 0;
        d |= (uint64_t)((addr >> 4) & 0x3FFF);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)lay << 61;
        return d;
      };
      const uint64_t ad = mk(smem_u32(sA) + k * step), bd = mk(smem_u32(sB) + k * step);
      const uint32_t acc = k != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  if (!skip_mma) {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    if (!ok) { printf("timeout\n"); __trap(); }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncwarp();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[tid * N + c * 32 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.0f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.0f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      R[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int skip = 1; skip >= 0; --skip)
    for (int variant = 0; variant < (skip ? 1 : 4); ++variant) {
      cudaMemset(dD, 0xFF, D.size() * 4);
      probe<<<1, 128, 40000>>>(dA, dB, dD, variant, skip);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0, mx = 0;
      int nz = 0;
      for (int i = 0; i < M * N; ++i) {
        const double want = skip ? (double)((i / N) * 1000 + (i % N)) : R[i];
        err = fmax(err, fabs(D[i] - want));
        mx = fmax(mx, fabs(want));
        nz += D[i] != 0.0f;
      }
      printf("skip_mma=%d variant=%d: %s  max|err|=%.4g (max|ref|=%.3g) nonzero=%d  D[0..3]=%g %g %g %g  ref=%g %g %g %g\n", skip, variant,
             cudaGetErrorString(e), err, mx, nz, D[0], D[1], D[2], D[3], skip ? 0.0 : R[0], skip ? 1.0 : R[1], skip ? 2.0 : R[2], skip ? 3.0 : R[3]);
    }
  return 0;
}
