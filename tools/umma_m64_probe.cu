// Probe (not product code): where does tcgen05.mma.cta_group::1.kind::tf32 with M = 64 put its rows in TMEM, and which
// lane offsets of the D address are legal?  A[128x32], B[128x32] in the product's operand layout (MN-major, 128B swizzle
// with 32-byte atoms).  MMA 1: rows 0..63 of A x B[0..63]^T at D = tmem + (off1 << 16); MMA 2 (optional): rows 64..127 of
// A x B[64..127]^T at D = tmem + (off2 << 16).  Dumps, for every TMEM lane, which row of which product it holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_m64_probe umma_m64_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __host__ inline uint32_t elem_off(int mn, int k) {
  uint32_t off = (mn / 32) * 4096 + (k / 4) * 512 + (k % 4) * 128 + (mn % 32) * 4;
  return off ^ (((off >> 7) & 3) << 5);
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int off1, int off2, int two) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + 16384;
  uint64_t* bar = (uint64_t*)(base + 32768);
  uint32_t* slot = (uint32_t*)(base + 32768 + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < M * K; e += 128) {
    const int mn = e / K, k = e % K;
    *(float*)(sA + elem_off(mn, k)) = A[e];
    *(float*)(sB + elem_off(mn, k)) = B[e];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  // fill the whole allocation with a marker so untouched lanes are recognisable
  for (int c = 0; c < 128; ++c) {
    uint32_t v = __float_as_uint(-12345.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1 && (tid & 31) == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    auto mk = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((4096u >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)1 << 61;
      return d;
    };
    for (int pass = 0; pass < (two ? 2 : 1); ++pass) {
      const uint32_t d_addr = tmem + ((uint32_t)(pass ? off2 : off1) << 16);
      for (int k = 0; k < K / 8; ++k) {
        const uint64_t ad = mk(smem_u32(sA) + pass * 2 * 4096 + k * 1024), bd = mk(smem_u32(sB) + pass * 2 * 4096 + k * 1024);
        const uint32_t acc = k != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_addr),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
  if (!ok) { printf("timeout\n"); __trap(); }
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
  std::vector<float> A(M * K), B(M * K), D(128 * N);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.0f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.0f;
  // reference products: P1[r][n] = A[r] . B[n] (r, n < 64), P2[r][n] = A[64 + r] . B[64 + n]
  auto dot = [&](int ar, int br) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[ar * K + k] * B[br * K + k]; return (float)s; };
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int cases[][3] = {{0, 0, 0}, {0, 16, 1}, {0, 64, 1}, {0, 32, 1}, {16, 0, 0}};
  for (auto& cs : cases) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, 40000>>>(dA, dB, dD, cs[0], cs[1], cs[2]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("=== off1=%d off2=%d two=%d: %s\n", cs[0], cs[1], cs[2], cudaGetErrorString(e));
    if (e != cudaSuccess) { cudaGetLastError(); break; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    for (int lane = 0; lane < 128; ++lane) {
      int found = -1, which = 0;
      for (int w = 0; w < 2 && found < 0; ++w)
        for (int r = 0; r < 64 && found < 0; ++r) {
          bool match = true;
          for (int n = 0; n < 8 && match; ++n) match = fabsf(D[lane * N + n] - dot(w * 64 + r, w * 64 + n)) < 2e-3f;
          if (match) { found = r; which = w + 1; }
        }
      if (found >= 0) printf("lane %3d: P%d row %2d\n", lane, which, found);
      else if (D[lane * N] != -12345.0f) printf("lane %3d: ??? %g %g\n", lane, D[lane * N], D[lane * N + 1]);
    }
  }
  return 0;
}
