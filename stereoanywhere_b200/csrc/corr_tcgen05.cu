// A1 - stereo correlation volume on the 5th-gen tensor cores
// (reference: models/stereoanywhere/corr.py:117-132, torch.einsum('aijk,aijh->ajkh') / sqrt(C)).
//
// Per image row (b,h):  V[w2,w3] = sum_c L[b,c,h,w2] R[b,c,h,w3].  Both operands are W-contiguous in
// NCHW, i.e. "MN-major" for the MMA: a 4-D TMA tensor map {W,H,C,B} with box {32,1,BK,1} and the
// 128-byte swizzle drops a [BK channels][32 columns] slab straight into shared memory in the
// canonical UMMA MN-major SWIZZLE_128B (32-byte atom) layout - no transpose, no conversion pass; the fp32 bits
// are consumed as TF32 by tcgen05.mma.kind::tf32 with an fp32 accumulator in TMEM.
//
// One persistent CTA per SM walks over [128 x BN<=256] output tiles; the two 256-column halves of
// TMEM hold two accumulators so the tensor core works on tile i+1 while tile i is being stored.
//   warp 0 / lane 0 : TMA producer (mbarrier full/empty ring, BK = 32 channels per stage)
//   warp 1 / lane 0 : tcgen05.mma issuer, tcgen05.commit frees the stage / publishes the accumulator
//   warps 2..5      : epilogue - tcgen05.ld (lane = output row), scale, swizzled st.shared,
//                     TMA store of [128 x 32] boxes (rows / columns outside the image are clipped
//                     by the 3-D output tensor map {W3, W2, B*H}).
// HBM-bound by the fp32 volume write (AI ~ 36 flop/B at C=256, W=312); the tensor pipe idles ~2/3.
#include <cuda.h>
#include <stdlib.h>

#include "sa_common.cuh"
#include "tc_common.cuh"

namespace sa {

constexpr int kBM = 128;       // rows of the accumulator tile (w2)
constexpr int kBK = 32;        // channels per pipeline stage (4 UMMA k-steps of 8)
constexpr int kBox = 32;       // fp32 columns per TMA box = 128 B = one swizzle row
constexpr int kBoxBytes = kBox * kBK * 4;  // 4096
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 224 * 1024;    // one persistent CTA per SM

struct CorrTcArgs {
  int C, H, W2, W3;
  int m_tiles, n_tiles, BN;  // BN: accumulator columns per tile (multiple of 32, <= 256)
  int nstage;
  long long tiles;
  float inv_divisor, post_scale;
};

constexpr int kEpiWarp0 = 2;     // warps 2..5 drain the accumulators
constexpr int kThreads = 192;
constexpr int kStagingBytes = 2 * kBM * 128;  // two [128 rows][128 B] tiles for the TMA stores

// Persistent, warp-specialised: one CTA per SM loops over output tiles (tile id = blockIdx.x + i *
// gridDim.x, n-tile fastest so that CTAs running side by side share their operands in L2).
//   warp 0 lane 0   TMA producer   : ring of `nstage` [BK=32 channel] operand slabs
//   warp 1 lane 0   MMA issuer     : 4 x tcgen05.mma (K=8) per slab into one of TWO 256-column TMEM
//                                    accumulators, so tile i+1 is multiplied while tile i is drained
//   warps 2..5      epilogue       : tcgen05.ld -> scale -> swizzled st.shared -> TMA store
__global__ void __launch_bounds__(kThreads, 1)
corr_tf32_kernel(const __grid_constant__ CUtensorMap map_l, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ CUtensorMap map_o, const CorrTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_boxes_b = a.BN / kBox;
  const uint32_t stage_bytes = (uint32_t)(kBM / kBox + n_boxes_b) * kBoxBytes;
  uint8_t* stag = base;                             // staging tiles (1024-byte aligned)
  uint8_t* pipe = base + kStagingBytes;             // operand ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(pipe + (size_t)a.nstage * stage_bytes);
  uint64_t* full = bars;                        // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;          // [kMaxStages]
  uint64_t* acc_full = bars + 2 * kMaxStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int kchunks = a.C / kBK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_l) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp loops,
    // one elected lane issues: addresses and coordinates stay in uniform registers)
    uint32_t it = 0;  // slab counter across tiles
    for (long long tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
      long long t = tile;
      const int tn = (int)(t % a.n_tiles); t /= a.n_tiles;
      const int tm = (int)(t % a.m_tiles);
      const int bh = (int)(t / a.m_tiles);
      const int b = bh / a.H, h = bh % a.H;
      const int m0 = tm * kBM, n0 = tn * a.BN;
      // boxes wholly outside the image are neither loaded nor counted: their accumulator rows /
      // columns are clipped by the output tensor map
      const int a_boxes = min(kBM / kBox, (a.W2 - m0 + kBox - 1) / kBox);
      const int b_boxes = min(n_boxes_b, (a.W3 - n0 + kBox - 1) / kBox);
      const uint32_t tx_bytes = (uint32_t)(a_boxes + b_boxes) * kBoxBytes;
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int s = it % a.nstage;
        const uint32_t ph = (it / a.nstage) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[s], tx_bytes);
          uint8_t* sa_ = pipe + (size_t)s * stage_bytes;
          uint8_t* sb_ = sa_ + (kBM / kBox) * kBoxBytes;
#pragma unroll
          for (int g = 0; g < kBM / kBox; ++g)
            if (g < a_boxes) tma_load_4d(sa_ + g * kBoxBytes, &map_l, &full[s], m0 + g * kBox, h, kc * kBK, b);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g < b_boxes) tma_load_4d(sb_ + g * kBoxBytes, &map_r, &full[s], n0 + g * kBox, h, kc * kBK, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer (same pattern)
    // instruction descriptor: D=f32, A=B=tf32, both MN-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    uint32_t it = 0, lt = 0;
    for (long long tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++lt) {
      const uint32_t ab = lt & 1u;
      mbar_wait(&acc_empty[ab], ((lt >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * (uint32_t)kTmemCols;
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int s = it % a.nstage;
        const uint32_t ph = (it / a.nstage) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa_ = smem_u32(pipe + (size_t)s * stage_bytes);
          const uint32_t sb_ = sa_ + (kBM / kBox) * kBoxBytes;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t ad = make_desc(sa_ + k * 1024, kBoxBytes, 512);
            const uint64_t bd = make_desc(sb_ + k * 1024, kBoxBytes, 512);
            umma_tf32(d_tmem, ad, bd, idesc, (uint32_t)((kc | k) != 0));
          }
          umma_commit(&empty[s]);  // slab reusable once these MMAs have read it
          if (kc == kchunks - 1) umma_commit(&acc_full[ab]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps
    const int et = tid - kEpiWarp0 * 32;   // 0..127
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;   // accumulator lane = output row m0 + row
    const float scale = a.inv_divisor * a.post_scale;
    uint32_t lt = 0, cc = 0;               // local tile counter, staging-chunk counter
    for (long long tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++lt) {
      long long t = tile;
      const int tn = (int)(t % a.n_tiles); t /= a.n_tiles;
      const int tm = (int)(t % a.m_tiles);
      const int bh = (int)(t / a.m_tiles);
      const int m0 = tm * kBM, n0 = tn * a.BN;
      const uint32_t ab = lt & 1u;
      mbar_wait(&acc_full[ab], (lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ab * (uint32_t)kTmemCols + ((uint32_t)(quarter * 32) << 16);
      const int nchunks = (min(a.BN, a.W3 - n0) + kBox - 1) / kBox;
      const bool live = m0 + quarter * 32 < a.W2;  // warp-uniform: any row of this quarter inside the image
      for (int c = 0; c < nchunks; ++c, ++cc) {
        uint32_t v[32];
        if (live) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(t_lane + (uint32_t)(c * kBox)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (c == nchunks - 1) {  // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[ab])) : "memory");
        }
        uint8_t* tile_s = stag + (cc & 1u) * (kBM * 128);
        // the store issued two chunks ago (same staging tile) must have finished reading it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = __uint_as_float(v[4 * j + 0]) * scale;
            o.y = __uint_as_float(v[4 * j + 1]) * scale;
            o.z = __uint_as_float(v[4 * j + 2]) * scale;
            o.w = __uint_as_float(v[4 * j + 3]) * scale;
            *reinterpret_cast<float4*>(tile_s + row * 128 + ((j ^ (row & 7)) << 4)) = o;  // 128B swizzle
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          tma_store_3d(&map_o, tile_s, n0 + c * kBox, m0, bh);
          tma_commit();
          tma_wait_read<1>();  // <= 1 store in flight: the other staging tile is free again
        }
      }
    }
    if (et == 0) tma_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace sa

extern "C" int sa_corr_tf32(const float* fmap_l, const float* fmap_r, float* vol, int B, int C, int H, int W2, int W3,
                            float divisor, float post_scale, const float* trunc_disp, const float* trunc_conf,
                            double trunc_gain, float* pyr1, float* pyr2, float* pyr3, int64_t pitch1, int64_t pitch2,
                            int64_t pitch3, void* stream) {
  using namespace sa;
  (void)trunc_gain; (void)pitch1; (void)pitch2; (void)pitch3;
  SA_REQUIRE(fmap_l && fmap_r && vol, SA_E_INVALID, "sa_corr_tf32: null pointer");
  SA_REQUIRE(B > 0 && C > 0 && H > 0 && W2 > 0 && W3 > 0, SA_E_INVALID, "sa_corr_tf32: sizes must be positive");
  SA_REQUIRE(divisor != 0.f, SA_E_INVALID, "sa_corr_tf32: divisor == 0");
  SA_REQUIRE(C % kBK == 0, SA_E_UNSUPPORTED, "sa_corr_tf32: C must be a multiple of %d (got %d)", kBK, C);
  SA_REQUIRE(W2 % 4 == 0 && W3 % 4 == 0, SA_E_UNSUPPORTED, "sa_corr_tf32: W2 and W3 must be multiples of 4");
  SA_REQUIRE(aligned16(fmap_l) && aligned16(fmap_r) && aligned16(vol), SA_E_ALIGN,
             "sa_corr_tf32: pointers must be 16-byte aligned");
  SA_REQUIRE(!trunc_disp && !trunc_conf && !pyr1 && !pyr2 && !pyr3, SA_E_UNSUPPORTED,
             "sa_corr_tf32: fused truncation / pyramid epilogue is not available in this build");

  CorrTcArgs a = {};
  a.C = C; a.H = H; a.W2 = W2; a.W3 = W3;
  a.inv_divisor = kernel_inv_divisor(divisor); a.post_scale = post_scale;
  a.m_tiles = (W2 + kBM - 1) / kBM;
  a.n_tiles = (W3 + 255) / 256;
  const int per = (W3 + a.n_tiles - 1) / a.n_tiles;
  a.BN = (per + kBox - 1) / kBox * kBox;  // multiple of 32 (hence of 16: legal UMMA N for M = 128)
  const int stage_bytes = (kBM / kBox + a.BN / kBox) * kBoxBytes;
  a.nstage = (kSmemBudget - kStagingBytes) / stage_bytes;
  if (a.nstage > kMaxStages) a.nstage = kMaxStages;
  SA_REQUIRE(a.nstage >= 2, SA_E_UNSUPPORTED, "sa_corr_tf32: tile does not fit shared memory");
  const long long tiles = (long long)B * H * a.m_tiles * a.n_tiles;
  a.tiles = tiles;

  CUtensorMap ml, mr, mo;
  {
    cuuint64_t dims[4] = {(cuuint64_t)W2, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W2 * 4, (cuuint64_t)H * W2 * 4, (cuuint64_t)C * H * W2 * 4};
    cuuint32_t box[4] = {kBox, 1, kBK, 1};
    int rc = make_map(&ml, fmap_l, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "fmap_l");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)W3, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W3 * 4, (cuuint64_t)H * W3 * 4, (cuuint64_t)C * H * W3 * 4};
    cuuint32_t box[4] = {kBox, 1, kBK, 1};
    int rc = make_map(&mr, fmap_r, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "fmap_r");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)W3, (cuuint64_t)W2, (cuuint64_t)B * H};
    cuuint64_t str[2] = {(cuuint64_t)W3 * 4, (cuuint64_t)W2 * W3 * 4};
    cuuint32_t box[3] = {kBox, kBM, 1};
    int rc = make_map(&mo, vol, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "vol");
    if (rc) return rc;
  }
  const size_t smem = 1024 + kStagingBytes + (size_t)a.nstage * stage_bytes + (2 * kMaxStages + 5) * sizeof(uint64_t);
  {  // per device and cheap: always (re)state the dynamic shared-memory opt-in
    cudaError_t e = cudaFuncSetAttribute(corr_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) SA_FAIL((int)e, "sa_corr_tf32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const long long grid = tiles < (long long)num_sms() ? tiles : (long long)num_sms();
  corr_tf32_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(ml, mr, mo, a);
  return finish_launch("sa_corr_tf32");
}
