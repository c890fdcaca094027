// Backward of A1 on the 5th-gen tensor cores (reference: autograd through
// torch.einsum('aijk,aijh->ajkh') / sqrt(C), models/stereoanywhere/corr.py:130-132 under train.py:383).
//
// With G = dLoss/dV [B,H,W2,W3] and s = post_scale / sqrt(C), per image row (b,h):
//     dL[c, w2] = s * sum_w3 R[c, w3] G[w2, w3]          (DR = false: K = W3)
//     dR[c, w3] = s * sum_w2 L[c, w2] G[w2, w3]          (DR = true : K = W2)
// i.e. D[M = channels, N = width] = A[M, K] B[N, K]^T with the OTHER feature map as A.  In NCHW the feature row
// (b, :, h, :) is [C][W] with W contiguous: K-MAJOR for this product (it was MN-major in the forward) - a 4-D TMA
// map {W, H, C, B} with box {32, 1, 128, 1} and the plain 128-byte swizzle drops a [128 channels][32 columns] slab
// into shared memory in the canonical UMMA K-major SWIZZLE_128B layout.  G is [W2][W3] with W3 contiguous:
//     dL: B[n = w2, k = w3] is K-major   -> one box {32, BN} per stage, SWIZZLE_128B;
//     dR: B[n = w3, k = w2] is MN-major  -> BN/32 boxes {32, 32} per stage, SWIZZLE_128B with 32-byte atoms,
//         exactly the forward kernel's operand layout.
// K (= 312 at KITTI size) is not a multiple of the 32-column stage: the TMA unit zero-fills the columns / rows
// beyond the tensor, in both operands.  Same persistent, warp-specialised structure as the forward kernel
// (TMA producer warp, MMA issuer warp, four epilogue warps, two 256-column TMEM accumulators); the epilogue scales
// and TMA-stores [128 channels][32 columns] boxes straight into the NCHW gradient.
// TF32 operands (rounded to nearest by the TMA unit), fp32 accumulate: normwise error ~3e-4 (gate 1e-3).
#include <cuda.h>
#include <stdlib.h>

#include "sa_common.cuh"
#include "tc_common.cuh"

namespace sa {
namespace bwd {

constexpr int kBM = 128;        // channels per accumulator tile
constexpr int kBK = 32;         // K columns per pipeline stage (4 UMMA k-steps of 8)
constexpr int kBox = 32;
constexpr int kABytes = kBM * 128;   // A slab: [128 rows][128 B]
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 224 * 1024;
constexpr int kEpiWarp0 = 2;
constexpr int kThreads = 192;
constexpr int kStagingBytes = 2 * kBM * 128;

struct Args {
  int C, H, WN, WK;          // N extent (output width), K extent (contracted width)
  int m_tiles, n_tiles, BN;
  int nstage;
  long long tiles;
  float scale;
};

template <bool DR>
__global__ void __launch_bounds__(kThreads, 1)
corr_bwd_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_g,
                     const __grid_constant__ CUtensorMap map_o, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_bytes = (uint32_t)a.BN * 128;          // B slab: BN rows x 128 B (dL) or BN/32 boxes of 4 KB (dR)
  const uint32_t stage_bytes = kABytes + b_bytes;
  uint8_t* stag = base;
  uint8_t* pipe = base + kStagingBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pipe + (size_t)a.nstage * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int kchunks = (a.WK + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
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

  // tile id -> (bh, tn, tm): the m-tiles (channel halves) of one (b, h, n-tile) run side by side and share G in L2
  auto decode = [&](long long tile, int& bh, int& tm, int& tn) {
    tm = (int)(tile % a.m_tiles); tile /= a.m_tiles;
    tn = (int)(tile % a.n_tiles);
    bh = (int)(tile / a.n_tiles);
  };

  if (warp == 0) {
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
      int bh, tm, tn;
      decode(tile, bh, tm, tn);
      const int b = bh / a.H, h = bh % a.H;
      const int m0 = tm * kBM, n0 = tn * a.BN;
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int s = it % a.nstage;
        const uint32_t ph = (it / a.nstage) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[s], stage_bytes);   // boxes are always counted whole: the unit zero-fills what is outside
          uint8_t* sa_ = pipe + (size_t)s * stage_bytes;
          uint8_t* sb_ = sa_ + kABytes;
          tma_load_4d(sa_, &map_a, &full[s], kc * kBK, h, m0, b);
          if (!DR) {
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(sb_)), "l"(&map_g), "r"(smem_u32(&full[s])), "r"(kc * kBK), "r"(n0), "r"(bh)
                : "memory");
          } else {
            const int nb = a.BN / kBox;
            for (int g = 0; g < nb; ++g)
              asm volatile(
                  "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  ::"r"(smem_u32(sb_ + g * 4096)), "l"(&map_g), "r"(smem_u32(&full[s])), "r"(n0 + g * kBox), "r"(kc * kBK),
                  "r"(bh)
                  : "memory");
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D = f32, A = B = tf32, A K-major, B K-major (dL) / MN-major (dR), N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((DR ? 1u : 0u) << 16) |
                           ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    uint32_t it = 0, lt = 0;
    for (long long tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++lt) {
      const uint32_t ab = lt & 1u;
      mbar_wait(&acc_empty[ab], ((lt >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * (uint32_t)kTmemCols;
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int s = it % a.nstage;
        const uint32_t ph = (it / a.nstage) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa_ = smem_u32(pipe + (size_t)s * stage_bytes);
          const uint32_t sb_ = sa_ + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            // K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart; a k-step of 8 tf32 = +32 B in the row
            const uint64_t ad = make_desc(sa_ + k * 32, 16, 1024, 2);
            const uint64_t bd = DR ? make_desc(sb_ + k * 1024, 4096, 512)      // MN-major, 32-byte atoms (as the forward)
                                   : make_desc(sb_ + k * 32, 16, 1024, 2);
            umma_tf32(d_tmem, ad, bd, idesc, (uint32_t)((kc | k) != 0));
          }
          umma_commit(&empty[s]);
          if (kc == kchunks - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    const int et = tid - kEpiWarp0 * 32;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;   // accumulator lane = channel m0 + row
    uint32_t lt = 0, cc = 0;
    for (long long tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++lt) {
      int bh, tm, tn;
      decode(tile, bh, tm, tn);
      const int b = bh / a.H, h = bh % a.H;
      const int m0 = tm * kBM, n0 = tn * a.BN;
      const uint32_t ab = lt & 1u;
      mbar_wait(&acc_full[ab], (lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ab * (uint32_t)kTmemCols + ((uint32_t)(quarter * 32) << 16);
      const int nchunks = (min(a.BN, a.WN - n0) + kBox - 1) / kBox;
      const bool live = m0 + quarter * 32 < a.C;
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
        if (c == nchunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[ab])) : "memory");
        }
        uint8_t* tile_s = stag + (cc & 1u) * (kBM * 128);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = __uint_as_float(v[4 * j + 0]) * a.scale;
            o.y = __uint_as_float(v[4 * j + 1]) * a.scale;
            o.z = __uint_as_float(v[4 * j + 2]) * a.scale;
            o.w = __uint_as_float(v[4 * j + 3]) * a.scale;
            *reinterpret_cast<float4*>(tile_s + row * 128 + ((j ^ (row & 7)) << 4)) = o;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&map_o),
                       "r"(smem_u32(tile_s)), "r"(n0 + c * kBox), "r"(h), "r"(m0), "r"(b)
                       : "memory");
          tma_commit();
          tma_wait_read<1>();
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

// plain (non-TF32-typed) map helper for the 128-byte swizzle with fp32 -> tf32 rounding by the TMA unit
static int make_map_tf32(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                         const cuuint32_t* box, CUtensorMapSwizzle swz, bool round_tf32, const char* what) {
  EncodeTiledFn fn = encode_fn();
  SA_REQUIRE(fn != nullptr, SA_E_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable (no driver?)");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void*>(ptr), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SA_REQUIRE(r == CUDA_SUCCESS, SA_E_INVALID, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}

template <bool DR>
static int launch(const float* other, const float* grad_vol, float* grad_out, int B, int C, int H, int WN, int WK, int W2,
                  int W3, float scale, cudaStream_t st) {
  Args a = {};
  a.C = C; a.H = H; a.WN = WN; a.WK = WK; a.scale = scale;
  a.m_tiles = (C + kBM - 1) / kBM;
  a.n_tiles = (WN + 255) / 256;
  const int per = (WN + a.n_tiles - 1) / a.n_tiles;
  a.BN = (per + kBox - 1) / kBox * kBox;
  const int stage_bytes = kABytes + a.BN * 128;
  a.nstage = (kSmemBudget - kStagingBytes) / stage_bytes;
  if (a.nstage > kMaxStages) a.nstage = kMaxStages;
  SA_REQUIRE(a.nstage >= 2, SA_E_UNSUPPORTED, "sa_corr_backward_tf32: tile does not fit shared memory");
  a.tiles = (long long)B * H * a.m_tiles * a.n_tiles;

  CUtensorMap ma, mg, mo;
  {  // the other feature map, K-major slabs [128 channels][32 columns of the contracted width]
    cuuint64_t dims[4] = {(cuuint64_t)WK, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)WK * 4, (cuuint64_t)H * WK * 4, (cuuint64_t)C * H * WK * 4};
    cuuint32_t box[4] = {kBox, 1, kBM, 1};
    int rc = make_map_tf32(&ma, other, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true, "feature map");
    if (rc) return rc;
  }
  {  // the volume gradient [B*H][W2][W3]
    cuuint64_t dims[3] = {(cuuint64_t)W3, (cuuint64_t)W2, (cuuint64_t)B * H};
    cuuint64_t str[2] = {(cuuint64_t)W3 * 4, (cuuint64_t)W2 * W3 * 4};
    cuuint32_t box_dl[3] = {kBox, (cuuint32_t)a.BN, 1};   // K-major: [BN rows of w2][32 columns of w3]
    cuuint32_t box_dr[3] = {kBox, kBK, 1};                // MN-major: [32 rows of w2 (K)][32 columns of w3 (N)]
    int rc = make_map_tf32(&mg, grad_vol, 3, dims, str, DR ? box_dr : box_dl,
                           DR ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, true, "grad_vol");
    if (rc) return rc;
  }
  {  // the feature-map gradient, stored as [128 channels][32 columns] boxes
    cuuint64_t dims[4] = {(cuuint64_t)WN, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)WN * 4, (cuuint64_t)H * WN * 4, (cuuint64_t)C * H * WN * 4};
    cuuint32_t box[4] = {kBox, 1, kBM, 1};
    int rc = make_map_tf32(&mo, grad_out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, false, "grad_fmap");
    if (rc) return rc;
  }
  const size_t smem = 1024 + kStagingBytes + (size_t)a.nstage * stage_bytes + (2 * kMaxStages + 5) * sizeof(uint64_t);
  auto kern = corr_bwd_tf32_kernel<DR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) SA_FAIL((int)e, "sa_corr_backward_tf32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const long long grid = a.tiles < (long long)num_sms() ? a.tiles : (long long)num_sms();
  kern<<<(unsigned)grid, kThreads, smem, st>>>(ma, mg, mo, a);
  return finish_launch("sa_corr_backward_tf32");
}

}  // namespace bwd
}  // namespace sa

extern "C" int sa_corr_backward_tf32(const float* grad_vol, const float* fmap_l, const float* fmap_r, float* grad_l,
                                     float* grad_r, int B, int C, int H, int W2, int W3, float divisor, float post_scale,
                                     void* stream) {
  using namespace sa;
  SA_REQUIRE(grad_vol && fmap_l && fmap_r && (grad_l || grad_r), SA_E_INVALID, "sa_corr_backward_tf32: null pointer");
  SA_REQUIRE(B > 0 && C > 0 && H > 0 && W2 > 0 && W3 > 0 && divisor != 0.f, SA_E_INVALID, "sa_corr_backward_tf32: bad sizes");
  SA_REQUIRE(W2 % 4 == 0 && W3 % 4 == 0, SA_E_UNSUPPORTED, "sa_corr_backward_tf32: W2 and W3 must be multiples of 4");
  SA_REQUIRE(aligned16(grad_vol) && aligned16(fmap_l) && aligned16(fmap_r) && (!grad_l || aligned16(grad_l)) &&
                 (!grad_r || aligned16(grad_r)),
             SA_E_ALIGN, "sa_corr_backward_tf32: pointers must be 16-byte aligned");
  const float scale = kernel_inv_divisor(divisor) * post_scale;
  if (grad_l) {
    int rc = bwd::launch<false>(fmap_r, grad_vol, grad_l, B, C, H, W2, W3, W2, W3, scale, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (grad_r) {
    int rc = bwd::launch<true>(fmap_l, grad_vol, grad_r, B, C, H, W3, W2, W2, W3, scale, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}
