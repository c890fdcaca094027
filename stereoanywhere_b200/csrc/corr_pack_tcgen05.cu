// A1 + A5 + A3 in ONE kernel: the stereo correlation on the tensor cores with the truncation product and
// the avg-pooled pyramid formed in the GEMM epilogue, written once, directly as the line-packed pyramid
// that the lookups read (csrc/packed.cu).  The fp32 volume itself never exists in memory.
//   reference: models/stereoanywhere/corr.py:117-132 (einsum / sqrt(C)), :76-91 (avg_pool2d pyramid),
//              utils/utils.py:216-238 + stereoanywhere.py:203,253-255 (truncation mask product).
//
// Work unit = a "stripe": 128 left pixels (w2) of one image row (b,h) against ALL W3 right columns.  A
// persistent CTA per SM walks over stripes; a stripe is multiplied in n-tiles of <= 256 columns into the two
// 256-column TMEM accumulators (same operand path as corr_tcgen05.cu: TMA boxes straight from NCHW, MN-major
// TF32).  The four epilogue warps own one output row per thread (TMEM lane = w2) and STREAM over the row in
// chunks of 32 columns, left to right, across the n-tiles of the stripe:
//     chunk cc -> scale, x truncation mask -> 16 + 8 + 4 pooled values (levels 1..3)
// A line q of the packed layout needs L0[8q-4 .. 8q+12] and pooled values of blocks q-4 .. q+5, so step cc
// can emit the four lines q = 4cc-5 .. 4cc-2 from a register window of the last 44 level-0 values and
// 24 / 14 / 9 values of levels 1 / 2 / 3 - all indices compile-time constants.  The four lines go through four
// swizzled [128 rows][128 B] staging tiles and leave as one TMA store per line index (4-D map {32, line, w2, b*h}: rows
// beyond W2 are clipped by the map).  Columns outside [0, W3) are zeros (TMA zero fill / virtual chunks),
// which is exactly the layout's zero padding.
//
// Same arithmetic, in the same order, as sa_corr_tf32 followed by sa_pack_pyramid: the results are
// bit-identical (tests/test_gpu_parity.py::test_corr_pack_fused_matches_two_step).
// HBM: reads 2*B*C*H*W*4, writes B*H*W2*(W3/8+9)*128; bound by the packed write.
#include <cuda.h>

#include "pack_stream.cuh"
#include "tc_common.cuh"

namespace sa {
namespace cpk {

constexpr int kBM = 128;       // rows of the accumulator tile (w2)
constexpr int kBK = 32;        // channels per pipeline stage (4 UMMA k-steps of 8)
constexpr int kBox = 32;       // fp32 columns per TMA box = 128 B = one swizzle row = one epilogue chunk
constexpr int kBoxBytes = kBox * kBK * 4;  // 4096
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 224 * 1024;
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;
constexpr int kNStg = 4;                         // [128 rows][128 B] staging tiles: the four lines of one step
constexpr int kStagingBytes = kNStg * kBM * 128;

struct Args {
  int C, H, W2, W3;
  int m_tiles, n_tiles;
  int cpt;      // chunks (of 32 columns) per n-tile, <= 8
  int nch;      // chunks that hold real columns = ceil(W3 / 32)
  int n_steps;  // epilogue steps per stripe (4 lines each)
  int nblk;     // lines per row = W3 / 8 + 9
  int nstage;
  long long stripes;
  float scale;
  const float* disp;  // truncation (TRUNC only): one float per (b,h,w2)
  const float* conf;
  float gain, one_minus_gain;
};

// HK = 0: fp32 lines (128 B); 1 / 2: fp16 / bf16 lines (64 B) - the packed array, the staging tiles and the TMA
// stores shrink to half (opt-in storage mode; the fp32 arithmetic up to the rounding of the stored values is the same)
template <bool TRUNC, int HK>
__global__ void __launch_bounds__(kThreads, 1)
corr_pack_tf32_kernel(const __grid_constant__ CUtensorMap map_l, const __grid_constant__ CUtensorMap map_r,
                      const __grid_constant__ CUtensorMap map_o, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = (uint32_t)(kBM / kBox + a.cpt) * kBoxBytes;
  constexpr int kLineBytes = HK ? 64 : 128;
  constexpr int kTileBytes = kBM * kLineBytes;
  uint8_t* stag = base;
  uint8_t* pipe = base + kNStg * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pipe + (size_t)a.nstage * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int kchunks = a.C / kBK;

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
    // ------------------------------------------------------------------ TMA producer
    uint32_t it = 0;
    for (long long stripe = blockIdx.x; stripe < a.stripes; stripe += gridDim.x) {
      const int tm = (int)(stripe % a.m_tiles);
      const int bh = (int)(stripe / a.m_tiles);
      const int b = bh / a.H, h = bh % a.H;
      const int m0 = tm * kBM;
      const int a_boxes = min(kBM / kBox, (a.W2 - m0 + kBox - 1) / kBox);
      for (int tn = 0; tn < a.n_tiles; ++tn) {
        const int c0 = tn * a.cpt;
        const int b_boxes = min(a.cpt, a.nch - c0);
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
              if (g < b_boxes) tma_load_4d(sb_ + g * kBoxBytes, &map_r, &full[s], (c0 + g) * kBox, h, kc * kBK, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kBM >> 4) << 24);
    uint32_t it = 0, lt = 0;
    for (long long stripe = blockIdx.x; stripe < a.stripes; stripe += gridDim.x) {
      for (int tn = 0; tn < a.n_tiles; ++tn, ++lt) {
        const int nb = min(a.cpt, a.nch - tn * a.cpt);
        const uint32_t idesc = idesc0 | ((uint32_t)((nb * kBox) >> 3) << 17);
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
            const uint32_t sb_ = sa_ + (kBM / kBox) * kBoxBytes;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t ad = make_desc(sa_ + k * 1024, kBoxBytes, 512);
              const uint64_t bd = make_desc(sb_ + k * 1024, kBoxBytes, 512);
              umma_tf32(d_tmem, ad, bd, idesc, (uint32_t)((kc | k) != 0));
            }
            umma_commit(&empty[s]);
            if (kc == kchunks - 1) umma_commit(&acc_full[ab]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps (streaming packer)
    const int et = tid - kEpiWarp0 * 32;   // 0..127
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;   // accumulator lane = output row m0 + row
    const int swz = row & 7;
    uint32_t lt = 0;                       // n-tile counter
    for (long long stripe = blockIdx.x; stripe < a.stripes; stripe += gridDim.x) {
      const int tm = (int)(stripe % a.m_tiles);
      const int bh = (int)(stripe / a.m_tiles);
      const int m0 = tm * kBM;
      const bool live = m0 + quarter * 32 < a.W2;  // warp-uniform
      float tc = 0.f, omc = 1.f, centre = 0.f;
      if (TRUNC) {
        const int w2 = min(m0 + row, a.W2 - 1);
        const long long r = (long long)bh * a.W2 + w2;
        tc = __ldg(a.conf + r);
        omc = 1.0f - tc;
        centre = (float)w2 - __ldg(a.disp + r);
      }
      PackWindow win;  // register window of the streaming packer (pack_stream.cuh)
      win.reset();
      int tn = 0, wi = 0;  // n-tile / chunk-in-tile of the current chunk
      for (int cc = 0; cc < a.n_steps; ++cc) {
        float v[32];
        if (cc < a.nch) {
          const uint32_t ltc = lt + (uint32_t)tn;
          const uint32_t ab = ltc & 1u;
          const int nb = min(a.cpt, a.nch - tn * a.cpt);
          if (wi == 0) {
            mbar_wait(&acc_full[ab], (ltc >> 1) & 1u);
            tc_fence_after();
          }
          if (live) {
            uint32_t u[32];
            const uint32_t taddr = tmem_base + ab * (uint32_t)kTmemCols + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(wi * kBox);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                  "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                  "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                  "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 q4;
              q4.x = __uint_as_float(u[4 * g + 0]) * a.scale;
              q4.y = __uint_as_float(u[4 * g + 1]) * a.scale;
              q4.z = __uint_as_float(u[4 * g + 2]) * a.scale;
              q4.w = __uint_as_float(u[4 * g + 3]) * a.scale;
              if (TRUNC) trunc_mask_mul4(q4, centre, (float)(cc * kBox + 4 * g), tc, omc, a.gain, a.one_minus_gain);
              v[4 * g + 0] = q4.x; v[4 * g + 1] = q4.y; v[4 * g + 2] = q4.z; v[4 * g + 3] = q4.w;
            }
          }
          if (wi == nb - 1) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[ab])) : "memory");
            ++tn;
            wi = 0;
          } else {
            ++wi;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (live) win.pool(v);
        // ---- the four lines of this step: one staging tile each, one barrier round, four TMA stores
        if (et == 0) tma_wait_read<0>();                // the previous step's stores have read the tiles
        asm volatile("bar.sync 1, 128;" ::: "memory");  // (they had this step's TMEM load + pooling to do so)
        if (live) {
          const int li = 4 * cc;
          if (HK == 0) {
            if (li + 0 < a.nblk) win.store_line<0>(stag + 0 * kTileBytes + row * 128, swz);
            if (li + 1 < a.nblk) win.store_line<1>(stag + 1 * kTileBytes + row * 128, swz);
            if (li + 2 < a.nblk) win.store_line<2>(stag + 2 * kTileBytes + row * 128, swz);
            if (li + 3 < a.nblk) win.store_line<3>(stag + 3 * kTileBytes + row * 128, swz);
          } else {
            constexpr int K = HK ? HK : 1;
            if (li + 0 < a.nblk) win.store_line_half<0, K>(stag + 0 * kTileBytes + row * 64, row);
            if (li + 1 < a.nblk) win.store_line_half<1, K>(stag + 1 * kTileBytes + row * 64, row);
            if (li + 2 < a.nblk) win.store_line_half<2, K>(stag + 2 * kTileBytes + row * 64, row);
            if (li + 3 < a.nblk) win.store_line_half<3, K>(stag + 3 * kTileBytes + row * 64, row);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (4 * cc + j < a.nblk) tma_store_4d(&map_o, stag + j * kTileBytes, 0, 4 * cc + j, m0, bh);
          tma_commit();
        }
        if (live) win.advance(v);
      }
      lt += (uint32_t)a.n_tiles;
    }
    if (et == 0) tma_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace cpk
}  // namespace sa

namespace sa {
namespace cpk {

static int launch_corr_pack(const float* fmap_l, const float* fmap_r, int B, int C, int H, int W2, int W3, float divisor,
                            float post_scale, const float* trunc_disp, const float* trunc_conf, double trunc_gain,
                            void* packed, int hk, cudaStream_t stream) {
  SA_REQUIRE(fmap_l && fmap_r && packed, SA_E_INVALID, "sa_corr_pack_tf32: null pointer");
  SA_REQUIRE(B > 0 && C > 0 && H > 0 && W2 > 0 && W3 > 0, SA_E_INVALID, "sa_corr_pack_tf32: sizes must be positive");
  SA_REQUIRE(divisor != 0.f, SA_E_INVALID, "sa_corr_pack_tf32: divisor == 0");
  SA_REQUIRE((trunc_disp == nullptr) == (trunc_conf == nullptr), SA_E_INVALID,
             "sa_corr_pack_tf32: trunc_disp / trunc_conf must come together");
  SA_REQUIRE(C % kBK == 0, SA_E_UNSUPPORTED, "sa_corr_pack_tf32: C must be a multiple of %d (got %d)", kBK, C);
  SA_REQUIRE(W2 % 4 == 0 && W3 % 8 == 0, SA_E_UNSUPPORTED,
             "sa_corr_pack_tf32: W2 must be a multiple of 4 and W3 a multiple of 8 (got %d, %d)", W2, W3);
  SA_REQUIRE((long long)B * H <= 0x7fffffffLL, SA_E_UNSUPPORTED, "sa_corr_pack_tf32: B*H too large");
  SA_REQUIRE(aligned16(fmap_l) && aligned16(fmap_r) && aligned16(packed), SA_E_ALIGN,
             "sa_corr_pack_tf32: pointers must be 16-byte aligned");
  SA_REQUIRE(hk >= 0 && hk <= 2, SA_E_INVALID, "sa_corr_pack_tf32: storage kind must be 0 (fp32), 1 (fp16) or 2 (bf16)");

  Args a = {};
  a.C = C; a.H = H; a.W2 = W2; a.W3 = W3;
  a.scale = kernel_inv_divisor(divisor) * post_scale;
  a.m_tiles = (W2 + kBM - 1) / kBM;
  a.nch = (W3 + kBox - 1) / kBox;
  a.n_tiles = (a.nch + 7) / 8;
  a.cpt = (a.nch + a.n_tiles - 1) / a.n_tiles;
  a.nblk = W3 / 8 + 9;
  a.n_steps = (a.nblk - 1) / 4 + 1;
  const int line_bytes = hk ? 64 : 128;
  const int staging = kNStg * kBM * line_bytes;
  const int stage_bytes = (kBM / kBox + a.cpt) * kBoxBytes;
  a.nstage = (kSmemBudget - staging) / stage_bytes;
  if (a.nstage > kMaxStages) a.nstage = kMaxStages;
  SA_REQUIRE(a.nstage >= 2, SA_E_UNSUPPORTED, "sa_corr_pack_tf32: tile does not fit shared memory");
  a.stripes = (long long)B * H * a.m_tiles;
  a.disp = trunc_disp; a.conf = trunc_conf;
  a.gain = (float)trunc_gain; a.one_minus_gain = (float)(1.0 - trunc_gain);

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
  {  // lines as 32-bit words: 32 per line (fp32) or 16 (two 16-bit values per word)
    const cuuint64_t lb = (cuuint64_t)line_bytes;
    cuuint64_t dims[4] = {lb / 4, (cuuint64_t)a.nblk, (cuuint64_t)W2, (cuuint64_t)B * H};
    cuuint64_t str[3] = {lb, (cuuint64_t)a.nblk * lb, (cuuint64_t)W2 * a.nblk * lb};
    cuuint32_t box[4] = {(cuuint32_t)(lb / 4), 1, kBM, 1};
    int rc = make_map(&mo, packed, 4, dims, str, box, hk ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, "packed");
    if (rc) return rc;
  }
  const size_t smem = 1024 + staging + (size_t)a.nstage * stage_bytes + (2 * kMaxStages + 5) * sizeof(uint64_t);
  const bool trunc = trunc_disp != nullptr;
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Args) =
      hk == 0 ? (trunc ? corr_pack_tf32_kernel<true, 0> : corr_pack_tf32_kernel<false, 0>)
      : hk == 1 ? (trunc ? corr_pack_tf32_kernel<true, 1> : corr_pack_tf32_kernel<false, 1>)
                : (trunc ? corr_pack_tf32_kernel<true, 2> : corr_pack_tf32_kernel<false, 2>);
  {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) SA_FAIL((int)e, "sa_corr_pack_tf32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const long long grid = a.stripes < (long long)num_sms() ? a.stripes : (long long)num_sms();
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(ml, mr, mo, a);
  return finish_launch("sa_corr_pack_tf32");
}

}  // namespace cpk
}  // namespace sa

extern "C" int sa_corr_pack_tf32(const float* fmap_l, const float* fmap_r, int B, int C, int H, int W2, int W3,
                                 float divisor, float post_scale, const float* trunc_disp, const float* trunc_conf,
                                 double trunc_gain, float* packed, void* stream) {
  return sa::cpk::launch_corr_pack(fmap_l, fmap_r, B, C, H, W2, W3, divisor, post_scale, trunc_disp, trunc_conf, trunc_gain,
                                   packed, 0, (cudaStream_t)stream);
}

extern "C" int sa_corr_pack_tf32_half(const float* fmap_l, const float* fmap_r, int B, int C, int H, int W2, int W3,
                                      float divisor, float post_scale, const float* trunc_disp, const float* trunc_conf,
                                      double trunc_gain, int half_kind, void* packed_h, void* stream) {
  SA_REQUIRE(half_kind == 1 || half_kind == 2, SA_E_INVALID, "sa_corr_pack_tf32_half: half_kind must be 1 (fp16) or 2 (bf16)");
  return sa::cpk::launch_corr_pack(fmap_l, fmap_r, B, C, H, W2, W3, divisor, post_scale, trunc_disp, trunc_conf, trunc_gain,
                                   packed_h, half_kind, (cudaStream_t)stream);
}
