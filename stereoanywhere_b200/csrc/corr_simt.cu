// A1 / A2 - all-pairs 1-D correlation volume, fp32 SIMT kernel
// (reference: models/stereoanywhere/corr.py:117-132, `1.73 *` of stereoanywhere.py:136).
//
// vol[b,h,w2,w3] = (sum_c L[b,c,h,w2] R[b,c,h,w3]) / divisor * post_scale, exact fp32 FMA
// accumulation.  This is the kernel of record for the mono volume (C = 3: three FMAs per output,
// a pure HBM streaming write) and the "fp32" precision mode of the stereo volume; the tensor
// core kernel (corr_tcgen05.cu) is the fast path for C = 128 / 256.
//
// A CTA computes a 64 x 64 (w2 x w3) tile of one image row (b,h): both operands are W-contiguous
// in NCHW, so a 16-channel slab of each is staged in shared memory with coalesced loads and each
// thread keeps a 4 x 4 register tile.
#include "sa_common.cuh"

namespace sa {

constexpr int kTile = 64;
constexpr int kChunk = 16;

__global__ void __launch_bounds__(256)
corr_simt_kernel(const float* __restrict__ fl, const float* __restrict__ fr, float* __restrict__ vol, int C, int H,
                 int W2, int W3, int tiles_m, int tiles_n, float divisor, float inv_divisor, float post_scale) {
  __shared__ __align__(16) float sA[kChunk][kTile];
  __shared__ __align__(16) float sB[kChunk][kTile];

  int tile = blockIdx.x;
  const int tn = tile % tiles_n;
  tile /= tiles_n;
  const int tm = tile % tiles_m;
  const int bh = tile / tiles_m;
  const int b = bh / H, h = bh % H;
  const int m0 = tm * kTile, n0 = tn * kTile;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long plane2 = (long long)H * W2, plane3 = (long long)H * W3;
  const float* pl = fl + ((long long)b * C * H + h) * W2;  // + c*plane2 + w2
  const float* pr = fr + ((long long)b * C * H + h) * W3;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < C; c0 += kChunk) {
    const int kc = min(kChunk, C - c0);
    // 16 x 64 floats per operand, 256 threads -> 4 elements each, coalesced along w
    for (int e = tid; e < kChunk * kTile; e += 256) {
      const int k = e / kTile, w = e % kTile;
      float va = 0.f, vb = 0.f;
      if (k < kc) {
        if (m0 + w < W2) va = __ldg(pl + (long long)(c0 + k) * plane2 + m0 + w);
        if (n0 + w < W3) vb = __ldg(pr + (long long)(c0 + k) * plane3 + n0 + w);
      }
      sA[k][w] = va;
      sB[k][w] = vb;
    }
    __syncthreads();
    for (int k = 0; k < kc; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool vec = (W3 & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= W2) continue;
    float* orow = vol + (((long long)bh * W2) + m) * W3;
    const int n = n0 + tx * 4;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = div_const(acc[i][j], divisor, inv_divisor) * post_scale;
    if (vec && n + 3 < W3) {
      st_stream_v4(orow + n, make_float4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < W3) orow[n + j] = o[j];
    }
  }
}

}  // namespace sa

extern "C" int sa_corr_fp32(const float* fmap_l, const float* fmap_r, float* vol, int B, int C, int H, int W2,
                            int W3, float divisor, float post_scale, void* stream) {
  using namespace sa;
  SA_REQUIRE(fmap_l && fmap_r && vol, SA_E_INVALID, "sa_corr_fp32: null pointer");
  SA_REQUIRE(B > 0 && C > 0 && H > 0 && W2 > 0 && W3 > 0, SA_E_INVALID, "sa_corr_fp32: sizes must be positive");
  SA_REQUIRE(divisor != 0.f, SA_E_INVALID, "sa_corr_fp32: divisor == 0");
  SA_REQUIRE(aligned16(vol), SA_E_ALIGN, "sa_corr_fp32: vol must be 16-byte aligned");
  if (C == 3 && (W3 & 3) == 0 && aligned16(fmap_r))  // the mono volume: a streaming write, one warp per volume row
    return launch_mono_volume_rows(fmap_l, fmap_r, vol, B, H, W2, W3, divisor, post_scale, (cudaStream_t)stream);
  const int tiles_m = (W2 + kTile - 1) / kTile, tiles_n = (W3 + kTile - 1) / kTile;
  const long long blocks = (long long)B * H * tiles_m * tiles_n;
  SA_REQUIRE(blocks < (1ll << 31), SA_E_UNSUPPORTED, "sa_corr_fp32: too many tiles");
  corr_simt_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(fmap_l, fmap_r, vol, C, H, W2, W3, tiles_m,
                                                                      tiles_n, kernel_divisor(divisor), kernel_inv_divisor(divisor), post_scale);
  return finish_launch("sa_corr_fp32");
}
