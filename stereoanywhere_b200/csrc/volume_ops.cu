// A6 - depth-bin masked volume (reference: utils/utils.py:48-54 `generate_masks`,
//      stereoanywhere.py:138-139,161) and
// A7 - training-only volume corruption (reference: stereoanywhere.py:214-251,
//      utils/utils.py:200-214 `gauss_corr_volume_naive`).
// Both are full-volume streaming passes; every thread owns one float4 of a volume row.
#include "sa_common.cuh"

namespace sa {

struct BinEdges {
  float e[SA_MAX_BINS + 1];
};

// bin(x) = n iff e[n] <= x < e[n+1]; -1 when x is in no bin (x == 1.0, NaN, out of range).
__device__ __forceinline__ int depth_bin(float x, const BinEdges& ed, int n_bins) {
  int bin = -1;
  for (int n = 0; n < n_bins; ++n)
    if (x >= ed.e[n] && x < ed.e[n + 1]) bin = n;
  return bin;
}

template <bool FROM_NORMALS>
__global__ void __launch_bounds__(256)
masked_volume_kernel(const float* __restrict__ vol, const float* __restrict__ nl, const float* __restrict__ nr,
                     float divisor, float inv_divisor, float post_scale, const float* __restrict__ mde_l,
                     const float* __restrict__ mde_r, const BinEdges ed, int n_bins, float* __restrict__ out,
                     int H, int W2, int W3, long long nvec) {
  const int W34 = (W3 + 3) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long hw2 = (long long)H * W2;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const long long row = v / W34;            // (b*H + h)*W2 + w2
    const int col = (int)(v - row * W34) * 4;  // w3
    const long long b = row / hw2;
    const long long rem = row - b * hw2;       // h*W2 + w2
    const long long h = rem / W2;
    const int bin_l = depth_bin(__ldg(mde_l + row), ed, n_bins);
    const float* mr = mde_r + (b * H + h) * W3 + col;
    float val[4];
    int bin_r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool in = col + e < W3;
      bin_r[e] = in ? depth_bin(__ldg(mr + e), ed, n_bins) : -1;
      float x = 0.f;
      if (in) {
        if (FROM_NORMALS) {
          const int w2 = (int)(rem - h * W2);
          const float* pl = nl + ((b * 3) * H + h) * W2 + w2;
          const float* pr = nr + ((b * 3) * H + h) * W3 + col + e;
          float acc = __ldg(pl) * __ldg(pr);
          acc = fmaf(__ldg(pl + hw2), __ldg(pr + (long long)H * W3), acc);
          acc = fmaf(__ldg(pl + 2 * hw2), __ldg(pr + 2ll * H * W3), acc);
          x = div_const(acc, divisor, inv_divisor) * post_scale;
        } else {
          x = ld_stream_f32(vol + row * W3 + col + e);
        }
      }
      val[e] = x;
    }
    const bool vec_store = ((W3 & 3) == 0);
    for (int n = 0; n < n_bins; ++n) {
      float* o = out + ((b * n_bins + n) * hw2 + rem) * W3 + col;
      float w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = (bin_l == n && bin_r[e] == n) ? val[e] : val[e] * 0.0f;
      if (vec_store) {
        st_stream_v4(o, make_float4(w[0], w[1], w[2], w[3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < W3) o[e] = w[e];
      }
    }
  }
}

// mode 0 roll / 1 noise / 2 gauss; one thread per element (training-only, not on any bench config)
__global__ void corrupt_kernel(const float* __restrict__ vol, const float* __restrict__ bin_mask, int mode, int shift,
                               const float* __restrict__ noise, float gauss_k, float* __restrict__ out, int W2,
                               int W3, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const long long row = idx / W3;
    const int w3 = (int)(idx - row * W3);
    const int w2 = (int)(row % W2);
    const float m = __ldg(bin_mask + row);
    const float x = vol[idx];
    float other;
    if (mode == 0) {
      int src = (w2 - shift) % W2;
      if (src < 0) src += W2;
      other = vol[(row - w2 + src) * W3 + w3];
    } else if (mode == 1) {
      other = x * __ldg(noise + row);
    } else {
      const float d = (float)w2 - (float)w3;
      other = x * (gauss_k * expf(-(d * d) / 2.0f));
    }
    out[idx] = x * (1.0f - m) + other * m;
  }
}

}  // namespace sa

extern "C" int sa_masked_volume(const float* vol, const float* normals_l, const float* normals_r, float divisor,
                                float post_scale, const float* mde_l, const float* mde_r, const float* h_edges,
                                int n_bins, float* out, int B, int H, int W2, int W3, void* stream) {
  using namespace sa;
  SA_REQUIRE(mde_l && mde_r && h_edges && out, SA_E_INVALID, "sa_masked_volume: null pointer");
  SA_REQUIRE(vol || (normals_l && normals_r && divisor != 0.f), SA_E_INVALID,
             "sa_masked_volume: need vol or both normal maps");
  SA_REQUIRE(n_bins >= 1 && n_bins <= SA_MAX_BINS, SA_E_INVALID, "sa_masked_volume: 1 <= n_bins <= %d", SA_MAX_BINS);
  SA_REQUIRE(B > 0 && H > 0 && W2 > 0 && W3 > 0, SA_E_INVALID, "sa_masked_volume: sizes must be positive");
  SA_REQUIRE((W3 & 3) != 0 || aligned16(out), SA_E_ALIGN, "sa_masked_volume: out must be 16-byte aligned");
  if ((W3 & 3) == 0 && aligned16(out) && aligned16(mde_r) && (vol ? aligned16(vol) : aligned16(normals_r)))
    return launch_masked_volume_rows(vol, normals_l, normals_r, divisor, post_scale, mde_l, mde_r, h_edges, n_bins, out, B, H,
                                     W2, W3, (cudaStream_t)stream);
  BinEdges ed;
  for (int i = 0; i <= n_bins; ++i) ed.e[i] = h_edges[i];
  for (int i = n_bins + 1; i <= SA_MAX_BINS; ++i) ed.e[i] = 0.f;
  const long long nvec = (long long)B * H * W2 * ((W3 + 3) / 4);
  const long long want = (nvec + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 16 ? want : (long long)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (vol)
    masked_volume_kernel<false><<<grid, 256, 0, st>>>(vol, nullptr, nullptr, 1.f, 1.f, 1.f, mde_l, mde_r, ed, n_bins, out,
                                                      H, W2, W3, nvec);
  else
    masked_volume_kernel<true><<<grid, 256, 0, st>>>(nullptr, normals_l, normals_r, kernel_divisor(divisor), kernel_inv_divisor(divisor),
                                                     post_scale, mde_l,
                                                     mde_r, ed, n_bins, out, H, W2, W3, nvec);
  return finish_launch("sa_masked_volume");
}

extern "C" int sa_corrupt(const float* vol, const float* bin_mask, int mode, int shift, const float* noise,
                          float gauss_k, float* out, int B, int H, int W2, int W3, void* stream) {
  using namespace sa;
  SA_REQUIRE(vol && bin_mask && out && vol != out, SA_E_INVALID, "sa_corrupt: null pointer or in-place call");
  SA_REQUIRE(mode >= 0 && mode <= 2, SA_E_INVALID, "sa_corrupt: mode must be 0 (roll), 1 (noise) or 2 (gauss)");
  SA_REQUIRE(mode != 1 || noise, SA_E_INVALID, "sa_corrupt: noise mode needs a noise map");
  SA_REQUIRE(B > 0 && H > 0 && W2 > 0 && W3 > 0, SA_E_INVALID, "sa_corrupt: sizes must be positive");
  if ((W3 & 3) == 0 && aligned16(vol) && aligned16(out))
    return launch_corrupt_rows(vol, bin_mask, mode, shift, noise, gauss_k, out, (long long)B * H * W2, W2, W3,
                               (cudaStream_t)stream);
  const long long n = (long long)B * H * W2 * W3;
  const long long want = (n + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 32 ? want : (long long)num_sms() * 32);
  corrupt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vol, bin_mask, mode, shift, noise, gauss_k, out, W2, W3, n);
  return finish_launch("sa_corrupt");
}
