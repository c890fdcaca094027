// Producer chain of the path's inputs, two kernels:
//
//   sa_mono_inputs    mono depth [B,1,H,W] -> 1/2^n bilinear resize (align_corners), unit normals of the resized
//                     map and its one-hot depth bins, in ONE pass.  Reference: F.interpolate(mde, scale_factor=1/4,
//                     mode="bilinear", align_corners=True) (stereoanywhere.py:109-110), estimate_normals
//                     (utils/utils.py:73-77: -spatial_gradient(gain * d, "diff"), append 1, normalise) and
//                     generate_masks (utils/utils.py:48-54) - about forty small launches per forward there.
//   sa_weighted_lsq   per-sample scale / shift of the mono depth against the coarse disparity (weighted_lsq,
//                     utils/utils.py:345-384), one CTA per sample: the two quantiles of relu(disp) by an exact radix
//                     select (four 8-bit passes, all four order statistics at once), then the masked weighted normal
//                     equations in double and the 2 x 2 solve.  The reference loops over the samples in Python with
//                     two torch.quantile (a full sort each), three boolean-mask gathers (a device->host sync each)
//                     and a QR lstsq per sample.
#include <cuda_fp16.h>

#include "sa_common.cuh"

namespace sa {

struct MonoInArgs {
  const float* mde;
  int B, H, W, Hl, Wl;
  float rh, rw;        // (H - 1) / (Hl - 1), (W - 1) / (Wl - 1): align_corners source step
  float gain;
  int n_bins;
  float edges[SA_MAX_BINS + 1];
  float* lowres;       // [B,1,Hl,Wl]
  float* normals;      // [B,3,Hl,Wl]
  __half* masks;       // [B,n_bins,Hl,Wl] or null
};

// ATen's upsample_bilinear2d (align_corners): src = scale * dst, i0 = (int)src, i1 = i0 + (i0 < in - 1), lambda = src - i0
__device__ __forceinline__ float bilinear_ac(const float* img, int H, int W, float rh, float rw, int y, int x) {
  const float h1r = rh * (float)y, w1r = rw * (float)x;
  const int h1 = (int)h1r, w1 = (int)w1r;
  const int h1p = h1 < H - 1 ? 1 : 0, w1p = w1 < W - 1 ? 1 : 0;
  const float h1l = h1r - (float)h1, h0l = 1.0f - h1l;
  const float w1l = w1r - (float)w1, w0l = 1.0f - w1l;
  const float* p = img + (long long)h1 * W + w1;
  return h0l * (w0l * __ldg(p) + w1l * __ldg(p + w1p)) + h1l * (w0l * __ldg(p + (long long)h1p * W) + w1l * __ldg(p + (long long)h1p * W + w1p));
}

__global__ void __launch_bounds__(256) mono_inputs_kernel(const MonoInArgs a) {
  const long long n = (long long)a.B * a.Hl * a.Wl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.Wl);
    const long long r = i / a.Wl;
    const int y = (int)(r % a.Hl), b = (int)(r / a.Hl);
    const float* img = a.mde + (long long)b * a.H * a.W;
    auto lr = [&](int yy, int xx) {   // replicate padding of the gradient filter
      yy = min(max(yy, 0), a.Hl - 1);
      xx = min(max(xx, 0), a.Wl - 1);
      return bilinear_ac(img, a.H, a.W, a.rh, a.rw, yy, xx);
    };
    const float c = lr(y, x);
    const float gx = a.gain * lr(y, x + 1) - a.gain * lr(y, x - 1);
    const float gy = a.gain * lr(y + 1, x) - a.gain * lr(y - 1, x);
    const float nx = -gx, ny = -gy;
    const float norm = sqrtf(nx * nx + ny * ny + 1.0f);
    const long long plane = (long long)a.Hl * a.Wl, hw = (long long)y * a.Wl + x;
    a.lowres[(long long)b * plane + hw] = c;
    float* np_ = a.normals + (long long)b * 3 * plane + hw;
    np_[0] = nx / norm;
    np_[plane] = ny / norm;
    np_[2 * plane] = 1.0f / norm;
    if (a.masks) {
      __half* mp = a.masks + (long long)b * a.n_bins * plane + hw;
      for (int k = 0; k < a.n_bins; ++k) mp[k * plane] = __float2half((c < a.edges[k + 1] && c >= a.edges[k]) ? 1.0f : 0.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------------ weighted_lsq
struct LsqArgs {
  const float* mono;   // [B, n]
  const float* disp;   // [B, n]
  const float* conf;   // [B, n]
  int n;
  float qmin, qmax;
  float* scale;        // [B]
  float* shift;        // [B]
};

constexpr int kLsqThreads = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = threadIdx.x < (kLsqThreads >> 5) ? sh[threadIdx.x] : 0.0;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

__global__ void __launch_bounds__(kLsqThreads) weighted_lsq_kernel(const LsqArgs a) {
  __shared__ unsigned hist[4][256];
  __shared__ unsigned prefix[4], want[4];
  __shared__ double red[32];
  const int b = blockIdx.x, n = a.n;
  const float* disp = a.disp + (long long)b * n;
  // torch.quantile(x, q) (linear): rank = q * (n - 1) in float32; value = lerp(sorted[floor], sorted[ceil], frac)
  const float r0 = a.qmin * (float)(n - 1), r1 = a.qmax * (float)(n - 1);
  const unsigned lo0 = (unsigned)floorf(r0), lo1 = (unsigned)floorf(r1);
  if (threadIdx.x == 0) {
    want[0] = lo0; want[1] = min(lo0 + 1u, (unsigned)n - 1u);
    want[2] = lo1; want[3] = min(lo1 + 1u, (unsigned)n - 1u);
    for (int k = 0; k < 4; ++k) prefix[k] = 0u;
  }
  // exact k-th smallest of relu(disp) for the four ranks: relu makes every value >= +0, whose float bits order as
  // unsigned integers.  8 bits per pass, most significant first; `prefix[k]` is the matched high part of rank k.
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const unsigned himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    unsigned pf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pf[k] = prefix[k];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned u = __float_as_uint(fmaxf(__ldg(disp + i), 0.0f));
      const unsigned digit = (u >> shift) & 255u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((u & himask) == pf[k]) atomicAdd(&hist[k][digit], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      const int k = threadIdx.x;
      unsigned rank = want[k], d = 0;
      for (; d < 256; ++d) {
        const unsigned c = hist[k][d];
        if (rank < c) break;
        rank -= c;
      }
      want[k] = rank;                       // rank among the elements sharing the longer prefix
      prefix[k] |= d << shift;
    }
    __syncthreads();
  }
  const float s_lo0 = __uint_as_float(prefix[0]), s_hi0 = __uint_as_float(prefix[1]);
  const float s_lo1 = __uint_as_float(prefix[2]), s_hi1 = __uint_as_float(prefix[3]);
  auto lerp = [](float x, float y, float w) { return w < 0.5f ? x + w * (y - x) : y - (y - x) * (1.0f - w); };   // at::lerp
  const float qlo = lerp(s_lo0, s_hi0, r0 - floorf(r0)), qhi = lerp(s_lo1, s_hi1, r1 - floorf(r1));

  const float* mono = a.mono + (long long)b * n;
  const float* conf = a.conf + (long long)b * n;
  double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = fmaxf(__ldg(disp + i), 0.0f);
    if (qlo <= s && s <= qhi) {
      const double m = (double)fabsf(__ldg(mono + i));
      const double w2 = (double)(fabsf(__ldg(conf + i)) * (1.0f - 0.1f) + 0.1f);   // (sqrt(conf'))^2
      const double sd = (double)s;
      a00 += w2 * m * m; a01 += w2 * m; a11 += w2; b0 += w2 * m * sd; b1 += w2 * sd;
    }
  }
  a00 = block_sum(a00, red); a01 = block_sum(a01, red); a11 = block_sum(a11, red);
  b0 = block_sum(b0, red); b1 = block_sum(b1, red);
  if (threadIdx.x == 0) {
    const double det = a00 * a11 - a01 * a01;
    a.scale[b] = (float)((a11 * b0 - a01 * b1) / det);
    a.shift[b] = (float)((a00 * b1 - a01 * b0) / det);
  }
}

}  // namespace sa

extern "C" int sa_mono_inputs(const float* mde, int B, int H, int W, int n_downsample, float normal_gain,
                              const float* h_edges, int n_bins, float* lowres, float* normals, void* masks_f16,
                              void* stream) {
  using namespace sa;
  SA_REQUIRE(mde && lowres && normals && B > 0 && H > 0 && W > 0, SA_E_INVALID, "sa_mono_inputs: null pointer / bad sizes");
  SA_REQUIRE(n_downsample >= 0 && n_downsample <= 6, SA_E_INVALID, "sa_mono_inputs: n_downsample out of range");
  SA_REQUIRE(!masks_f16 || (h_edges && n_bins >= 1 && n_bins <= SA_MAX_BINS), SA_E_INVALID,
             "sa_mono_inputs: masks need 1..%d bins and their edges", SA_MAX_BINS);
  MonoInArgs a = {};
  a.mde = mde; a.B = B; a.H = H; a.W = W;
  // F.interpolate(scale_factor = 1 / 2^n): output size = floor(in * scale)
  a.Hl = (int)((double)H * (1.0 / (double)(1 << n_downsample)));
  a.Wl = (int)((double)W * (1.0 / (double)(1 << n_downsample)));
  SA_REQUIRE(a.Hl >= 1 && a.Wl >= 1, SA_E_INVALID, "sa_mono_inputs: image too small for the resize");
  a.rh = a.Hl > 1 ? (float)(H - 1) / (float)(a.Hl - 1) : 0.f;
  a.rw = a.Wl > 1 ? (float)(W - 1) / (float)(a.Wl - 1) : 0.f;
  a.gain = normal_gain; a.n_bins = masks_f16 ? n_bins : 0;
  if (masks_f16)
    for (int i = 0; i <= n_bins; ++i) a.edges[i] = h_edges[i];
  a.lowres = lowres; a.normals = normals; a.masks = reinterpret_cast<__half*>(masks_f16);
  const long long n = (long long)B * a.Hl * a.Wl;
  const long long want = (n + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  mono_inputs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return finish_launch("sa_mono_inputs");
}

extern "C" int sa_weighted_lsq(const float* mono, const float* disp, const float* conf, int B, int n, float min_quantile,
                               float max_quantile, float* scale, float* shift, void* stream) {
  using namespace sa;
  SA_REQUIRE(mono && disp && conf && scale && shift && B > 0 && n > 1, SA_E_INVALID, "sa_weighted_lsq: null pointer / bad sizes");
  SA_REQUIRE(min_quantile >= 0.f && max_quantile <= 1.f && min_quantile <= max_quantile, SA_E_INVALID,
             "sa_weighted_lsq: quantiles must satisfy 0 <= min <= max <= 1");
  LsqArgs a = {mono, disp, conf, n, min_quantile, max_quantile, scale, shift};
  weighted_lsq_kernel<<<B, kLsqThreads, 0, (cudaStream_t)stream>>>(a);
  return finish_launch("sa_weighted_lsq");
}
