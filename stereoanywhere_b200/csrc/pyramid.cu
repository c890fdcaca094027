// A3 - avg-pooled disparity pyramid (reference: models/stereoanywhere/corr.py:76-91) and
// A5 - truncation mask (reference: utils/utils.py:216-238, applied at stereoanywhere.py:253-255).
//
// One pass over level 0: every thread owns one aligned float4 of a volume row, forms two level-1
// values and one level-2 value in registers, and pairs up with its neighbour lane (shuffle) for
// the level-3 value.  With W % 8 == 0 (always true for the model: W = image width / 4, image
// width % 32 == 0) an aligned group of 8 columns never straddles a row, so the volume is treated
// as one flat float4 stream.  HBM-bound: reads W floats per row, writes 0.875 W (+ W when the
// truncation product is materialised as the block's level 0).
#include "sa_common.cuh"

namespace sa {

struct PyrArgs {
  const float* src;
  float* dst[3];
  long long pitch[3];
  long long src_pitch;
  long long rows;
  int W;
  int n_out;
  // truncation (optional)
  const float* disp;
  const float* conf;
  float gain, one_minus_gain;  // both rounded from the caller's double, like the reference's python floats
  int w2_size;
  float* masked0;
};

template <bool TRUNC, typename IdxT>
__global__ void __launch_bounds__(256) pyramid_vec_kernel(const PyrArgs a, const IdxT nvec) {
  const IdxT W4 = (IdxT)(a.W >> 2);
  const IdxT stride = (IdxT)gridDim.x * blockDim.x;
  // all lanes of a warp stay in the loop together (shuffle below): the bound is rounded up to a warp
  const IdxT first = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
  const IdxT nvec_up = (nvec + 31) & ~(IdxT)31;
  for (IdxT v = first; v < nvec_up; v += stride) {
    const bool ok = v < nvec;
    const IdxT rowi = ok ? v / W4 : 0;
    const int c4 = ok ? (int)(v - rowi * W4) : 0;
    const long long row = (long long)rowi;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) q = ld_stream_v4(a.src + row * a.src_pitch + 4 * c4);
    if (TRUNC) {
      if (ok) {
        const float c = __ldg(a.conf + row);
        const float centre = (float)(int)(rowi % (IdxT)a.w2_size) - __ldg(a.disp + row);
        const float omc = 1.0f - c, g = a.gain, omg = a.one_minus_gain;
        trunc_mask_mul4(q, centre, (float)(4 * c4), c, omc, g, omg);
        st_stream_v4(a.masked0 + row * (long long)a.W + 4 * c4, q);
      }
    }
    const float2 l1 = make_float2((q.x + q.y) * 0.5f, (q.z + q.w) * 0.5f);
    const float l2 = (l1.x + l1.y) * 0.5f;
    const float l2n = __shfl_down_sync(0xffffffffu, l2, 1);
    if (ok) {
      st_stream_v2(a.dst[0] + row * a.pitch[0] + 2 * c4, l1);
      if (a.n_out > 1) st_stream_f32(a.dst[1] + row * a.pitch[1] + c4, l2);
      if (a.n_out > 2 && !(c4 & 1)) st_stream_f32(a.dst[2] + row * a.pitch[2] + (c4 >> 1), (l2 + l2n) * 0.5f);
    }
  }
}

// Generic kernel: any W (odd tails are dropped level by level exactly like avg_pool2d with
// floor), any pitch / alignment.  One thread per level-(n_out) output column group: thread
// (row, j) owns level-0 columns [8j, 8j+8).
template <bool TRUNC>
__global__ void pyramid_generic_kernel(const PyrArgs a) {
  const int groups = (a.W + 7) / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * groups) return;
  const long long row = idx / groups;
  const int j = (int)(idx - row * groups);
  const int w0 = a.W, w1 = w0 / 2, w2 = w1 / 2, w3 = w2 / 2;
  float q[8];
  float c = 0.f, centre = 0.f;
  if (TRUNC) {
    c = a.conf[row];
    centre = (float)(int)(row % a.w2_size) - a.disp[row];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = 8 * j + e;
    float val = 0.f;
    if (col < w0) {
      val = a.src[row * a.src_pitch + col];
      if (TRUNC) {
        val *= trunc_mask(centre, (float)col, c, 1.0f - c, a.gain, a.one_minus_gain);
        a.masked0[row * (long long)a.W + col] = val;
      }
    }
    q[e] = val;
  }
  float l1[4], l2[2];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    l1[e] = (q[2 * e] + q[2 * e + 1]) * 0.5f;
    if (4 * j + e < w1) a.dst[0][row * a.pitch[0] + 4 * j + e] = l1[e];
  }
  if (a.n_out > 1) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      l2[e] = (l1[2 * e] + l1[2 * e + 1]) * 0.5f;
      if (2 * j + e < w2) a.dst[1][row * a.pitch[1] + 2 * j + e] = l2[e];
    }
    if (a.n_out > 2 && j < w3) a.dst[2][row * a.pitch[2] + j] = (l2[0] + l2[1]) * 0.5f;
  }
}

// Standalone A5: out = T * vol (or T itself when vol == nullptr).
__global__ void truncate_kernel(const float* vol, const float* disp, const float* conf, float gain, float omg,
                                float* out, long long rows, int w2_size, int W3) {
  const long long n = rows * W3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const long long row = idx / W3;
    const int col = (int)(idx - row * W3);
    const float c = __ldg(conf + row);
    const float centre = (float)(int)(row % w2_size) - __ldg(disp + row);
    const float t = trunc_mask(centre, (float)col, c, 1.0f - c, gain, omg);
    out[idx] = vol ? t * vol[idx] : t;
  }
}

}  // namespace sa

extern "C" int sa_pyramid(const float* src, int64_t rows, int W, int64_t src_pitch, int n_out, float* dst1,
                          float* dst2, float* dst3, int64_t pitch1, int64_t pitch2, int64_t pitch3,
                          const float* trunc_disp, const float* trunc_conf, double trunc_gain, int w2_size,
                          float* masked0, void* stream) {
  using namespace sa;
  SA_REQUIRE(src && rows > 0 && W >= 2, SA_E_INVALID, "sa_pyramid: need src, rows > 0, W >= 2");
  SA_REQUIRE(n_out >= 1 && n_out <= 3, SA_E_INVALID, "sa_pyramid: n_out must be 1..3");
  SA_REQUIRE(src_pitch >= W, SA_E_INVALID, "sa_pyramid: src_pitch < W");
  float* dst[3] = {dst1, dst2, dst3};
  const int64_t pitch[3] = {pitch1, pitch2, pitch3};
  int wl = W;
  for (int i = 0; i < n_out; ++i) {
    wl /= 2;
    SA_REQUIRE(wl >= 1, SA_E_INVALID, "sa_pyramid: level %d would be empty (W=%d)", i + 1, W);
    SA_REQUIRE(dst[i] != nullptr && pitch[i] >= wl, SA_E_INVALID, "sa_pyramid: bad dst / pitch for level %d", i + 1);
  }
  const bool trunc = trunc_disp != nullptr;
  if (trunc)
    SA_REQUIRE(trunc_conf && masked0 && w2_size > 0 && rows % w2_size == 0, SA_E_INVALID,
               "sa_pyramid: truncation needs conf, masked0 and rows %% w2_size == 0");
  PyrArgs a = {};
  a.src = src;
  a.rows = rows;
  a.W = W;
  a.src_pitch = src_pitch;
  a.n_out = n_out;
  for (int i = 0; i < 3; ++i) {
    a.dst[i] = dst[i];
    a.pitch[i] = pitch[i];
  }
  a.disp = trunc_disp;
  a.conf = trunc_conf;
  a.gain = (float)trunc_gain;
  a.one_minus_gain = (float)(1.0 - trunc_gain);
  a.w2_size = w2_size;
  a.masked0 = masked0;
  cudaStream_t st = (cudaStream_t)stream;

  bool vec = (W % 8 == 0) && (src_pitch % 4 == 0) && aligned16(src) && (!trunc || aligned16(masked0));
  for (int i = 0; i < n_out; ++i) vec = vec && (pitch[i] % 2 == 0) && ((reinterpret_cast<uintptr_t>(dst[i]) & 7u) == 0);
  if (vec) {
    const long long nvec = rows * (W / 4);
    const long long want = (nvec + 255) / 256;
    const int grid = (int)(want < (long long)num_sms() * 32 ? want : (long long)num_sms() * 32);
    const bool small = nvec < (1ll << 31);
    if (trunc && small)
      pyramid_vec_kernel<true, uint32_t><<<grid, 256, 0, st>>>(a, (uint32_t)nvec);
    else if (trunc)
      pyramid_vec_kernel<true, unsigned long long><<<grid, 256, 0, st>>>(a, (unsigned long long)nvec);
    else if (small)
      pyramid_vec_kernel<false, uint32_t><<<grid, 256, 0, st>>>(a, (uint32_t)nvec);
    else
      pyramid_vec_kernel<false, unsigned long long><<<grid, 256, 0, st>>>(a, (unsigned long long)nvec);
    return finish_launch("sa_pyramid (vec)");
  }
  const long long nthreads = rows * ((W + 7) / 8);
  const long long blocks = (nthreads + 255) / 256;
  SA_REQUIRE(blocks < (1ll << 31), SA_E_UNSUPPORTED, "sa_pyramid: volume too large for the generic kernel");
  if (trunc)
    pyramid_generic_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(a);
  else
    pyramid_generic_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(a);
  return finish_launch("sa_pyramid (generic)");
}

extern "C" int sa_truncate(const float* vol, const float* disp, const float* conf, double gain, float* out,
                           int64_t rows, int W2, int W3, void* stream) {
  using namespace sa;
  SA_REQUIRE(disp && conf && out && rows > 0 && W2 > 0 && W3 > 0 && rows % W2 == 0, SA_E_INVALID,
             "sa_truncate: bad arguments");
  if ((W3 & 3) == 0 && aligned16(out) && (!vol || aligned16(vol)))
    return launch_truncate_rows(vol, disp, conf, (float)gain, (float)(1.0 - gain), out, rows, W2, W3, (cudaStream_t)stream);
  const long long n = rows * (long long)W3;
  const long long want = (n + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 32 ? want : (long long)num_sms() * 32);
  truncate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vol, disp, conf, (float)gain, (float)(1.0 - gain), out, rows,
                                                          W2, W3);
  return finish_launch("sa_truncate");
}
