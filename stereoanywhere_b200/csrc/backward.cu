// backward of the lookup and of the pyramid (the adjoints of A4 and A3, optionally through the
// truncation product A5), so that the block can sit inside the reference's training step
// (train.py:277,383 - autograd through grid_sample / avg_pool2d / the mask product; coords are detached before
// every lookup, stereoanywhere.py:268, so only the VOLUME receives a gradient).
//
//   lookup   out[b, i*(2r+1)+k, h, w] = (1-f) P_i[row, x0+k-r] + f P_i[row, x0+k-r+1]   (corr.py:93-115)
//   adjoint  dP_i[row, x0-r+j] += (1-f) g[j] + f g[j-1],  j = 0 .. 2r+1   (g[-1] = g[2r+1] = 0; entries outside
//            [0, W_i) are dropped - they were zero padding)
// Every pixel (b,h,w) owns its own volume row, so the accumulation is a plain read-modify-write: no atomics,
// successive iterations are successive launches on the same stream.
//
//   pyramid  P_{i+1}[j] = 0.5 (P_i[2j] + P_i[2j+1]),  j < floor(W_i / 2)               (corr.py:88-91)
//   adjoint  dP_i[m] += 0.5 dP_{i+1}[m >> 1]  for m < 2 floor(W_i / 2);  finally dV = T * dP_0 when the block
//            was built from the truncation product T * V (T is detached in the reference, stereoanywhere.py:203).
#include "sa_common.cuh"

namespace sa {

struct LkBwdArgs {
  const float* grad_out;  // [B, L*(2r+1), H, W]
  const float* coords;
  long long coords_bstride;
  float* dlvl[SA_MAX_LEVELS];
  long long pitch[SA_MAX_LEVELS];
  int width[SA_MAX_LEVELS];
  int HW, num_levels, radius;
  float xoff;
  long long total;  // B * HW * num_levels work items
};

// one thread per (pixel, level); consecutive threads = consecutive pixels, so the grad_out reads are coalesced
template <int R>
__global__ void __launch_bounds__(256) lookup_backward_kernel(const LkBwdArgs a) {
  constexpr int NT = 2 * R + 1;
  const long long per_level = a.total / a.num_levels;  // B * HW
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < a.total; it += (long long)gridDim.x * blockDim.x) {
    const int lvl = (int)(it / per_level);
    const long long px = it - (long long)lvl * per_level;  // b * HW + hw
    const long long b = px / a.HW;
    const int hw = (int)(px - b * a.HW);
    const float x = __ldg(a.coords + b * a.coords_bstride + hw) + a.xoff;
    const float xs = x / (float)(1 << lvl);
    const float fl = fminf(fmaxf(floorf(xs), -1.0e6f), 1.0e6f);
    const float f = xs - floorf(xs);
    const int x0 = (int)fl - R;  // column of tap k = -R, left neighbour
    const float* g = a.grad_out + ((b * a.num_levels + lvl) * NT) * (long long)a.HW + hw;
    float gk[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) gk[k] = __ldg(g + (long long)k * a.HW);
    float* d = a.dlvl[lvl] + px * a.pitch[lvl];
    const int w = a.width[lvl];
#pragma unroll
    for (int j = 0; j <= NT; ++j) {
      const int c = x0 + j;
      if (c >= 0 && c < w) {
        const float add = (j < NT ? (1.0f - f) * gk[j < NT ? j : 0] : 0.f) + (j > 0 ? f * gk[j > 0 ? j - 1 : 0] : 0.f);
        d[c] += add;
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(256) lookup_backward_generic_kernel(const LkBwdArgs a, int nt) {
  const long long per_level = a.total / a.num_levels;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < a.total; it += (long long)gridDim.x * blockDim.x) {
    const int lvl = (int)(it / per_level);
    const long long px = it - (long long)lvl * per_level;
    const long long b = px / a.HW;
    const int hw = (int)(px - b * a.HW);
    const float x = __ldg(a.coords + b * a.coords_bstride + hw) + a.xoff;
    const float xs = x / (float)(1 << lvl);
    const float fl = fminf(fmaxf(floorf(xs), -1.0e6f), 1.0e6f);
    const float f = xs - floorf(xs);
    const int x0 = (int)fl - a.radius;
    const float* g = a.grad_out + ((b * a.num_levels + lvl) * nt) * (long long)a.HW + hw;
    float* d = a.dlvl[lvl] + px * a.pitch[lvl];
    const int w = a.width[lvl];
    float prev = 0.f;
    for (int j = 0; j <= nt; ++j) {
      const float cur = j < nt ? __ldg(g + (long long)j * a.HW) : 0.f;
      const int c = x0 + j;
      if (c >= 0 && c < w) d[c] += (1.0f - f) * cur + f * prev;
      prev = cur;
    }
  }
}

struct PyrBwdArgs {
  float* d0;  // [rows, pitch0]: in = dP_0, out = dV (in place)
  const float* dl[SA_MAX_LEVELS];  // dl[i] = dP_i for i >= 1
  long long pitch[SA_MAX_LEVELS];
  int width[SA_MAX_LEVELS];
  int num_levels;
  long long rows;
  const float* disp;  // truncation (optional)
  const float* conf;
  float gain, one_minus_gain;
  int w2_size;
};

// one thread per level-0 element: dV[row, m] = T * (dP0[m] + 0.5 (dP1[m>>1] + 0.5 (dP2[m>>2] + ...))) with every
// term present only while the parent index lies inside the pooled range of its level
template <bool TRUNC>
__global__ void __launch_bounds__(256) pyramid_backward_kernel(const PyrBwdArgs a) {
  const int W = a.width[0];
  const long long total = a.rows * W;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const long long row = it / W;
    const int m = (int)(it - row * W);
    // walk up while the index is covered by the next level, then fold back down
    float acc = 0.f;
    int top = 0;
    int idx = m;
    for (int i = 0; i + 1 < a.num_levels; ++i) {
      if (idx >= 2 * a.width[i + 1]) break;  // odd tail column of level i: not pooled
      idx >>= 1;
      top = i + 1;
    }
    for (int i = top; i >= 1; --i) {
      acc = 0.5f * (a.dl[i][row * a.pitch[i] + (m >> i)] + acc);
    }
    float v = a.d0[row * a.pitch[0] + m] + acc;
    if (TRUNC) {
      const float c = __ldg(a.conf + row);
      const float centre = (float)(int)(row % a.w2_size) - __ldg(a.disp + row);
      v *= trunc_mask(centre, (float)m, c, 1.0f - c, a.gain, a.one_minus_gain);
    }
    a.d0[row * a.pitch[0] + m] = v;
  }
}

}  // namespace sa

extern "C" int sa_lookup_backward(const float* grad_out, const float* coords, int64_t coords_bstride,
                                  float* const* h_dlevels, const int* h_widths, const int64_t* h_pitches, int num_levels,
                                  int radius, int B, int H, int W, int pad0, void* stream) {
  using namespace sa;
  SA_REQUIRE(grad_out && coords && h_dlevels && h_widths && h_pitches, SA_E_INVALID, "sa_lookup_backward: null pointer");
  SA_REQUIRE(num_levels >= 1 && num_levels <= SA_MAX_LEVELS && radius >= 0, SA_E_INVALID, "sa_lookup_backward: bad levels / radius");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), SA_E_INVALID, "sa_lookup_backward: bad sizes");
  LkBwdArgs a = {};
  a.grad_out = grad_out; a.coords = coords; a.coords_bstride = coords_bstride;
  for (int i = 0; i < num_levels; ++i) {
    SA_REQUIRE(h_dlevels[i] && h_widths[i] >= 1 && h_pitches[i] >= h_widths[i], SA_E_INVALID, "sa_lookup_backward: bad level %d", i);
    a.dlvl[i] = h_dlevels[i]; a.width[i] = h_widths[i]; a.pitch[i] = h_pitches[i];
  }
  a.HW = H * W; a.num_levels = num_levels; a.radius = radius; a.xoff = (float)pad0;
  a.total = (long long)B * a.HW * num_levels;
  const long long want = (a.total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 16 ? want : (long long)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (radius == 4) lookup_backward_kernel<4><<<grid, 256, 0, st>>>(a);
  else lookup_backward_generic_kernel<0><<<grid, 256, 0, st>>>(a, 2 * radius + 1);
  return finish_launch("sa_lookup_backward");
}

extern "C" int sa_pyramid_backward(float* d0, const float* const* h_dlevels, const int* h_widths, const int64_t* h_pitches,
                                   int num_levels, int64_t rows, const float* trunc_disp, const float* trunc_conf,
                                   double trunc_gain, int w2_size, void* stream) {
  using namespace sa;
  SA_REQUIRE(d0 && h_widths && h_pitches && rows > 0, SA_E_INVALID, "sa_pyramid_backward: null pointer / no rows");
  SA_REQUIRE(num_levels >= 1 && num_levels <= SA_MAX_LEVELS, SA_E_INVALID, "sa_pyramid_backward: bad level count");
  SA_REQUIRE((trunc_disp == nullptr) == (trunc_conf == nullptr), SA_E_INVALID, "sa_pyramid_backward: trunc_disp / trunc_conf must come together");
  PyrBwdArgs a = {};
  a.d0 = d0; a.num_levels = num_levels; a.rows = rows;
  for (int i = 0; i < num_levels; ++i) {
    a.width[i] = h_widths[i]; a.pitch[i] = h_pitches[i];
    if (i >= 1) {
      SA_REQUIRE(h_dlevels && h_dlevels[i], SA_E_INVALID, "sa_pyramid_backward: missing level %d", i);
      a.dl[i] = h_dlevels[i];
    }
  }
  a.disp = trunc_disp; a.conf = trunc_conf;
  a.gain = (float)trunc_gain; a.one_minus_gain = (float)(1.0 - trunc_gain); a.w2_size = w2_size;
  if (trunc_disp) SA_REQUIRE(w2_size > 0 && rows % w2_size == 0, SA_E_INVALID, "sa_pyramid_backward: rows %% w2_size != 0");
  const long long total = rows * h_widths[0];
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 16 ? want : (long long)num_sms() * 16);
  if (trunc_disp) pyramid_backward_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else pyramid_backward_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return finish_launch("sa_pyramid_backward");
}
