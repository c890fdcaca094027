// Row-streaming forms of the full-volume passes A2 (C = 3 mono volume), A5 (truncation mask / product),
// A6 (depth-bin masked volume) and A7 (training corruption) for W3 % 4 == 0 - the model's shapes.
//   reference: stereoanywhere.py:136 (1.73 * corr(nL, nR)), utils/utils.py:216-238 + stereoanywhere.py:253-255,
//              utils/utils.py:48-54 + stereoanywhere.py:138-139,161, stereoanywhere.py:214-251.
// All four are out[row, :] = f(row constants, column constants of the image row (b,h), vol[row, :]).  A warp owns
// a CONTIGUOUS chunk of volume rows, so everything that depends on (b,h) only - the right normals, the right
// depth bins - is loaded when (b,h) changes and otherwise stays in registers (W3 <= 384), and the per-row
// constants cost three or four scalar loads; no division or 64-bit multiply per element (the first versions
// of these kernels spent their time there: 0.2-0.5 of the HBM roofline).  The generic kernels in
// volume_ops.cu / pyramid.cu / corr_simt.cu remain for other shapes.
#include "sa_common.cuh"

namespace sa {

constexpr int kRowsMaxV4 = 3;  // float4 column groups per lane kept in registers (W3 <= 384)
constexpr int kRowWarps = 8;

struct RowChunk {
  long long row, row_end;
};
__device__ __forceinline__ RowChunk my_rows(long long rows) {
  const long long warps_total = (long long)gridDim.x * kRowWarps;
  const long long gwarp = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const long long per = (rows + warps_total - 1) / warps_total;
  RowChunk c;
  c.row = gwarp * per;
  c.row_end = min(rows, c.row + per);
  return c;
}

// position of a volume row inside [B][H][W2], advanced without divisions (rows of a chunk are consecutive)
struct RowPos {
  long long bh;
  int b, h, w2;
  __device__ __forceinline__ void init(long long row, int H, int W2) {
    bh = row / W2;
    w2 = (int)(row - bh * W2);
    b = (int)(bh / H);
    h = (int)(bh - (long long)b * H);
  }
  __device__ __forceinline__ void next(int H, int W2) {
    if (++w2 == W2) {
      w2 = 0;
      ++bh;
      if (++h == H) { h = 0; ++b; }
    }
  }
};

// ------------------------------------------------------------------------------------------------ A2, C = 3
__global__ void __launch_bounds__(kRowWarps * 32)
mono_volume_rows_kernel(const float* __restrict__ nl, const float* __restrict__ nr, float* __restrict__ out, int H, int W2,
                        int W3, long long rows, float divisor, float inv_divisor, float post_scale) {
  const int lane = threadIdx.x & 31, W4 = W3 >> 2;
  const bool resident = W4 <= 32 * kRowsMaxV4;
  const long long plane2 = (long long)H * W2, plane3 = (long long)H * W3;
  RowChunk rc = my_rows(rows);
  float4 r0[kRowsMaxV4], r1[kRowsMaxV4], r2[kRowsMaxV4];
  long long cur_bh = -1;
  // the three left-normal scalars of the NEXT row are loaded while this row is written (software pipeline)
  auto left_normal = [&](const RowPos& q, float& a0, float& a1, float& a2) {
    const float* nlp = nl + ((long long)q.b * 3 * H + q.h) * W2 + q.w2;
    a0 = __ldg(nlp); a1 = __ldg(nlp + plane2); a2 = __ldg(nlp + 2 * plane2);
  };
  float p0 = 0.f, p1 = 0.f, p2 = 0.f;
  RowPos pos, nxp;
  pos.init(rc.row < rc.row_end ? rc.row : 0, H, W2);
  nxp = pos;
  if (rc.row < rc.row_end) left_normal(pos, p0, p1, p2);
  for (long long row = rc.row; row < rc.row_end; ++row, pos.next(H, W2)) {
    const long long bh = pos.bh;
    const long long b = pos.b, h = pos.h;
    const float n0 = p0, n1 = p1, n2 = p2;
    nxp.next(H, W2);
    if (row + 1 < rc.row_end) left_normal(nxp, p0, p1, p2);
    const float* nrp = nr + (b * 3 * H + h) * (long long)W3;
    if (resident && bh != cur_bh) {
      cur_bh = bh;
#pragma unroll
      for (int i = 0; i < kRowsMaxV4; ++i) {
        const int v = lane + 32 * i;
        if (v < W4) {
          r0[i] = __ldg(reinterpret_cast<const float4*>(nrp + 4 * v));
          r1[i] = __ldg(reinterpret_cast<const float4*>(nrp + plane3 + 4 * v));
          r2[i] = __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + 4 * v));
        }
      }
    }
    // same FMA order as corr_simt_kernel (c = 0, 1, 2 from a zero accumulator): bit-identical to it
    auto mono4 = [&](const float4& a0, const float4& a1, const float4& a2) {
      float4 q;
      q.x = div_const(fmaf(n2, a2.x, fmaf(n1, a1.x, fmaf(n0, a0.x, 0.f))), divisor, inv_divisor) * post_scale;
      q.y = div_const(fmaf(n2, a2.y, fmaf(n1, a1.y, fmaf(n0, a0.y, 0.f))), divisor, inv_divisor) * post_scale;
      q.z = div_const(fmaf(n2, a2.z, fmaf(n1, a1.z, fmaf(n0, a0.z, 0.f))), divisor, inv_divisor) * post_scale;
      q.w = div_const(fmaf(n2, a2.w, fmaf(n1, a1.w, fmaf(n0, a0.w, 0.f))), divisor, inv_divisor) * post_scale;
      return q;
    };
    float* o = out + row * W3;
    if (resident) {
#pragma unroll
      for (int i = 0; i < kRowsMaxV4; ++i) {
        const int v = lane + 32 * i;
        if (v < W4) st_stream_v4(o + 4 * v, mono4(r0[i], r1[i], r2[i]));
      }
    } else {
      for (int v = lane; v < W4; v += 32)
        st_stream_v4(o + 4 * v, mono4(__ldg(reinterpret_cast<const float4*>(nrp + 4 * v)),
                                      __ldg(reinterpret_cast<const float4*>(nrp + plane3 + 4 * v)),
                                      __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + 4 * v))));
    }
  }
}

// ------------------------------------------------------------------------------------------------ A5
template <bool HAS_VOL>
__global__ void __launch_bounds__(kRowWarps * 32)
truncate_rows_kernel(const float* __restrict__ vol, const float* __restrict__ disp, const float* __restrict__ conf, float gain,
                     float omg, float* __restrict__ out, long long rows, int w2_size, int W3) {
  const int lane = threadIdx.x & 31, W4 = W3 >> 2;
  RowChunk rc = my_rows(rows);
  const bool resident = W4 <= 32 * kRowsMaxV4;
  float pc = 0.f, pd = 0.f;
  float4 nxt[kRowsMaxV4];
  auto fetch = [&](long long row) {
    pc = __ldg(conf + row);
    pd = __ldg(disp + row);
    if (HAS_VOL && resident) {
#pragma unroll
      for (int i = 0; i < kRowsMaxV4; ++i)
        if (lane + 32 * i < W4) nxt[i] = ld_stream_v4(vol + row * W3 + 4 * (lane + 32 * i));
    }
  };
  if (rc.row < rc.row_end) fetch(rc.row);
  int w2 = (int)((rc.row < rc.row_end ? rc.row : 0) % w2_size);
  for (long long row = rc.row; row < rc.row_end; ++row, w2 = (w2 + 1 == w2_size ? 0 : w2 + 1)) {
    const float c = pc;
    const float centre = (float)w2 - pd;
    const float omc = 1.0f - c;
    float4 cur[kRowsMaxV4];
#pragma unroll
    for (int i = 0; i < kRowsMaxV4; ++i) cur[i] = nxt[i];
    if (row + 1 < rc.row_end) fetch(row + 1);  // the next row's loads fly while this row is written
    if (resident) {
#pragma unroll
      for (int i = 0; i < kRowsMaxV4; ++i) {
        const int v = lane + 32 * i;
        if (v < W4) {
          float4 q = HAS_VOL ? cur[i] : make_float4(1.f, 1.f, 1.f, 1.f);
          trunc_mask_mul4(q, centre, (float)(4 * v), c, omc, gain, omg);  // x 1.0 is exact: the mask itself
          st_stream_v4(out + row * W3 + 4 * v, q);
        }
      }
    } else {
      for (int v = lane; v < W4; v += 32) {
        float4 q = HAS_VOL ? ld_stream_v4(vol + row * W3 + 4 * v) : make_float4(1.f, 1.f, 1.f, 1.f);
        trunc_mask_mul4(q, centre, (float)(4 * v), c, omc, gain, omg);
        st_stream_v4(out + row * W3 + 4 * v, q);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ A6
struct RowBinEdges {
  float e[SA_MAX_BINS + 1];
};
__device__ __forceinline__ int row_depth_bin(float x, const RowBinEdges& ed, int n_bins) {
  int bin = -1;
  for (int n = 0; n < n_bins; ++n)
    if (x >= ed.e[n] && x < ed.e[n + 1]) bin = n;
  return bin;
}

template <bool FROM_NORMALS>
__global__ void __launch_bounds__(kRowWarps * 32)
masked_volume_rows_kernel(const float* __restrict__ vol, const float* __restrict__ nl, const float* __restrict__ nr, float divisor,
                          float inv_divisor, float post_scale, const float* __restrict__ mde_l,
                          const float* __restrict__ mde_r, const RowBinEdges ed, int n_bins, float* __restrict__ out, int H,
                          int W2, int W3, long long rows) {
  const int lane = threadIdx.x & 31, W4 = W3 >> 2;
  const long long plane2 = (long long)H * W2, plane3 = (long long)H * W3;
  RowChunk rc = my_rows(rows);
  // per (b,h): bins of the right pixels of this lane's columns (4 x int8 per group) and, from normals, nR
  uint32_t rb[kRowsMaxV4];
  float4 r0[kRowsMaxV4], r1[kRowsMaxV4], r2[kRowsMaxV4];
  long long cur_bh = -1;
  const bool resident = W4 <= 32 * kRowsMaxV4;
  float4 nxt[kRowsMaxV4];  // volume source: the next row's values, loaded while this row's N planes are written
  auto fetch = [&](long long row) {
#pragma unroll
    for (int i = 0; i < kRowsMaxV4; ++i)
      if (lane + 32 * i < W4) nxt[i] = ld_stream_v4(vol + row * W3 + 4 * (lane + 32 * i));
  };
  if (!FROM_NORMALS && resident && rc.row < rc.row_end) fetch(rc.row);
  RowPos pos;
  for (long long row = rc.row; row < rc.row_end; ++row) {
    float4 cur[kRowsMaxV4];
#pragma unroll
    for (int i = 0; i < kRowsMaxV4; ++i) cur[i] = nxt[i];
    if (!FROM_NORMALS && resident && row + 1 < rc.row_end) fetch(row + 1);
    if (row == rc.row) pos.init(row, H, W2); else pos.next(H, W2);
    const long long bh = pos.bh;
    const int w2 = pos.w2;
    const long long b = pos.b, h = pos.h;
    const int bin_l = row_depth_bin(__ldg(mde_l + row), ed, n_bins);
    const float* mrp = mde_r + bh * W3;
    const float* nrp = FROM_NORMALS ? nr + (b * 3 * H + h) * (long long)W3 : nullptr;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (FROM_NORMALS) {
      const float* nlp = nl + (b * 3 * H + h) * W2 + w2;
      n0 = __ldg(nlp); n1 = __ldg(nlp + plane2); n2 = __ldg(nlp + 2 * plane2);
    }
    auto bins4 = [&](int v) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mrp + 4 * v));
      return (uint32_t)(row_depth_bin(m.x, ed, n_bins) & 0xff) | ((uint32_t)(row_depth_bin(m.y, ed, n_bins) & 0xff) << 8) |
             ((uint32_t)(row_depth_bin(m.z, ed, n_bins) & 0xff) << 16) | ((uint32_t)(row_depth_bin(m.w, ed, n_bins) & 0xff) << 24);
    };
    if (resident && bh != cur_bh) {
      cur_bh = bh;
#pragma unroll
      for (int i = 0; i < kRowsMaxV4; ++i) {
        const int v = lane + 32 * i;
        if (v < W4) {
          rb[i] = bins4(v);
          if (FROM_NORMALS) {
            r0[i] = __ldg(reinterpret_cast<const float4*>(nrp + 4 * v));
            r1[i] = __ldg(reinterpret_cast<const float4*>(nrp + plane3 + 4 * v));
            r2[i] = __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + 4 * v));
          }
        }
      }
    }
    const long long rem = row - b * plane2;  // h * W2 + w2
    float* obase = out + (b * n_bins * plane2 + rem) * W3;
    const long long nstride = plane2 * W3;
    auto emit = [&](int v, uint32_t bins, float4 val) {
      // the reference multiplies by 0/1 masks: entries outside the bin are val * 0 (keeps the sign of zero)
      const float4 z = make_float4(val.x * 0.0f, val.y * 0.0f, val.z * 0.0f, val.w * 0.0f);
      for (int n = 0; n < n_bins; ++n) {
        float4 w = z;
        if (n == bin_l) {
          const uint32_t nn = (uint32_t)n;
          if ((bins & 0xff) == nn) w.x = val.x;
          if (((bins >> 8) & 0xff) == nn) w.y = val.y;
          if (((bins >> 16) & 0xff) == nn) w.z = val.z;
          if ((bins >> 24) == nn) w.w = val.w;
        }
        st_stream_v4(obase + n * nstride + 4 * v, w);
      }
    };
    auto value4 = [&](int v, const float4& a0, const float4& a1, const float4& a2) {
      if (!FROM_NORMALS) return ld_stream_v4(vol + row * W3 + 4 * v);
      float4 q;
      q.x = div_const(fmaf(n2, a2.x, fmaf(n1, a1.x, n0 * a0.x)), divisor, inv_divisor) * post_scale;
      q.y = div_const(fmaf(n2, a2.y, fmaf(n1, a1.y, n0 * a0.y)), divisor, inv_divisor) * post_scale;
      q.z = div_const(fmaf(n2, a2.z, fmaf(n1, a1.z, n0 * a0.z)), divisor, inv_divisor) * post_scale;
      q.w = div_const(fmaf(n2, a2.w, fmaf(n1, a1.w, n0 * a0.w)), divisor, inv_divisor) * post_scale;
      return q;
    };
    if (resident) {
#pragma unroll
      for (int i = 0; i < kRowsMaxV4; ++i) {
        const int v = lane + 32 * i;
        if (v < W4) emit(v, rb[i], FROM_NORMALS ? value4(v, r0[i], r1[i], r2[i]) : cur[i]);
      }
    } else {
      for (int v = lane; v < W4; v += 32) {
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
        if (FROM_NORMALS) {
          a0 = __ldg(reinterpret_cast<const float4*>(nrp + 4 * v));
          a1 = __ldg(reinterpret_cast<const float4*>(nrp + plane3 + 4 * v));
          a2 = __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + 4 * v));
        }
        emit(v, bins4(v), value4(v, a0, a1, a2));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ A7
__global__ void __launch_bounds__(kRowWarps * 32)
corrupt_rows_kernel(const float* __restrict__ vol, const float* __restrict__ bin_mask, int mode, int shift,
                    const float* __restrict__ noise, float gauss_k, float* __restrict__ out, int W2, int W3, long long rows) {
  const int lane = threadIdx.x & 31, W4 = W3 >> 2;
  RowChunk rc = my_rows(rows);
  int w2 = (int)((rc.row < rc.row_end ? rc.row : 0) % W2);
  for (long long row = rc.row; row < rc.row_end; ++row, w2 = (w2 + 1 == W2 ? 0 : w2 + 1)) {
    const float m = __ldg(bin_mask + row), omm = 1.0f - m;
    const float* src = vol + row * W3;
    const float* other_row = src;
    float nz = 1.f;
    if (mode == 0) {
      int s = (w2 - shift) % W2;
      if (s < 0) s += W2;
      other_row = vol + (row - w2 + s) * W3;
    } else if (mode == 1) {
      nz = __ldg(noise + row);
    }
    for (int v = lane; v < W4; v += 32) {
      const float4 x = ld_stream_v4(src + 4 * v);
      float4 o;
      if (mode == 0) {
        o = __ldg(reinterpret_cast<const float4*>(other_row + 4 * v));
      } else if (mode == 1) {
        o = make_float4(x.x * nz, x.y * nz, x.z * nz, x.w * nz);
      } else {
        const float d0 = (float)w2 - (float)(4 * v);
        const float d1 = d0 - 1.f, d2 = d0 - 2.f, d3 = d0 - 3.f;
        o = make_float4(x.x * (gauss_k * expf(-(d0 * d0) / 2.0f)), x.y * (gauss_k * expf(-(d1 * d1) / 2.0f)),
                        x.z * (gauss_k * expf(-(d2 * d2) / 2.0f)), x.w * (gauss_k * expf(-(d3 * d3) / 2.0f)));
      }
      st_stream_v4(out + row * W3 + 4 * v,
                   make_float4(x.x * omm + o.x * m, x.y * omm + o.y * m, x.z * omm + o.z * m, x.w * omm + o.w * m));
    }
  }
}

static int rows_grid(long long rows) {
  const long long want = (rows + kRowWarps - 1) / kRowWarps;
  const long long cap = (long long)num_sms() * 8;
  return (int)(want < cap ? want : cap);
}

// ---- launchers used by the C entry points (declared in sa_rows.h style: plain functions inside namespace sa)
int launch_mono_volume_rows(const float* nl, const float* nr, float* out, int B, int H, int W2, int W3, float divisor,
                            float post_scale, cudaStream_t st) {
  const long long rows = (long long)B * H * W2;
  mono_volume_rows_kernel<<<rows_grid(rows), kRowWarps * 32, 0, st>>>(nl, nr, out, H, W2, W3, rows, kernel_divisor(divisor),
                                                                       kernel_inv_divisor(divisor), post_scale);
  return finish_launch("sa_corr_fp32 (C = 3 rows)");
}

int launch_truncate_rows(const float* vol, const float* disp, const float* conf, float gain, float omg, float* out,
                         long long rows, int W2, int W3, cudaStream_t st) {
  if (vol) truncate_rows_kernel<true><<<rows_grid(rows), kRowWarps * 32, 0, st>>>(vol, disp, conf, gain, omg, out, rows, W2, W3);
  else truncate_rows_kernel<false><<<rows_grid(rows), kRowWarps * 32, 0, st>>>(vol, disp, conf, gain, omg, out, rows, W2, W3);
  return finish_launch("sa_truncate (rows)");
}

int launch_masked_volume_rows(const float* vol, const float* nl, const float* nr, float divisor, float post_scale,
                              const float* mde_l, const float* mde_r, const float* h_edges, int n_bins, float* out, int B,
                              int H, int W2, int W3, cudaStream_t st) {
  RowBinEdges ed;
  for (int i = 0; i <= n_bins; ++i) ed.e[i] = h_edges[i];
  for (int i = n_bins + 1; i <= SA_MAX_BINS; ++i) ed.e[i] = 0.f;
  const long long rows = (long long)B * H * W2;
  if (vol)
    masked_volume_rows_kernel<false><<<rows_grid(rows), kRowWarps * 32, 0, st>>>(vol, nullptr, nullptr, 1.f, 1.f, 1.f, mde_l, mde_r,
                                                                                ed, n_bins, out, H, W2, W3, rows);
  else
    masked_volume_rows_kernel<true><<<rows_grid(rows), kRowWarps * 32, 0, st>>>(nullptr, nl, nr, kernel_divisor(divisor),
                                                                               kernel_inv_divisor(divisor), post_scale, mde_l,
                                                                               mde_r, ed, n_bins, out, H, W2, W3, rows);
  return finish_launch("sa_masked_volume (rows)");
}

int launch_corrupt_rows(const float* vol, const float* bin_mask, int mode, int shift, const float* noise, float gauss_k,
                        float* out, long long rows, int W2, int W3, cudaStream_t st) {
  corrupt_rows_kernel<<<rows_grid(rows), kRowWarps * 32, 0, st>>>(vol, bin_mask, mode, shift, noise, gauss_k, out, W2, W3, rows);
  return finish_launch("sa_corrupt (rows)");
}

}  // namespace sa
