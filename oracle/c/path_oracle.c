/* Plain-C restatement of the cost-volume path's closed forms.
 *
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Built by `__graft_entry__.build()` into
 * oracle/_build/libpath_oracle.so and loaded only by tests/ (ctypes).  It restates, in scalar C
 * with float arithmetic in the reference's operation order:
 *   A1/A2  models/stereoanywhere/corr.py:117-132      oracle_corr
 *   A3     models/stereoanywhere/corr.py:76-91        oracle_pyramid_level
 *   A4     models/stereoanywhere/corr.py:93-115 +
 *          utils/utils.py:19-35 (grid_sample, zeros)  oracle_lookup
 *   A5     utils/utils.py:216-238                     oracle_truncation
 * Pinned against tests/golden/path_small.npz (generated from the unmodified reference) by
 * tests/test_oracle_golden.py::test_c_restatement.
 */
#include <math.h>
#include <stdint.h>

/* vol[b,h,w2,w3] = (sum_c L[b,c,h,w2] R[b,c,h,w3]) / divisor * post */
void oracle_corr(const float* fl, const float* fr, float* vol, int B, int C, int H, int W2, int W3, float divisor,
                 float post) {
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int i = 0; i < W2; ++i)
        for (int j = 0; j < W3; ++j) {
          double acc = 0.0; /* the reference accumulates in fp32 (MKL order unknown); fp64 bounds it */
          for (int c = 0; c < C; ++c)
            acc += (double)fl[(((int64_t)b * C + c) * H + h) * W2 + i] * (double)fr[(((int64_t)b * C + c) * H + h) * W3 + j];
          vol[(((int64_t)b * H + h) * W2 + i) * W3 + j] = (float)acc / divisor * post;
        }
}

/* dst[r, j] = 0.5 * (src[r, 2j] + src[r, 2j+1]),  j < floor(w / 2) */
void oracle_pyramid_level(const float* src, float* dst, int64_t rows, int w) {
  const int wo = w / 2;
  for (int64_t r = 0; r < rows; ++r)
    for (int j = 0; j < wo; ++j) dst[r * wo + j] = (src[r * w + 2 * j] + src[r * w + 2 * j + 1]) * 0.5f;
}

static float tap(const float* row, int w, int j) { return (j >= 0 && j < w) ? row[j] : 0.0f; }

/* out[b, i*(2r+1)+k+r, h, w] for one level i (level rows: [B*H*W, wi]) */
void oracle_lookup(const float* level, int wi, int i, const float* coords_x, float* out, int B, int H, int W, int radius,
                   int num_levels) {
  const int nt = 2 * radius + 1;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const int64_t p = ((int64_t)b * H + h) * W + w;
        const float xs = coords_x[p] / (float)(1 << i);
        for (int k = -radius; k <= radius; ++k) {
          const float x = xs + (float)k;
          const float x0 = floorf(x);
          const float f = x - x0;
          const float* row = level + p * wi;
          const float v = (1.0f - f) * tap(row, wi, (int)x0) + f * tap(row, wi, (int)x0 + 1);
          out[(((int64_t)b * num_levels * nt + i * nt + (k + radius)) * H + h) * W + w] = v;
        }
      }
}

/* mask[b,h,w2,w3] = (1-c) + c * (sigmoid((w2 - d) - w3) * (1-g) + g) */
void oracle_truncation(const float* disp, const float* conf, float* mask, int64_t rows, int W2, int W3, double gain) {
  const float g = (float)gain, omg = (float)(1.0 - gain);
  for (int64_t r = 0; r < rows; ++r) {
    const float centre = (float)(r % W2) - disp[r];
    const float c = conf[r];
    for (int j = 0; j < W3; ++j) {
      const float s = 1.0f / (1.0f + expf(-(centre - (float)j)));
      mask[r * W3 + j] = (1.0f - c) + c * (s * omg + g);
    }
  }
}
