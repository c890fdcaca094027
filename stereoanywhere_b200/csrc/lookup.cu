// A4 - per-GRU-iteration multi-level lookup (reference: models/stereoanywhere/corr.py:93-115 +
// utils/utils.py:19-35).  HBM-bound gather: per left pixel and level one (2r+2)-float window
// at an arbitrary, unaligned column of that pixel's own volume row, one interpolation weight,
// 2r+1 outputs written channel-major (NCHW) for the motion encoder.
//
// Fast kernel (lookup_vec_kernel): a CTA owns TILE consecutive pixels of one batch image.
//   phase 1  four adjacent lanes own one (pixel, level) window: each issues ONE aligned 128-bit
//            load (4 x 16 B always cover a (2r+2 <= 12)-float window at any alignment), the
//            neighbour element comes from a shuffle, each lane blends the taps it holds and
//            drops them into a [channel][pixel] tile in shared memory;
//   phase 2  the tile is streamed out with 128-bit stores, one contiguous run per channel.
// Two-volume variant (NV = 2): the stereo and the mono block share one coords read and one launch.
#include "sa_common.cuh"

namespace sa {

constexpr int kVecLevels = 4;

struct LookupArgs {
  const float* lvl[2][SA_MAX_LEVELS];
  long long pitch[2][SA_MAX_LEVELS];
  int width[SA_MAX_LEVELS];
  const float* coords;
  long long coords_bstride;
  float* out[2];
  int HW;      // pixels per batch image
  int W;       // image width (generic kernel)
  int num_levels, radius, pad0, pad1;
  float xoff;  // pad0 as float
};

template <int R, int NL, int NV, int TILE, int THREADS>
__global__ void __launch_bounds__(THREADS) lookup_vec_kernel(const LookupArgs a) {
  constexpr int NT = 2 * R + 1;    // taps per level
  constexpr int NC = NL * NT;      // output channels per volume
  constexpr int SP = TILE + 4;     // smem pitch: keeps rows 16 B aligned, spreads channels over banks
  constexpr int GROUPS = THREADS / 4;
  constexpr int ITEMS = NL * TILE;             // (level, pixel) windows per volume
  static_assert(ITEMS % GROUPS == 0, "uniform trip count (shuffles inside the loop)");
  constexpr int PER = ITEMS / GROUPS;          // windows per 4-lane group and volume
  static_assert(2 * R + 4 <= 15, "window + misalignment must fit four 16-byte chunks");

  extern __shared__ __align__(16) float smem[];
  float* s_out = smem;                  // [NV*NC][SP]
  float* s_x = smem + NV * NC * SP;     // [TILE]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * TILE;
  const int npx = min(TILE, a.HW - hw0);

  const float* cx = a.coords + (long long)b * a.coords_bstride + hw0;
  for (int t = tid; t < TILE; t += THREADS) s_x[t] = (t < npx) ? (__ldg(cx + t) + a.xoff) : 0.0f;
  __syncthreads();

  const int grp = tid >> 2;
  const int j = tid & 3;
  const long long row0 = (long long)b * a.HW + hw0;

#pragma unroll
  for (int v = 0; v < NV; ++v) {
    // ---- issue every load of this volume before touching any result: PER independent 128-bit
    //      requests in flight per lane (the gather is latency-bound otherwise)
    float4 q[PER];
    float fr[PER];
    int kb[PER];
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int item = grp + n * GROUPS;
      const int t = item % TILE;
      const int i = item / TILE;
      const float xs = s_x[t] * __int_as_float((127 - i) << 23);  // x / 2^i, exact
      float fl = floorf(xs);
      fr[n] = xs - fl;
      fl = fminf(fmaxf(fl, -1.0e6f), 1.0e6f);  // far-away / inf coords: every tap lands outside
      const int start = (int)fl - R;
      const int o = start & 3;
      kb[n] = 4 * j - o;                        // tap index of this lane's first element
      const int e0 = ((start >> 2) + j) << 2;   // first column of this lane's 16-byte chunk
      const int wi = a.width[i];
      // chunk needed iff it holds one of the 2R+2 window elements and lies inside the row
      const bool ok = (t < npx) && (e0 >= 0) && (e0 < wi) && (4 * j <= o + 2 * R + 1);
      q[n] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) q[n] = ld_stream_v4(a.lvl[v][i] + (row0 + t) * a.pitch[v][i] + e0);
      if (e0 + 1 >= wi) q[n].y = 0.f;  // row padding / right border
      if (e0 + 2 >= wi) q[n].z = 0.f;
      if (e0 + 3 >= wi) q[n].w = 0.f;
    }
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int item = grp + n * GROUPS;
      const int t = item % TILE;
      const int i = item / TILE;
      const float nx = __shfl_down_sync(0xffffffffu, q[n].x, 1);
      const float f = fr[n];
      float* so = s_out + ((v * NL + i) * NT) * SP + t;
      const int k0 = kb[n];
      if ((unsigned)(k0 + 0) < (unsigned)NT) so[(k0 + 0) * SP] = blend(q[n].x, q[n].y, f);
      if ((unsigned)(k0 + 1) < (unsigned)NT) so[(k0 + 1) * SP] = blend(q[n].y, q[n].z, f);
      if ((unsigned)(k0 + 2) < (unsigned)NT) so[(k0 + 2) * SP] = blend(q[n].z, q[n].w, f);
      if ((unsigned)(k0 + 3) < (unsigned)NT) so[(k0 + 3) * SP] = blend(q[n].w, nx, f);
    }
  }
  __syncthreads();

  // phase 2: [channel][pixel] tile -> NCHW, one contiguous run of npx floats per channel
  if ((a.HW & 3) == 0) {
    constexpr int T4 = TILE / 4;
    for (int idx = tid; idx < NV * NC * T4; idx += THREADS) {
      const int c = idx / T4;
      const int t = (idx % T4) * 4;
      if (t < npx) {
        const float4 val = *reinterpret_cast<const float4*>(s_out + c * SP + t);
        const int v = c / NC;
        const int cc = c - v * NC;
        st_stream_v4(a.out[v] + ((long long)b * NC + cc) * a.HW + hw0 + t, val);
      }
    }
  } else {
    for (int idx = tid; idx < NV * NC * TILE; idx += THREADS) {
      const int c = idx / TILE;
      const int t = idx % TILE;
      if (t < npx) {
        const int v = c / NC;
        const int cc = c - v * NC;
        a.out[v][((long long)b * NC + cc) * a.HW + hw0 + t] = s_out[c * SP + t];
      }
    }
  }
}

// Generic kernel: any radius / level count / pitch / pad.  One thread per left pixel.
template <int NV>
__global__ void lookup_generic_kernel(const LookupArgs a) {
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (hw >= a.HW) return;
  const int w = hw % a.W;
  const int h = hw / a.W;
  const int wout = a.W - a.pad0 - a.pad1;
  if (w < a.pad0 || w >= a.W - a.pad1) return;
  const int r = a.radius;
  const int nt = 2 * r + 1;
  const int nc = a.num_levels * nt;
  const int H = a.HW / a.W;
  const float x = a.coords[(long long)b * a.coords_bstride + hw] + a.xoff;
  const long long row = (long long)b * a.HW + hw;
  for (int v = 0; v < NV; ++v) {
    float* outp = a.out[v] + (((long long)b * nc) * H + h) * wout + (w - a.pad0);
    for (int i = 0; i < a.num_levels; ++i) {
      const float xs = x * __int_as_float((127 - i) << 23);
      float fl = floorf(xs);
      const float f = xs - fl;
      fl = fminf(fmaxf(fl, -1.0e6f), 1.0e6f);
      const int start = (int)fl - r;
      const int wi = a.width[i];
      const float* rowp = a.lvl[v][i] + row * a.pitch[v][i];
      float prev = (start >= 0 && start < wi) ? rowp[start] : 0.f;
      for (int k = 0; k < nt; ++k) {
        const int c1 = start + k + 1;
        const float nxt = (c1 >= 0 && c1 < wi) ? rowp[c1] : 0.f;
        outp[(long long)(i * nt + k) * H * wout] = blend(prev, nxt, f);
        prev = nxt;
      }
    }
  }
}

template <int R, int NL, int NV>
static int launch_vec(const LookupArgs& a, int B, cudaStream_t st) {
  constexpr int TILE = 128, THREADS = 256;
  constexpr int NC = NL * (2 * R + 1);
  const size_t smem = (size_t)(NV * NC * (TILE + 4) + TILE) * sizeof(float);
  auto kern = lookup_vec_kernel<R, NL, NV, TILE, THREADS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) SA_FAIL((int)e, "lookup: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  dim3 grid((a.HW + TILE - 1) / TILE, B);
  kern<<<grid, THREADS, smem, st>>>(a);
  return finish_launch("sa_lookup (vec)");
}

template <int NV>
static int dispatch_vec(const LookupArgs& a, int B, cudaStream_t st, bool* handled) {
  *handled = true;
#define SA_CASE(R_, NL_) \
  if (a.radius == R_ && a.num_levels == NL_) return launch_vec<R_, NL_, NV>(a, B, st);
  SA_CASE(4, 4) SA_CASE(4, 3) SA_CASE(4, 2) SA_CASE(4, 1)
  SA_CASE(3, 4) SA_CASE(3, 3) SA_CASE(3, 2)
  SA_CASE(2, 4) SA_CASE(2, 3) SA_CASE(2, 2)
  SA_CASE(5, 4) SA_CASE(1, 4)
#undef SA_CASE
  *handled = false;
  return 0;
}

template <int NV>
static int lookup_common(LookupArgs& a, int B, int H, int W, cudaStream_t st) {
  SA_REQUIRE(B > 0 && H > 0 && W > 0, SA_E_INVALID, "lookup: B, H, W must be positive");
  SA_REQUIRE(a.num_levels >= 1 && a.num_levels <= SA_MAX_LEVELS, SA_E_INVALID, "lookup: 1 <= num_levels <= %d",
             SA_MAX_LEVELS);
  SA_REQUIRE(a.radius >= 0 && a.radius <= 64, SA_E_INVALID, "lookup: 0 <= radius <= 64");
  SA_REQUIRE(a.pad0 >= 0 && a.pad1 >= 0 && a.pad0 + a.pad1 < W, SA_E_INVALID, "lookup: bad pad");
  SA_REQUIRE((long long)H * W < (1ll << 31) && B <= 65535, SA_E_UNSUPPORTED, "lookup: H*W < 2^31, B <= 65535");
  SA_REQUIRE(a.coords && a.out[0], SA_E_INVALID, "lookup: null coords / out");
  bool vec = a.pad0 == 0 && a.pad1 == 0 && a.num_levels <= kVecLevels;
  for (int v = 0; v < NV; ++v)
    for (int i = 0; i < a.num_levels; ++i) {
      SA_REQUIRE(a.lvl[v][i] != nullptr, SA_E_INVALID, "lookup: null level pointer");
      SA_REQUIRE(a.width[i] >= 1 && a.pitch[v][i] >= a.width[i], SA_E_INVALID, "lookup: bad width / pitch");
      vec = vec && aligned16(a.lvl[v][i]) && (a.pitch[v][i] % 4 == 0);
    }
  for (int v = 0; v < NV; ++v) vec = vec && aligned16(a.out[v]);
  (void)num_sms();  // also applies the process-wide device limits once
  a.HW = H * W;
  a.W = W;
  a.xoff = (float)a.pad0;
  if (vec) {
    bool handled = false;
    int rc = dispatch_vec<NV>(a, B, st, &handled);
    if (handled) return rc;
  }
  dim3 grid((a.HW + 127) / 128, B);
  lookup_generic_kernel<NV><<<grid, 128, 0, st>>>(a);
  return finish_launch("sa_lookup (generic)");
}

}  // namespace sa

extern "C" int sa_lookup(const float* const* h_levels, const int* h_widths, const int64_t* h_pitches,
                         int num_levels, int radius, const float* coords, int64_t coords_bstride, float* out,
                         int B, int H, int W, int pad0, int pad1, void* stream) {
  SA_REQUIRE(h_levels && h_widths && h_pitches, SA_E_INVALID, "sa_lookup: null host arrays");
  SA_REQUIRE(num_levels >= 1 && num_levels <= SA_MAX_LEVELS, SA_E_INVALID, "sa_lookup: 1 <= num_levels <= %d",
             SA_MAX_LEVELS);
  sa::LookupArgs a = {};
  for (int i = 0; i < num_levels; ++i) {
    a.lvl[0][i] = h_levels[i];
    a.pitch[0][i] = h_pitches[i];
    a.width[i] = h_widths[i];
  }
  a.num_levels = num_levels;
  a.radius = radius;
  a.coords = coords;
  a.coords_bstride = coords_bstride;
  a.out[0] = out;
  a.pad0 = pad0;
  a.pad1 = pad1;
  return sa::lookup_common<1>(a, B, H, W, (cudaStream_t)stream);
}

extern "C" int sa_lookup2(const float* const* h_levels_a, const float* const* h_levels_b, const int* h_widths,
                          const int64_t* h_pitches_a, const int64_t* h_pitches_b, int num_levels, int radius,
                          const float* coords, int64_t coords_bstride, float* out_a, float* out_b, int B, int H,
                          int W, void* stream) {
  SA_REQUIRE(h_levels_a && h_levels_b && h_widths && h_pitches_a && h_pitches_b, SA_E_INVALID,
             "sa_lookup2: null host arrays");
  SA_REQUIRE(num_levels >= 1 && num_levels <= SA_MAX_LEVELS, SA_E_INVALID, "sa_lookup2: 1 <= num_levels <= %d",
             SA_MAX_LEVELS);
  SA_REQUIRE(out_b != nullptr, SA_E_INVALID, "sa_lookup2: null out_b");
  sa::LookupArgs a = {};
  for (int i = 0; i < num_levels; ++i) {
    a.lvl[0][i] = h_levels_a[i];
    a.lvl[1][i] = h_levels_b[i];
    a.pitch[0][i] = h_pitches_a[i];
    a.pitch[1][i] = h_pitches_b[i];
    a.width[i] = h_widths[i];
  }
  a.num_levels = num_levels;
  a.radius = radius;
  a.coords = coords;
  a.coords_bstride = coords_bstride;
  a.out[0] = out_a;
  a.out[1] = out_b;
  return sa::lookup_common<2>(a, B, H, W, (cudaStream_t)stream);
}
