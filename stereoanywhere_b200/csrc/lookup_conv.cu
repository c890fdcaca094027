// SURVEY 8f-1 - lookup fused with the motion encoder's front end
// (reference: models/stereoanywhere/update.py:74,80-84: `relu(convc1(corr))`, `relu(convc1(corr_mono))`,
// convc1 = Conv2d(36, 64, 1), the SAME weights for both volumes; lookups: stereoanywhere.py:270-271).
//
//   out_v[b, n, h, w] = relu( bias[n] + sum_k weight[n, k] * lookup_v[b, k, h, w] ),  v = stereo, mono
//
// The 2 x 36-channel lookup result never goes to HBM: the blended taps are written straight into the
// canonical UMMA operand layout in shared memory (MN-major, 128B swizzle with 32-byte atoms, pixels
// = M, channels = K padded to 40) and one elected thread issues 2 x 5 tcgen05.mma (M=128, N=64, K=8,
// TF32, fp32 accumulate in TMEM).  The epilogue reads the accumulator with tcgen05.ld (lane = pixel),
// adds the bias, applies ReLU and stores channel-major: for every output channel a warp writes 32
// consecutive pixels = one 128-byte line.
// Traffic per pixel: 2 x 128 B read, 2 x 256 B written, versus 288 B written + 288 B re-read +
// 512 B written by lookup + cuDNN 1x1 convolutions.
#include "sa_common.cuh"
#include "tc_common.cuh"

namespace sa {

constexpr int kLcTile = 128;           // pixels per CTA = MMA M
constexpr int kLcK = 40;               // 36 lookup channels padded to a multiple of the MMA K (8)
constexpr int kLcN = 64;               // convc1 output channels
constexpr int kLcGroupBytes = kLcK * 128;          // one 32-pixel group of the A operand: 40 rows x 128 B
constexpr int kLcABytes = 4 * kLcGroupBytes;       // per volume
constexpr int kLcBBytes = (kLcN / 8) * (kLcK / 4) * 128;  // weights, K-major core matrices (8 x 16 B)

struct LcArgs {
  const float* packed[2];
  float* out[2];
  const float* coords;
  const float* weight;  // [64][36]
  const float* bias;    // [64]
  long long coords_bstride;
  int HW, nblk, B;
  int reverse;  // walk the tiles backwards (alternate launches on the same volume: csrc/packed.cu, next_direction)
  // FACT: packed[1] is the packed pyramid of the right normal map's rows ([(b*3 + c)*H + h][nblk][32], see
  // csrc/packed.cu, factored mono volume), nl the left normals [B,3,H,Wimg], kscale = post_scale / divisor
  const float* nl;
  int H, Wimg;
  float kscale;
};

__device__ __forceinline__ void lc_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
template <int N>
__device__ __forceinline__ void lc_shift_if(float (&v)[N], bool on, int by, int keep) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < keep && i + by < N) v[i] = on ? v[i + by] : v[i];
}
__device__ __forceinline__ float to_tf32(float x) {  // round to nearest (the MMA alone would truncate)
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// byte offset of element (pixel p, channel k) inside one volume's A operand
__device__ __forceinline__ uint32_t a_off(int p, int k) {
  const uint32_t off = (uint32_t)(p >> 5) * kLcGroupBytes + (uint32_t)(k >> 2) * 512 + (uint32_t)(k & 3) * 128 +
                       (uint32_t)(p & 31) * 4;
  return off ^ (((off >> 7) & 3) << 5);
}

template <bool FACT>
__global__ void __launch_bounds__(2 * kLcTile, 4) lookup_conv_kernel(const LcArgs a) {
  constexpr int TILE = kLcTile, THREADS = 2 * kLcTile;
  extern __shared__ uint8_t lc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                                   // 2 x 20 KB A operands; first 32 KB double as line staging
  uint8_t* sB = base + 2 * kLcABytes;                   // 10 KB weights
  float* s_bias = reinterpret_cast<float*>(sB + kLcBBytes);
  float* s_x = s_bias + kLcN;
  unsigned* s_line = reinterpret_cast<unsigned*>(s_x + TILE);   // chunk (16 B) index of the pixel's line, or ~0u
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_line + TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // weights -> K-major 8x(16 B) core matrices, rounded to tf32; channels 36..39 are zero padding
  if ((reinterpret_cast<uintptr_t>(a.weight) & 15) == 0) {
    // one float4 (four consecutive k of one output channel) per step: 128-bit load, 4 x cvt, 128-bit store
    for (int e = tid; e < kLcN * (kLcK / 4); e += THREADS) {
      const int n = e / (kLcK / 4), kq = e % (kLcK / 4);
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kq < 9) {
        wv = __ldg(reinterpret_cast<const float4*>(a.weight + n * 36 + kq * 4));
        wv = make_float4(to_tf32(wv.x), to_tf32(wv.y), to_tf32(wv.z), to_tf32(wv.w));
      }
      *reinterpret_cast<float4*>(sB + ((n >> 3) * (kLcK / 4) + kq) * 128 + (n & 7) * 16) = wv;
    }
  } else {
    for (int e = tid; e < kLcN * kLcK; e += THREADS) {
      const int n = e / kLcK, k = e % kLcK;
      const float wv = (k < 36) ? to_tf32(__ldg(a.weight + n * 36 + k)) : 0.0f;
      *reinterpret_cast<float*>(sB + ((n >> 3) * (kLcK / 4) + (k >> 2)) * 128 + (n & 7) * 16 + (k & 3) * 4) = wv;
    }
  }
  if (tid < kLcN) s_bias[tid] = __ldg(a.bias + tid);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // persistent: the weights, the TMEM allocation and the barrier are set up once per CTA; tiles of 128 pixels
  // are taken round-robin
  const int tiles_x = (a.HW + TILE - 1) / TILE;
  const long long ntiles = (long long)tiles_x * a.B;
  uint32_t phase = 0;
  for (long long it = blockIdx.x; it < ntiles; it += gridDim.x) {
    const long long tile = a.reverse ? ntiles - 1 - it : it;
    const unsigned b = (unsigned)(tile / tiles_x);
    const unsigned hw0 = (unsigned)(tile - (long long)b * tiles_x) * TILE;
  const int npx = min(TILE, a.HW - (int)hw0);

  // All global addressing: uniform 64-bit base + 32-bit index in 16-byte units (as lookup_packed_kernel, round 2).
  // FACT: {k n0, k n1, k n2, chunk index of line (c = 0, h, blk)} per pixel, in the 8 KB of the A operands that the
  // line staging leaves free
  float4* s_n = reinterpret_cast<float4*>(sA + 2 * TILE * 128);
  if (tid < TILE) {
    float x = 0.f;
    int blk = -1;
    unsigned line = ~0u;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (tid < npx) {
      x = __ldg(a.coords + (size_t)b * a.coords_bstride + hw0 + tid);
      if (FACT) {
        const unsigned plane = (unsigned)a.H * a.Wimg;
        const float* nlp = a.nl + (size_t)(b * 3 * plane + hw0 + tid);
        n0 = __ldg(nlp); n1 = __ldg(nlp + plane); n2 = __ldg(nlp + 2 * plane);
      }
      const float fl = fminf(fmaxf(floorf(x), -1.0e6f), 1.0e6f);
      const int q = ((int)fl >> 3) + 5;  // blocks start at q = -5 (csrc/packed.cu)
      if (q >= 0 && q < a.nblk) {
        blk = q;
        line = ((b * (unsigned)a.HW + hw0 + tid) * (unsigned)a.nblk + (unsigned)q) * 8u;
      }
    }
    s_x[tid] = x;
    s_line[tid] = line;
    if (FACT) {  // a pixel without a line reads line 0 with zero coefficients
      const float k = blk >= 0 ? a.kscale : 0.f;
      const unsigned h0 = hw0 / (unsigned)a.Wimg;      // uniform; a pixel of the tile is a few rows further down
      const unsigned t = hw0 - h0 * (unsigned)a.Wimg + tid;
      unsigned hh = h0;
      for (unsigned w = a.Wimg; w <= t; w += a.Wimg) ++hh;
      const unsigned off = blk >= 0 ? (((b * 3u) * a.H + hh) * (unsigned)a.nblk + (unsigned)blk) * 8u : 0u;
      s_n[tid] = make_float4(n0 * k, n1 * k, n2 * k, __uint_as_float(off));
    }
  }
  __syncthreads();

  // ---- stage one packed line per (pixel, volume), chunk c of pixel p at chunk c ^ (p & 7); pixels without a line
  // are zero-filled by cp.async's src-size 0 form
  float* stage = reinterpret_cast<float*>(sA);
  {
    constexpr int UPS = THREADS / 8, SPV = 4;  // units per step, steps per volume
    const unsigned ch = tid & 7, u0 = tid >> 3;
    const float4* p0 = reinterpret_cast<const float4*>(a.packed[0]);
    const float4* p1 = reinterpret_cast<const float4*>(a.packed[1]);
#pragma unroll
    for (int m = 0; m < SPV; ++m) {
      const unsigned pm = u0 + m * UPS;
      const unsigned line = s_line[pm];
      const unsigned ok = line != ~0u ? 16u : 0u;
      const unsigned idx = (line != ~0u ? line : 0u) + ch;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        if (FACT && v == 1) continue;
        float* dst = stage + (v * TILE + pm) * 32 + ((ch ^ (pm & 7)) << 2);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                     "l"((v ? p1 : p0) + idx), "r"(ok)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (FACT) {  // the mono line = three right-normal lines combined with the pixel's scaled left normal
      const unsigned cplane = (unsigned)a.H * a.nblk * 8u;
      const float4* rp = reinterpret_cast<const float4*>(a.packed[1]);
      float4 r[SPV][3], nn[SPV];
#pragma unroll
      for (int m = 0; m < SPV; ++m) {
        nn[m] = s_n[u0 + m * UPS];
        const unsigned off = __float_as_uint(nn[m].w) + ch;
#pragma unroll
        for (int c = 0; c < 3; ++c) r[m][c] = __ldg(rp + (off + c * cplane));
      }
#pragma unroll
      for (int m = 0; m < SPV; ++m) {
        const int pm = u0 + m * UPS;
        float4 o;
        o.x = fmaf(nn[m].z, r[m][2].x, fmaf(nn[m].y, r[m][1].x, nn[m].x * r[m][0].x));
        o.y = fmaf(nn[m].z, r[m][2].y, fmaf(nn[m].y, r[m][1].y, nn[m].x * r[m][0].y));
        o.z = fmaf(nn[m].z, r[m][2].z, fmaf(nn[m].y, r[m][1].z, nn[m].x * r[m][0].z));
        o.w = fmaf(nn[m].z, r[m][2].w, fmaf(nn[m].y, r[m][1].w, nn[m].x * r[m][0].w));
        *reinterpret_cast<float4*>(stage + (TILE + pm) * 32 + ((ch ^ (pm & 7)) << 2)) = o;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int p = tid % TILE, v = tid / TILE;
  float l0[20], e[12];
  {
    const float* src = stage + tid * 32;
#pragma unroll
    for (int ch = 0; ch < 5; ++ch) {
      const float4 t = *reinterpret_cast<const float4*>(src + ((ch ^ (p & 7)) << 2));
      l0[4 * ch] = t.x; l0[4 * ch + 1] = t.y; l0[4 * ch + 2] = t.z; l0[4 * ch + 3] = t.w;
    }
#pragma unroll
    for (int ch = 5; ch < 8; ++ch) {
      const float4 t = *reinterpret_cast<const float4*>(src + ((ch ^ (p & 7)) << 2));
      e[4 * (ch - 5)] = t.x; e[4 * (ch - 5) + 1] = t.y; e[4 * (ch - 5) + 2] = t.z; e[4 * (ch - 5) + 3] = t.w;
    }
  }
  __syncthreads();  // staging is dead: the same bytes now become the A operands

  float l1[13], l2[11], l3[10];
  l1[0] = l0[17]; l1[1] = l0[18]; l1[10] = l0[19]; l1[11] = e[0]; l1[12] = e[1];
#pragma unroll
  for (int t = 0; t < 8; ++t) l1[2 + t] = (l0[2 * t] + l0[2 * t + 1]) * 0.5f;
  l2[0] = e[2]; l2[1] = e[3]; l2[8] = e[4]; l2[9] = e[5]; l2[10] = e[6];
#pragma unroll
  for (int t = 0; t < 6; ++t) l2[2 + t] = (l1[2 * t] + l1[2 * t + 1]) * 0.5f;
  l3[0] = e[7]; l3[1] = e[8]; l3[7] = e[9]; l3[8] = e[10]; l3[9] = e[11];
#pragma unroll
  for (int t = 0; t < 5; ++t) l3[2 + t] = (l2[2 * t] + l2[2 * t + 1]) * 0.5f;

  const float x = s_x[p];
  const int x0 = (int)fminf(fmaxf(floorf(x), -1.0e6f), 1.0e6f);
  float w0[17];
#pragma unroll
  for (int i = 0; i < 17; ++i) w0[i] = l0[i];
  lc_shift_if(w0, (x0 & 1) != 0, 1, 16);
  lc_shift_if(w0, (x0 & 2) != 0, 2, 14);
  lc_shift_if(w0, (x0 & 4) != 0, 4, 10);
  const int x1 = x0 >> 1;
  lc_shift_if(l1, (x1 & 1) != 0, 1, 12);
  lc_shift_if(l1, (x1 & 2) != 0, 2, 10);
  lc_shift_if(l2, ((x0 >> 2) & 1) != 0, 1, 10);

  uint8_t* Av = sA + v * kLcABytes;
  {
    const float f = x - floorf(x);
#pragma unroll
    for (int k = 0; k < 9; ++k) *reinterpret_cast<float*>(Av + a_off(p, k)) = to_tf32(blend(w0[k], w0[k + 1], f));
  }
  {
    const float xs = x * 0.5f, f = xs - floorf(xs);
#pragma unroll
    for (int k = 0; k < 9; ++k) *reinterpret_cast<float*>(Av + a_off(p, 9 + k)) = to_tf32(blend(l1[k], l1[k + 1], f));
  }
  {
    const float xs = x * 0.25f, f = xs - floorf(xs);
#pragma unroll
    for (int k = 0; k < 9; ++k) *reinterpret_cast<float*>(Av + a_off(p, 18 + k)) = to_tf32(blend(l2[k], l2[k + 1], f));
  }
  {
    const float xs = x * 0.125f, f = xs - floorf(xs);
#pragma unroll
    for (int k = 0; k < 9; ++k) *reinterpret_cast<float*>(Av + a_off(p, 27 + k)) = to_tf32(blend(l3[k], l3[k + 1], f));
  }
#pragma unroll
  for (int k = 36; k < kLcK; ++k) *reinterpret_cast<float*>(Av + a_off(p, k)) = 0.0f;

  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      // D = f32, A = tf32 MN-major (transpose bit 15), B = tf32 K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(kLcN >> 3) << 17) |
                             ((uint32_t)(TILE >> 4) << 24);
      const uint32_t sb_ = smem_u32(sB);
#pragma unroll
      for (int vv = 0; vv < 2; ++vv) {
        const uint32_t sa_ = smem_u32(sA + vv * kLcABytes);
#pragma unroll
        for (int ks = 0; ks < kLcK / 8; ++ks) {
          const uint64_t ad = make_desc(sa_ + ks * 1024, kLcGroupBytes, 512, 1);
          const uint64_t bd = make_desc(sb_ + ks * 256, 128, (kLcK / 4) * 128, 0);
          umma_tf32(tmem + vv * kLcN, ad, bd, idesc, (uint32_t)(ks != 0));
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, phase);
  phase ^= 1u;
  tc_fence_after();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. (its pixels), columns of its volume (w/4)
  const int vq = warp >> 2, quarter = warp & 3;
  const int px = quarter * 32 + lane;
  float* outp = (vq ? a.out[1] : a.out[0]) + ((size_t)b * kLcN * a.HW + hw0 + px);
  const bool live = px < npx;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(vq * kLcN + half * 32)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (live) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int n = half * 32 + c;
        st_stream_f32(outp + (unsigned)(n * a.HW), fmaxf(__uint_as_float(r[c]) + s_bias[n], 0.0f));
      }
    }
  }
  tc_fence_before();
  __syncthreads();  // the next tile rewrites the shared-memory operands and the accumulators
  tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

}  // namespace sa

namespace sa {
int lookup_next_direction(const void* key);  // csrc/packed.cu

template <bool FACT>
static int launch_lookup_conv(LcArgs a, cudaStream_t st, const char* what) {
  a.reverse = lookup_next_direction(a.packed[0]);
  const size_t smem = 1024 + 2 * kLcABytes + kLcBBytes + (kLcN + 2 * kLcTile) * sizeof(float) + 32;
  cudaError_t e = cudaFuncSetAttribute(lookup_conv_kernel<FACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) SA_FAIL((int)e, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
  const long long ntiles = (long long)((a.HW + kLcTile - 1) / kLcTile) * a.B;
  const long long cap = (long long)num_sms() * 4;  // 4 CTAs per SM fit (shared memory, 128 TMEM columns each)
  const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
  lookup_conv_kernel<FACT><<<grid, 2 * kLcTile, smem, st>>>(a);
  return finish_launch(what);
}
}  // namespace sa

extern "C" int sa_lookup_packed_conv(const float* packed_a, const float* packed_b, int W3, const float* coords,
                                     int64_t coords_bstride, const float* weight, const float* bias, float* out_a,
                                     float* out_b, int B, int H, int W, void* stream) {
  using namespace sa;
  SA_REQUIRE(packed_a && packed_b && coords && weight && bias && out_a && out_b, SA_E_INVALID,
             "sa_lookup_packed_conv: null pointer");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 31), SA_E_INVALID,
             "sa_lookup_packed_conv: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_lookup_packed_conv: W3 must be a multiple of 8");
  SA_REQUIRE((long long)B * H * W * (W3 / 8 + 9) < (1ll << 29) && (long long)B * 64 * H * W < (1ll << 32), SA_E_UNSUPPORTED,
             "sa_lookup_packed_conv: arrays of 64 GB or more (32-bit chunk indices)");
  SA_REQUIRE(aligned16(packed_a) && aligned16(packed_b), SA_E_ALIGN, "sa_lookup_packed_conv: packed arrays must be 16-byte aligned");
  LcArgs a = {};
  a.packed[0] = packed_a; a.packed[1] = packed_b;
  a.out[0] = out_a; a.out[1] = out_b;
  a.coords = coords; a.coords_bstride = coords_bstride;
  a.weight = weight; a.bias = bias;
  a.HW = H * W;
  a.nblk = W3 / 8 + 9;
  a.B = B;
  return launch_lookup_conv<false>(a, (cudaStream_t)stream, "sa_lookup_packed_conv");
}

extern "C" int sa_lookup_factored_conv(const float* packed_a, const float* packed_normals_r, const float* normals_l,
                                       float divisor, float post_scale, int W3, const float* coords,
                                       int64_t coords_bstride, const float* weight, const float* bias, float* out_a,
                                       float* out_mono, int B, int H, int W, void* stream) {
  using namespace sa;
  SA_REQUIRE(packed_a && packed_normals_r && normals_l && coords && weight && bias && out_a && out_mono, SA_E_INVALID,
             "sa_lookup_factored_conv: null pointer");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 31) && divisor != 0.f, SA_E_INVALID,
             "sa_lookup_factored_conv: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_lookup_factored_conv: W3 must be a multiple of 8");
  SA_REQUIRE((long long)B * H * W * (W3 / 8 + 9) < (1ll << 29) && (long long)B * 64 * H * W < (1ll << 32), SA_E_UNSUPPORTED,
             "sa_lookup_factored_conv: arrays of 64 GB or more (32-bit chunk indices)");
  SA_REQUIRE(aligned16(packed_a) && aligned16(packed_normals_r), SA_E_ALIGN,
             "sa_lookup_factored_conv: packed arrays must be 16-byte aligned");
  LcArgs a = {};
  a.packed[0] = packed_a; a.packed[1] = packed_normals_r;
  a.out[0] = out_a; a.out[1] = out_mono;
  a.coords = coords; a.coords_bstride = coords_bstride;
  a.weight = weight; a.bias = bias;
  a.HW = H * W;
  a.nblk = W3 / 8 + 9;
  a.B = B;
  a.nl = normals_l; a.H = H; a.Wimg = W;
  a.kscale = post_scale * kernel_inv_divisor(divisor);
  return launch_lookup_conv<true>(a, (cudaStream_t)stream, "sa_lookup_factored_conv");
}
