// the four soft-argmax / entropy reductions over an aggregated volume [B,1,H,W2,W3], two per
// launch with ONE read of the volume from HBM (the reference runs four separate softmax passes,
// models/stereoanywhere/utils/utils.py:112-170, called at stereoanywhere.py:174-177):
//   DISP  left [b,h,w2] = w2 - sum_w3 softmax_w3(V)[w3] * w3          (estimate_left_disparity,  :112-131)
//         right[b,h,w3] = sum_w2 softmax_w2(V)[w2] * w2 - w3          (estimate_right_disparity, :133-152)
//   CONF  left [b,h,w2] = 1 + sum_w3 p log2(p + 1e-6) / log2(W3)      (estimate_left_confidence,  :154-161)
//         right[b,h,w3] = 1 + sum_w2 p log2(p + 1e-6) / log2(W2)      (estimate_right_confidence, :163-170)
//
// HBM: reads B*H*W2*W3*4 once, writes B*H*(W2+W3)*4 (design notes at the kernel).
#include "sa_common.cuh"

namespace sa {

struct VReduceArgs {
  const float* vol;
  float* out_l;  // [BH, W2]
  float* out_r;  // [BH, W3]
  int W2, W3;
  float inv_log2_w2, inv_log2_w3;
};

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {  // 2^x, one MUFU; 2^-inf = 0
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float warp_max(float v) {  // one CREDUX.MAX.F32 on sm_100a
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// Sum N (power of two, <= 4... 8) per-lane values over the warp with N-1 + log2(32/N) shuffles instead of 5 N:
// every butterfly step halves the number of live values (the lane keeps one half and ships the other).  On
// return v[0] of lanes [i * 32/N, (i+1) * 32/N) holds the warp total of the original v[i].
template <int N>
__device__ __forceinline__ void warp_sum_multi(float (&v)[N], int lane) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = hi ? v[i] : v[i + n / 2];
      const float keep = hi ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

// One CTA (8 or 16 warps) per image row (b,h) = one [W2 x W3] slab.  A warp takes groups of G consecutive volume
// rows straight from global memory into registers (lane l holds columns l + 32k, k < KMAX: coalesced 128-byte
// loads, G*KMAX of them in flight per warp) and uses the same registers for BOTH directions:
//   rows    : the softmax's own two-pass form (row max by shuffle, exp, sums by shuffle);
//   columns : lane-private online softmax state (m, s, t) for its KMAX columns over the rows this warp sees,
//             rescaled once per group (group max first, then G independent exps), merged over the 8 warps
//             through shared memory at the end.
// Two MUFU.EX2 per element in total.  CONF (the +1e-6 sits inside the log, so p itself is needed) sweeps the
// slab a second time - out of L2 - with the final (m, s) of every column.
template <bool CONF, int KMAX, int G, int kVrWarps>
__global__ void __launch_bounds__(kVrWarps * 32) volume_reduce_kernel(const VReduceArgs a) {
  extern __shared__ __align__(16) float vr_smem[];  // [kVrWarps][3][W3] column states / partial sums
  const int W2 = a.W2, W3 = a.W3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long slab = blockIdx.x;
  const float* src = a.vol + slab * (long long)W2 * W3;
  const float lanef = (float)lane;

  float cm[KMAX], cs[KMAX], ct[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { cm[k] = -INFINITY; cs[k] = 0.f; ct[k] = 0.f; }

  for (int r0 = warp * G; r0 < W2; r0 += kVrWarps * G) {
    float x[G][KMAX];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float* row = src + (long long)min(r0 + g, W2 - 1) * W3 + lane;  // clamped: the tail is masked below
#pragma unroll
      for (int k = 0; k < KMAX; ++k) x[g][k] = (lane + 32 * k < W3) ? ld_stream_f32(row + 32 * k) : -INFINITY;
    }
    if (r0 + G > W2) {  // last, partial group of this warp (warp-uniform, at most once per slab)
#pragma unroll
      for (int g = 1; g < G; ++g) {
        if (r0 + g >= W2) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) x[g][k] = -INFINITY;
        }
      }
    }
    // ---- columns (before the rows overwrite x with their exponentials)
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (lane + 32 * k < W3) {
        float tm = cm[k];
#pragma unroll
        for (int g = 0; g < G; ++g) tm = fmaxf(tm, x[g][k]);
        const float tml = tm * kLog2e;
        const float resc = ex2_approx(__fmaf_rn(cm[k], kLog2e, -tml));  // 0 for the first group (cm = -inf)
        float sn = 0.f, tn = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float e = ex2_approx(__fmaf_rn(x[g][k], kLog2e, -tml));  // rows beyond W2 hold -inf -> 0
          sn += e;
          if (!CONF) tn = __fmaf_rn(e, (float)g, tn);
        }
        cs[k] = __fmaf_rn(cs[k], resc, sn);
        // weights are centred on the column (w2 - w3): the sum IS the disparity, no large-number cancellation
        if (!CONF) ct[k] = __fmaf_rn(ct[k], resc, __fmaf_rn((float)r0 - (lanef + (float)(32 * k)), sn, tn));
        cm[k] = tm;
      }
    }
    // ---- rows
    float m[G], sum[G], acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float mm = x[g][0];
#pragma unroll
      for (int k = 1; k < KMAX; ++k) mm = fmaxf(mm, x[g][k]);
      m[g] = warp_max(mm);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float ml = m[g] * kLog2e;
      float s1 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const float e = ex2_approx(__fmaf_rn(x[g][k], kLog2e, -ml));  // columns beyond W3 hold -inf -> 0
        x[g][k] = e;
        s1 += e;
        if (!CONF) a1 = __fmaf_rn(e, (float)(-32 * k), a1);
      }
      sum[g] = s1;
      // sum_k e_k * (w2 - w3), w3 = lane + 32 k: the per-lane cancellation costs ~W * 2^-24 px, far below 1e-3
      acc[g] = __fmaf_rn((float)(r0 + g) - lanef, s1, a1);
    }
    constexpr int kOwn = 32 / G;  // lanes [g * kOwn, (g+1) * kOwn) end up with the totals of row g
    warp_sum_multi<G>(sum, lane);
    if (CONF) {
      float tot[G];
#pragma unroll
      for (int g = 0; g < G; ++g) tot[g] = __shfl_sync(0xffffffffu, sum[0], g * kOwn);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float inv = 1.0f / tot[g];
        float a1 = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const float p = x[g][k] * inv;  // 0 beyond W3: 0 * log2(1e-6) = 0
          a1 = __fmaf_rn(p, __log2f(p + 1e-6f), a1);
        }
        acc[g] = a1;
      }
    }
    warp_sum_multi<G>(acc, lane);
    if ((lane & (kOwn - 1)) == 0) {
      const int g = lane / kOwn;
      if (r0 + g < W2) a.out_l[slab * W2 + r0 + g] = CONF ? 1.0f + acc[0] * a.inv_log2_w3 : __fdividef(acc[0], sum[0]);
    }
  }

  // ---- merge the column states of the 8 warps
  float* sm_m = vr_smem + (size_t)warp * 3 * W3;
  float* sm_s = sm_m + W3;
  float* sm_t = sm_s + W3;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = lane + 32 * k;
    if (c < W3) { sm_m[c] = cm[k]; sm_s[c] = cs[k]; if (!CONF) sm_t[c] = ct[k]; }
  }
  __syncthreads();
  for (int c = tid; c < W3; c += kVrWarps * 32) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kVrWarps; ++w) M = fmaxf(M, vr_smem[(size_t)w * 3 * W3 + c]);
    float S = 0.f, T = 0.f;
#pragma unroll
    for (int w = 0; w < kVrWarps; ++w) {
      const float* q = vr_smem + (size_t)w * 3 * W3;
      const float f = ex2_approx((q[c] - M) * kLog2e);  // warps that saw no row hold m = -inf, s = 0
      S = __fmaf_rn(q[W3 + c], f, S);
      if (!CONF) T = __fmaf_rn(q[2 * W3 + c], f, T);
    }
    if (!CONF) a.out_r[slab * W3 + c] = T / S;
    else { vr_smem[c] = M; vr_smem[W3 + c] = 1.0f / S; }  // warp 0's m / s slots now hold the final (M, 1/S)
  }
  if (!CONF) return;
  __syncthreads();
  // ---- CONF, column direction: second sweep (L2) with the final (M, 1/S) of every column
  float cml[KMAX], cinv[KMAX], cacc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = lane + 32 * k;
    cml[k] = c < W3 ? vr_smem[c] * kLog2e : 0.f;
    cinv[k] = c < W3 ? vr_smem[W3 + c] : 0.f;
    cacc[k] = 0.f;
  }
  __syncthreads();  // everyone has read (M, 1/S) before the slots are reused for the partial sums
  for (int r0 = warp * G; r0 < W2; r0 += kVrWarps * G) {
    float x[G][KMAX];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const bool rv = r0 + g < W2;
      const float* row = src + (long long)(r0 + g) * W3;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = lane + 32 * k;
        x[g][k] = (rv && c < W3) ? __ldg(row + c) : -INFINITY;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const float p = ex2_approx(__fmaf_rn(x[g][k], kLog2e, -cml[k])) * cinv[k];
        cacc[k] = __fmaf_rn(p, __log2f(p + 1e-6f), cacc[k]);
      }
    }
  }
  float* sm_a = vr_smem + (size_t)warp * 3 * W3;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = lane + 32 * k;
    if (c < W3) sm_a[c] = cacc[k];
  }
  __syncthreads();
  for (int c = tid; c < W3; c += kVrWarps * 32) {
    float A = 0.f;
#pragma unroll
    for (int w = 0; w < kVrWarps; ++w) A += vr_smem[(size_t)w * 3 * W3 + c];
    a.out_r[slab * W3 + c] = 1.0f + A * a.inv_log2_w2;
  }
}

template <bool CONF, int KMAX, int G>
static int launch_vreduce_k(const VReduceArgs& a, long long BH, cudaStream_t st, const char* what) {
  constexpr int kVrWarps = (KMAX <= 12 || KMAX > 24) ? 8 : 16;  // wide rows: more warps per slab (few slabs, long rows)
  const size_t smem = (size_t)kVrWarps * 3 * a.W3 * sizeof(float);
  auto kern = volume_reduce_kernel<CONF, KMAX, G, kVrWarps>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) SA_FAIL((int)e, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
  }
  kern<<<(unsigned)BH, kVrWarps * 32, smem, st>>>(a);
  return finish_launch(what);
}

template <bool CONF>
static int launch_vreduce(VReduceArgs& a, long long BH, cudaStream_t st, const char* what) {
  const int kk = (a.W3 + 31) / 32;  // columns per lane
  if (kk <= 4) return launch_vreduce_k<CONF, 4, 4>(a, BH, st, what);
  if (kk <= 6) return launch_vreduce_k<CONF, 6, 4>(a, BH, st, what);
  if (kk <= 8) return launch_vreduce_k<CONF, 8, 4>(a, BH, st, what);
  if (kk <= 10) return launch_vreduce_k<CONF, 10, 4>(a, BH, st, what);
  if (kk <= 12) return launch_vreduce_k<CONF, 12, 2>(a, BH, st, what);
  if (kk <= 16) return launch_vreduce_k<CONF, 16, 2>(a, BH, st, what);
  if (kk <= 24) return launch_vreduce_k<CONF, 24, 1>(a, BH, st, what);
  return launch_vreduce_k<CONF, 32, 1>(a, BH, st, what);
}

static int vreduce_common(const float* vol, int64_t BH, int W2, int W3, float* out_l, float* out_r, VReduceArgs& a,
                          const char* what) {
  SA_REQUIRE(vol && out_l && out_r, SA_E_INVALID, "%s: null pointer", what);
  SA_REQUIRE(BH > 0 && BH < (1ll << 31) && W2 > 0 && W3 > 0, SA_E_INVALID, "%s: sizes must be positive", what);
  SA_REQUIRE(W3 <= 1024, SA_E_UNSUPPORTED, "%s: W3 = %d > 1024 is not covered", what, W3);
  a.vol = vol; a.out_l = out_l; a.out_r = out_r; a.W2 = W2; a.W3 = W3;
  a.inv_log2_w2 = (float)(1.0 / log2((double)W2));
  a.inv_log2_w3 = (float)(1.0 / log2((double)W3));
  return 0;
}

}  // namespace sa

extern "C" int sa_volume_softargmax(const float* vol, int64_t BH, int W2, int W3, float* disp_left, float* disp_right,
                                    void* stream) {
  using namespace sa;
  VReduceArgs a = {};
  int rc = vreduce_common(vol, BH, W2, W3, disp_left, disp_right, a, "sa_volume_softargmax");
  if (rc) return rc;
  return launch_vreduce<false>(a, BH, (cudaStream_t)stream, "sa_volume_softargmax");
}

extern "C" int sa_volume_entropy_conf(const float* vol, int64_t BH, int W2, int W3, float* conf_left, float* conf_right,
                                      void* stream) {
  using namespace sa;
  VReduceArgs a = {};
  int rc = vreduce_common(vol, BH, W2, W3, conf_left, conf_right, a, "sa_volume_entropy_conf");
  if (rc) return rc;
  SA_REQUIRE(W2 > 1 && W3 > 1, SA_E_INVALID, "sa_volume_entropy_conf: widths must exceed 1 (log2(W) divides)");
  return launch_vreduce<true>(a, BH, (cudaStream_t)stream, "sa_volume_entropy_conf");
}
