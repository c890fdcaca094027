// Line-packed pyramid + lookup (A3 + A4 for num_levels = 4, radius = 4, W3 % 8 == 0 - the model's
// configuration; reference: models/stereoanywhere/corr.py:76-115).
//
// Why: on B200 every L2 miss fetches a whole 128-byte line from HBM (measured: random 16..128-byte
// reads all cost 128 B of dram__bytes_read).  A lookup needs four 10-float windows per pixel, one per
// level, each in a different row of a different array: >= 4 lines (~590 B measured) for 160 useful
// bytes.  The packed layout stores, for every pixel row and every block q of 8 level-0 columns, ONE
// line that is sufficient for any x with floor(x) in [8q, 8q+8):
//     slots  0..16  L0[8q-4 .. 8q+12]
//     slots 17..21  L1[4q-4], L1[4q-3], L1[4q+6], L1[4q+7], L1[4q+8]
//     slots 22..26  L2[2q-4], L2[2q-3], L2[2q+4], L2[2q+5], L2[2q+6]
//     slots 27..31  L3[ q-4], L3[ q-3], L3[ q+3], L3[ q+4], L3[ q+5]
// The remaining window entries are re-derived in registers with the pyramid's own formula
// 0.5*(a+b) - L1[4q-2..4q+5] from the stored L0, L2[2q-2..2q+3] from L1, L3[q-2..q+2] from L2 - so the
// result is bit-identical to pooling first and looking up afterwards.  Entries outside [0, W_i)
// are stored as zeros (grid_sample's zero padding), blocks q = -5 .. W/8+3 cover every x for which
// any tap of any level is inside the image.  One lookup = one line per (pixel, volume): 128 B read
// + 144 B written, below the 308 B/px "algorithmic" figure of the unpacked formulation.
#include <stdlib.h>
#include <string.h>

#include "pack_stream.cuh"
#include "sa_common.cuh"
#include "tc_common.cuh"

namespace sa {

constexpr int kQMin = -5;                       // first block index
__host__ __device__ inline int packed_blocks(int W) { return W / 8 + 9; }  // q in [-5, W/8 + 3]

// slot -> (level, offset): the entry stored in slot s of block q is L_level[(q << (3 - level)) + off]
__host__ __device__ inline void slot_map(int s, int& level, int& off) {
  if (s < 17) {
    level = 0;
    off = s - 4;
    return;
  }
  const int t = s - 17;
  level = 1 + t / 5;
  const int r = t % 5;
  const int hi = level == 1 ? 6 : (level == 2 ? 4 : 3);
  off = r < 2 ? r - 4 : hi + (r - 2);
}

// ---------------------------------------------------------------------------------------------
// pack: one warp per volume row.  The (optionally truncation-masked) row and its three pooled levels
// are formed in shared memory, then gathered into nblk lines with coalesced 128-bit stores.
// ---------------------------------------------------------------------------------------------
struct PackArgs {
  const float* src;
  // alternative source: the C=3 normals volume computed on the fly (A2), row = (b*H + h)*W2 + w2
  const float* nl;
  const float* nr;
  int H, W2;
  float divisor, inv_divisor, post_scale;
  float* packed;
  long long rows;
  int W;
  const float* disp;  // truncation (optional)
  const float* conf;
  float gain, one_minus_gain;
  int w2_size;
};

constexpr int kPackMaxV4 = 3;  // float4 per lane and row held in registers while prefetching (W <= 384)

template <bool TRUNC, bool NORMALS>
__global__ void __launch_bounds__(256, NORMALS ? 2 : 3) pack_kernel(const PackArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int W = a.W;
  const int W4 = W >> 2;
  const int per_warp = (W + W / 2 + W / 4 + W / 8 + 3) & ~3;  // keep every warp's slab 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s0 = sm + warp * per_warp;
  float* s1 = s0 + W;
  float* s2 = s1 + W / 2;
  float* s3 = s2 + W / 4;
  const int nblk = packed_blocks(W);
  // second half of every line (slots 16..31): a lane always owns the same 4 slots - resolve them once
  int lv[4], off[4], lbase[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    slot_map(16 + (lane & 3) * 4 + e, lv[e], off[e]);
    lbase[e] = lv[e] == 0 ? 0 : (lv[e] == 1 ? W : (lv[e] == 2 ? W + W / 2 : W + W / 2 + W / 4));
  }
  const bool prefetch = !NORMALS && (W4 <= 32 * kPackMaxV4);
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  // Volume source: rows round-robin over the warps (the next row is prefetched while this one is packed).
  // Normals source: a CONTIGUOUS chunk of rows per warp, so that the right normals of the image row (b,h) - the
  // same for all its W2 volume rows - stay in registers and only three scalars change from row to row.
  const bool resident = NORMALS && (W4 <= 32 * kPackMaxV4);
  const long long rows_per_warp = (a.rows + warps_total - 1) / warps_total;
  long long row = NORMALS ? gwarp * rows_per_warp : gwarp;
  const long long row_end = NORMALS ? min(a.rows, row + rows_per_warp) : a.rows;
  const long long row_step = NORMALS ? 1 : warps_total;
  float4 nxt[kPackMaxV4];
  float4 rr0[kPackMaxV4], rr1[kPackMaxV4], rr2[kPackMaxV4];
  long long cur_bh = -1;
  if (prefetch && row < row_end) {
#pragma unroll
    for (int i = 0; i < kPackMaxV4; ++i)
      if (lane + 32 * i < W4) nxt[i] = ld_stream_v4(a.src + row * W + 4 * (lane + 32 * i));
  }
  for (; row < row_end; row += row_step) {
    float c = 0.f, centre = 0.f, omc = 1.f;
    if (TRUNC) {
      c = __ldg(a.conf + row);
      omc = 1.0f - c;
      centre = (float)(int)(row % a.w2_size) - __ldg(a.disp + row);
    }
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    const float* nrp = nullptr;
    long long plane3 = 0;
    if (NORMALS) {  // same FMA order as corr_simt_kernel: bit-identical to sa_corr_fp32 + sa_pack_pyramid
      const long long bh = row / a.W2;
      const int w2 = (int)(row - bh * a.W2);
      const long long b = bh / a.H, h = bh - b * a.H;
      const long long plane2 = (long long)a.H * a.W2;
      plane3 = (long long)a.H * W;
      const float* nlp = a.nl + (b * 3 * a.H + h) * a.W2 + w2;
      n0 = __ldg(nlp); n1 = __ldg(nlp + plane2); n2 = __ldg(nlp + 2 * plane2);
      nrp = a.nr + (b * 3 * a.H + h) * (long long)W;
      if (resident && bh != cur_bh) {
        cur_bh = bh;
#pragma unroll
        for (int i = 0; i < kPackMaxV4; ++i) {
          if (lane + 32 * i < W4) {
            rr0[i] = __ldg(reinterpret_cast<const float4*>(nrp + 4 * (lane + 32 * i)));
            rr1[i] = __ldg(reinterpret_cast<const float4*>(nrp + plane3 + 4 * (lane + 32 * i)));
            rr2[i] = __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + 4 * (lane + 32 * i)));
          }
        }
      }
    }
    auto put_row = [&](int v, float4 q) {
      if (TRUNC) {
        trunc_mask_mul4(q, centre, (float)(4 * v), c, omc, a.gain, a.one_minus_gain);
      }
      *reinterpret_cast<float4*>(s0 + 4 * v) = q;
      *reinterpret_cast<float2*>(s1 + 2 * v) = make_float2((q.x + q.y) * 0.5f, (q.z + q.w) * 0.5f);
    };
    auto mono4 = [&](const float4& r0, const float4& r1, const float4& r2) {
      float4 q;
      q.x = div_const(fmaf(n2, r2.x, fmaf(n1, r1.x, fmaf(n0, r0.x, 0.f))), a.divisor, a.inv_divisor) * a.post_scale;
      q.y = div_const(fmaf(n2, r2.y, fmaf(n1, r1.y, fmaf(n0, r0.y, 0.f))), a.divisor, a.inv_divisor) * a.post_scale;
      q.z = div_const(fmaf(n2, r2.z, fmaf(n1, r1.z, fmaf(n0, r0.z, 0.f))), a.divisor, a.inv_divisor) * a.post_scale;
      q.w = div_const(fmaf(n2, r2.w, fmaf(n1, r1.w, fmaf(n0, r0.w, 0.f))), a.divisor, a.inv_divisor) * a.post_scale;
      return q;
    };
    if (NORMALS && resident) {
#pragma unroll
      for (int i = 0; i < kPackMaxV4; ++i)
        if (lane + 32 * i < W4) put_row(lane + 32 * i, mono4(rr0[i], rr1[i], rr2[i]));
    } else if (NORMALS) {
      for (int v = lane; v < W4; v += 32) {
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(nrp + 4 * v));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(nrp + plane3 + 4 * v));
        const float4 r2 = __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + 4 * v));
        put_row(v, mono4(r0, r1, r2));
      }
    } else if (prefetch) {
#pragma unroll
      for (int i = 0; i < kPackMaxV4; ++i)
        if (lane + 32 * i < W4) put_row(lane + 32 * i, nxt[i]);
      const long long nrow = row + warps_total;  // the next row's loads fly while this row is packed
      if (nrow < row_end) {
#pragma unroll
        for (int i = 0; i < kPackMaxV4; ++i)
          if (lane + 32 * i < W4) nxt[i] = ld_stream_v4(a.src + nrow * W + 4 * (lane + 32 * i));
      }
    } else {
      for (int v = lane; v < W4; v += 32) put_row(v, ld_stream_v4(a.src + row * W + 4 * v));
    }
    __syncwarp();
    for (int j = lane; j < W / 4; j += 32) s2[j] = (s1[2 * j] + s1[2 * j + 1]) * 0.5f;
    __syncwarp();
    for (int j = lane; j < W / 8; j += 32) s3[j] = (s2[2 * j] + s2[2 * j + 1]) * 0.5f;
    __syncwarp();
    float* dst = a.packed + row * (long long)nblk * 32;
    // first half of every line = L0[8q-4 .. 8q+11]: four aligned float4 copies of the row itself
    for (int v = lane; v < nblk * 4; v += 32) {
      const int q = (v >> 2) + kQMin;
      const int col = 8 * q - 4 + 4 * (v & 3);
      const float4 o = (col >= 0 && col < W) ? *reinterpret_cast<const float4*>(s0 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      st_stream_v4(dst + (v >> 2) * 32 + 4 * (v & 3), o);
    }
    // second half: L0[8q+12] and the five border entries of each pooled level
    for (int v = lane; v < nblk * 4; v += 32) {
      const int q = (v >> 2) + kQMin;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = q * (8 >> lv[e]) + off[e];
        o[e] = (idx >= 0 && idx < (W >> lv[e])) ? s0[lbase[e] + idx] : 0.0f;
      }
      st_stream_v4(dst + (v >> 2) * 32 + 16 + 4 * (v & 3), make_float4(o[0], o[1], o[2], o[3]));
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// lookup from the packed layout
// ---------------------------------------------------------------------------------------------
struct PLookupArgs {
  const float* packed[2];
  float* out[2];
  const float* coords;
  long long coords_bstride;
  int HW, W3, nblk;
  // on-the-fly mono volume (OTF >= 0): unit normals [B,3,H,Wimg] / [B,3,H,W3], A2's divisor and post_scale
  const float* nl;
  const float* nr;
  int H, Wimg;
  float divisor, inv_divisor, post_scale;
  int out_evict_first;  // TMA output stores carry the L2 evict_first hint (set by the launcher, see launch_packed_tt)
  int reverse;          // walk the pixel tiles backwards (alternate launches, see next_direction)
  // factored mono volume (FV >= 0): packed[FV] holds the packed pyramid of the RIGHT NORMAL MAP's rows,
  // [(b*3 + c)*H + h][nblk][32]; nl / H / Wimg / divisor / inv_divisor / post_scale as above
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}

template <int N>
__device__ __forceinline__ void shift_if(float (&v)[N], bool on, int by, int keep) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < keep && i + by < N) v[i] = on ? v[i + by] : v[i];
}

// One staged line (32 floats, chunk c of pixel p at chunk c ^ (p & 7)) -> the window entries of the four levels:
// L0[8q-4..8q+12], L1[4q-4..4q+8], L2[2q-4..2q+6], L3[q-4..q+5]; the entries that are not stored are re-derived
// with the pyramid's own 0.5 (a + b).
__device__ __forceinline__ void line_levels(const float* src, int p, float (&l0)[17], float (&l1)[13], float (&l2)[11],
                                            float (&l3)[10]) {
  float ln[32];
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const float4 t = *reinterpret_cast<const float4*>(src + ((ch ^ (p & 7)) << 2));
    ln[4 * ch] = t.x; ln[4 * ch + 1] = t.y; ln[4 * ch + 2] = t.z; ln[4 * ch + 3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 17; ++i) l0[i] = ln[i];
  // slots 17..31: full windows of levels 1..3 (relative index 0 = first stored entry of the level)
  l1[0] = ln[17]; l1[1] = ln[18]; l1[10] = ln[19]; l1[11] = ln[20]; l1[12] = ln[21];
#pragma unroll
  for (int t = 0; t < 8; ++t) l1[2 + t] = (l0[2 * t] + l0[2 * t + 1]) * 0.5f;
  l2[0] = ln[22]; l2[1] = ln[23]; l2[8] = ln[24]; l2[9] = ln[25]; l2[10] = ln[26];
#pragma unroll
  for (int t = 0; t < 6; ++t) l2[2 + t] = (l1[2 * t] + l1[2 * t + 1]) * 0.5f;
  l3[0] = ln[27]; l3[1] = ln[28]; l3[7] = ln[29]; l3[8] = ln[30]; l3[9] = ln[31];
#pragma unroll
  for (int t = 0; t < 5; ++t) l3[2 + t] = (l2[2 * t] + l2[2 * t + 1]) * 0.5f;
}

// The same from a line in 16-bit storage (HK = 1: fp16, 2: bf16; csrc/corr_pack_tcgen05.cu, sa_corr_pack_tf32_half):
// 64 bytes in the half of the pixel's 128-byte slot selected by p & 1, chunk c at c ^ ((p >> 1) & 3).
template <int HK>
__device__ __forceinline__ void line_levels_half(const float* slot, int p, float (&l0)[17], float (&l1)[13], float (&l2)[11],
                                                 float (&l3)[10]) {
  float ln[32];
  const uint8_t* src = reinterpret_cast<const uint8_t*>(slot) + ((p & 1) << 6);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint4 t = *reinterpret_cast<const uint4*>(src + ((ch ^ ((p >> 1) & 3)) << 4));
    unpack_half2<HK>(t.x, ln[8 * ch], ln[8 * ch + 1]);
    unpack_half2<HK>(t.y, ln[8 * ch + 2], ln[8 * ch + 3]);
    unpack_half2<HK>(t.z, ln[8 * ch + 4], ln[8 * ch + 5]);
    unpack_half2<HK>(t.w, ln[8 * ch + 6], ln[8 * ch + 7]);
  }
#pragma unroll
  for (int i = 0; i < 17; ++i) l0[i] = ln[i];
  l1[0] = ln[17]; l1[1] = ln[18]; l1[10] = ln[19]; l1[11] = ln[20]; l1[12] = ln[21];
#pragma unroll
  for (int t = 0; t < 8; ++t) l1[2 + t] = (l0[2 * t] + l0[2 * t + 1]) * 0.5f;
  l2[0] = ln[22]; l2[1] = ln[23]; l2[8] = ln[24]; l2[9] = ln[25]; l2[10] = ln[26];
#pragma unroll
  for (int t = 0; t < 6; ++t) l2[2 + t] = (l1[2 * t] + l1[2 * t + 1]) * 0.5f;
  l3[0] = ln[27]; l3[1] = ln[28]; l3[7] = ln[29]; l3[8] = ln[30]; l3[9] = ln[31];
#pragma unroll
  for (int t = 0; t < 5; ++t) l3[2 + t] = (l2[2 * t] + l2[2 * t + 1]) * 0.5f;
}

// The 4 x 9 taps at x from the window entries (which are consumed): put(channel, value).
template <class Put>
__device__ __forceinline__ void blend_windows(float (&l0)[17], float (&l1)[13], float (&l2)[11], float (&l3)[10], float x,
                                              Put put) {
  const float flx = fminf(fmaxf(floorf(x), -1.0e6f), 1.0e6f);
  const int x0 = (int)flx;
  // window starts inside the stored ranges: a0 in [0,8), a1 in [0,4), a2 in [0,2), a3 = 0
  {
    shift_if(l0, (x0 & 1) != 0, 1, 16);
    shift_if(l0, (x0 & 2) != 0, 2, 14);
    shift_if(l0, (x0 & 4) != 0, 4, 10);
    const float f = x - floorf(x);
#pragma unroll
    for (int k = 0; k < 9; ++k) put(k, blend(l0[k], l0[k + 1], f));
  }
  {
    const int x1 = x0 >> 1;
    shift_if(l1, (x1 & 1) != 0, 1, 12);
    shift_if(l1, (x1 & 2) != 0, 2, 10);
    const float xs = x * 0.5f;
    const float f = xs - floorf(xs);
#pragma unroll
    for (int k = 0; k < 9; ++k) put(9 + k, blend(l1[k], l1[k + 1], f));
  }
  {
    const int x2 = x0 >> 2;
    shift_if(l2, (x2 & 1) != 0, 1, 10);
    const float xs = x * 0.25f;
    const float f = xs - floorf(xs);
#pragma unroll
    for (int k = 0; k < 9; ++k) put(18 + k, blend(l2[k], l2[k + 1], f));
  }
  {
    const float xs = x * 0.125f;
    const float f = xs - floorf(xs);
#pragma unroll
    for (int k = 0; k < 9; ++k) put(27 + k, blend(l3[k], l3[k + 1], f));
  }
}

// OTF = index of a volume that is not read from a packed array but COMPUTED from the C = 3 normal maps (the mono
// volume of stereoanywhere.py:136 when it is looked up as it is, i.e. use_aggregate_mono_vol off), or -1.  The
// thread of such a (pixel, volume) forms the 80 level-0 values L0[8q-32 .. 8q+47] its line is made of - three
// FMAs, the division and the 1.73 scale in pack_kernel<normals>'s order - and pools them with the pyramid's own
// 0.5 (a + b): exactly the 17 + 13 + 11 + 10 window entries the packed path holds after its derivation step, bit
// for bit.  ~900 instructions per pixel instead of a 128-byte line read, and no mono pack pass at all.
//
// FV = index of a volume served in FACTORED form, or -1.  The mono volume has rank 3 and both the pyramid and the
// packed layout are linear in the volume, so the packed line of pixel (b,h,w2) is the combination, with the three
// left normals as coefficients, of the packed lines of the three right-normal rows (b,c,h) - a 14 MB array at
// KITTI size that stays in L2, instead of the 1.47 GB packed mono volume.  The staging lanes load the three
// 16-byte chunks, combine them with the pixel's left normal pre-multiplied by post_scale / divisor (one FMUL and
// two FFMA per entry) and store the chunk; everything after staging is the packed path unchanged.  Differences
// from the packed mono volume are fp32 rounding only (the scale is applied to the coefficients, and the 15 stored
// border entries of levels 1..3 are pooled before the contraction instead of after it): <= 4e-7 for unit normals.
// zero-filling form of cp.async: `bytes` = 16 copies the chunk, 0 writes sixteen zero bytes without reading (a pixel
// whose window lies wholly outside the row) - no branch in the staging loop
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, unsigned bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(bytes)
               : "memory");
}

// All global addressing below is `uniform 64-bit base + 32-bit unsigned index` in units of 16 bytes (one chunk of a
// line, one float4 of the output): a packed array of up to 64 GB.  The first version formed every address in 64 bits
// per thread - 370 of its 950 SASS instructions were integer address arithmetic in a kernel that is issue-bound
// (DESIGN 7c); see profiles/r2 for the before / after instruction mix.
// Register budgets (measured, profiles/r2/lookup_variants_sweep.txt): 56 for the TMA-store form, 48 for the 128-bit-store
// form (21 CTAs of 64 threads per SM instead of 18); the on-the-fly mono form forms 80 values per thread and keeps
// its free budget.
#ifndef SA_LOOKUP_REGS_TMA
#define SA_LOOKUP_REGS_TMA 56
#endif
#ifndef SA_LOOKUP_REGS_VEC
#define SA_LOOKUP_REGS_VEC 48
#endif
// H0 = storage of volume 0's packed array: 0 fp32 lines (128 B), 1 / 2 fp16 / bf16 lines (64 B, half the staging copies)
template <int NV, int TILE, int OTF, int FV, bool TMA, int H0>
__global__ void __maxnreg__(OTF >= 0 ? 128 : (TMA ? SA_LOOKUP_REGS_TMA : SA_LOOKUP_REGS_VEC)) lookup_packed_kernel(const PLookupArgs a, const __grid_constant__ CUtensorMap map_o0,
                                                     const __grid_constant__ CUtensorMap map_o1) {
  static_assert(OTF < 0 || FV < 0, "one special mono form at a time");
  constexpr int THREADS = NV * TILE;
  constexpr int NC = 36;
  // TMA: dense [channel][pixel] tile, stored by the TMA unit; else padded rows for the transposed 128-bit re-read
  constexpr int SP = TMA ? TILE : TILE + 4;
  constexpr int STAGE_FLOATS = NV * TILE * 32;
  constexpr int OUT_FLOATS = NV * NC * SP;
  constexpr int BUF_FLOATS = STAGE_FLOATS > OUT_FLOATS ? STAGE_FLOATS : OUT_FLOATS;
  extern __shared__ __align__(128) float smem[];
  float* buf = smem;                                        // staging lines, later the output tile
  float* s_x = smem + BUF_FLOATS;                           // [TILE] x coordinate
  unsigned* s_line = reinterpret_cast<unsigned*>(s_x + TILE);  // [TILE] index of the pixel's line in 16-byte chunks, or ~0u
  // FV >= 0: per pixel {k n0, k n1, k n2, chunk index of line (c = 0, h, blk) of the packed right normals}
  float4* s_n = reinterpret_cast<float4*>(s_x + 2 * TILE);

  const int tid = threadIdx.x;
  const unsigned b = a.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const unsigned hw0 = (a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * TILE;
  const int npx = min(TILE, a.HW - (int)hw0);

  if (tid < TILE) {
    float x = 0.f;
    unsigned line = ~0u;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    int blk = -1;
    if (tid < npx) {
      x = __ldg(a.coords + (size_t)b * a.coords_bstride + hw0 + tid);
      if (FV >= 0) {  // issued together with the coordinate load: they do not depend on it
        const unsigned plane = (unsigned)a.H * a.Wimg;
        const float* nlp = a.nl + (size_t)(b * 3 * plane + hw0 + tid);  // h * Wimg + w2 == hw
        n0 = __ldg(nlp); n1 = __ldg(nlp + plane); n2 = __ldg(nlp + 2 * plane);
      }
      const float fl = fminf(fmaxf(floorf(x), -1.0e6f), 1.0e6f);
      const int q = ((int)fl >> 3) - kQMin;
      if (q >= 0 && q < a.nblk) {
        blk = q;
        line = ((b * (unsigned)a.HW + hw0 + tid) * (unsigned)a.nblk + (unsigned)q) * 8u;
      }
    }
    s_x[tid] = x;
    s_line[tid] = line;
    if (FV >= 0) {
      // a pixel without a line (blk < 0) reads line 0 with zero coefficients: no branch in the staging loop.
      // Image row of the pixel: the CTA's first pixel is in row h0 (uniform division), a pixel of the tile at most
      // four rows further down (TILE <= 32 + ... and Wimg >= 8) - no per-thread division.
      const float k = blk >= 0 ? a.post_scale * a.inv_divisor : 0.f;
      const unsigned h0 = hw0 / (unsigned)a.Wimg;
      const unsigned t = hw0 - h0 * (unsigned)a.Wimg + tid;
      unsigned hh = h0;
      for (unsigned w = a.Wimg; w <= t; w += a.Wimg) ++hh;
      const unsigned off = blk >= 0 ? (((b * 3u) * a.H + hh) * (unsigned)a.nblk + (unsigned)blk) * 8u : 0u;
      s_n[tid] = make_float4(n0 * k, n1 * k, n2 * k, __uint_as_float(off));
    }
  }
  __syncthreads();

  // ---- stage one line per (pixel, volume): 8 lanes x 16 B, chunk c of pixel p lands at chunk c ^ (p & 7).
  // Thread tid copies chunk ch = tid & 7 of the units (tid >> 3) + n * THREADS/8, n < 8: the volume of a unit is
  // a compile-time function of n and its pixel one of 8/NV values, so a line index is read 8/NV times per thread
  // and shared by the volumes.
  {
    constexpr int UPS = THREADS / 8;   // units per step
    constexpr int SPV = 8 / NV;        // steps per volume
    const unsigned ch = tid & 7, u0 = tid >> 3;
    const float4* p0 = reinterpret_cast<const float4*>(a.packed[0]);
    const float4* p1 = reinterpret_cast<const float4*>(NV == 2 ? a.packed[1] : a.packed[0]);
#pragma unroll
    for (int m = 0; m < SPV; ++m) {
      const unsigned pm = u0 + m * UPS;
      const unsigned line = s_line[pm];
      const unsigned ok = line != ~0u ? 16u : 0u;
      const unsigned idx = (line != ~0u ? line : 0u) + ch;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (v == OTF || v == FV) continue;
        if (H0 != 0 && v == 0) {
          // 16-bit lines: four chunks; the line index counts 16-byte chunks of a 128-byte line - halve it
          if (ch < 4) {
            float* dst = buf + pm * 32 + ((((pm & 1) << 2) | (ch ^ ((pm >> 1) & 3))) << 2);
            cp_async16_zfill(dst, p0 + ((line != ~0u ? line : 0u) >> 1) + ch, ok);
          }
          continue;
        }
        float* dst = buf + (v * TILE + pm) * 32 + ((ch ^ (pm & 7)) << 2);
        cp_async16_zfill(dst, (v ? p1 : p0) + idx, ok);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (FV >= 0) {
    // the factored volume's chunks while the copies above are in flight: all loads first, then the arithmetic
    constexpr int UPS = THREADS / 8, SPV = 8 / NV;
    const unsigned ch = tid & 7, u0 = tid >> 3;
    const unsigned cplane = (unsigned)a.H * a.nblk * 8u;  // chunks between the channels of one (b, h)
    const float4* rp = reinterpret_cast<const float4*>(a.packed[FV >= 0 ? FV : 0]);
    float4 r[SPV][3], n[SPV];
#pragma unroll
    for (int m = 0; m < SPV; ++m) {
      n[m] = s_n[u0 + m * UPS];
      const unsigned off = __float_as_uint(n[m].w) + ch;
#pragma unroll
      for (int c = 0; c < 3; ++c) r[m][c] = __ldg(rp + (off + c * cplane));
    }
#pragma unroll
    for (int m = 0; m < SPV; ++m) {
      const int pm = u0 + m * UPS;
      float4 o;
      o.x = fmaf(n[m].z, r[m][2].x, fmaf(n[m].y, r[m][1].x, n[m].x * r[m][0].x));
      o.y = fmaf(n[m].z, r[m][2].y, fmaf(n[m].y, r[m][1].y, n[m].x * r[m][0].y));
      o.z = fmaf(n[m].z, r[m][2].z, fmaf(n[m].y, r[m][1].z, n[m].x * r[m][0].z));
      o.w = fmaf(n[m].z, r[m][2].w, fmaf(n[m].y, r[m][1].w, n[m].x * r[m][0].w));
      *reinterpret_cast<float4*>(buf + ((FV >= 0 ? FV : 0) * TILE + pm) * 32 + ((ch ^ (pm & 7)) << 2)) = o;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- thread (pixel p, volume v): the window entries of the four levels
  const int p = tid % TILE, v = tid / TILE;
  float l0[17], l1[13], l2[11], l3[10];  // L0[8q-4..8q+12], L1[4q-4..4q+8], L2[2q-4..2q+6], L3[q-4..q+5]
  if (OTF >= 0 && v == OTF) {
#pragma unroll
    for (int i = 0; i < 17; ++i) l0[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 13; ++i) l1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 11; ++i) l2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) l3[i] = 0.f;
    const unsigned line_p = s_line[p];
    if (line_p != ~0u) {
      const int blk = (int)((line_p >> 3) % (unsigned)a.nblk);
      const int c0 = 8 * (blk + kQMin) - 32;
      const int hw = (int)hw0 + p;
      const int hh = hw / a.Wimg, w2 = hw - hh * a.Wimg;
      const long long plane2 = (long long)a.H * a.Wimg, plane3 = (long long)a.H * a.W3;
      const float* nlp = a.nl + ((long long)b * 3 * a.H + hh) * a.Wimg + w2;
      const float n0 = __ldg(nlp), n1 = __ldg(nlp + plane2), n2 = __ldg(nlp + 2 * plane2);
      const float* nrp = a.nr + ((long long)b * 3 * a.H + hh) * a.W3;
#pragma unroll
      for (int m = 0; m < 10; ++m) {
        const int col = c0 + 8 * m;
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        if (col >= 0 && col < a.W3) {  // W3 % 8 == 0: a group of 8 columns is inside or outside as a whole
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            // (__fmul_rn: the product must be rounded on its own, as it is when pack_kernel stores the value - the
          // compiler would otherwise contract it into the pooling adds below)
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(nrp + col + 4 * hlf));
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(nrp + plane3 + col + 4 * hlf));
            const float4 r2 = __ldg(reinterpret_cast<const float4*>(nrp + 2 * plane3 + col + 4 * hlf));
            g[4 * hlf + 0] = __fmul_rn(div_const(fmaf(n2, r2.x, fmaf(n1, r1.x, fmaf(n0, r0.x, 0.f))), a.divisor, a.inv_divisor), a.post_scale);
            g[4 * hlf + 1] = __fmul_rn(div_const(fmaf(n2, r2.y, fmaf(n1, r1.y, fmaf(n0, r0.y, 0.f))), a.divisor, a.inv_divisor), a.post_scale);
            g[4 * hlf + 2] = __fmul_rn(div_const(fmaf(n2, r2.z, fmaf(n1, r1.z, fmaf(n0, r0.z, 0.f))), a.divisor, a.inv_divisor), a.post_scale);
            g[4 * hlf + 3] = __fmul_rn(div_const(fmaf(n2, r2.w, fmaf(n1, r1.w, fmaf(n0, r0.w, 0.f))), a.divisor, a.inv_divisor), a.post_scale);
          }
        }
        const float a10 = (g[0] + g[1]) * 0.5f, a11 = (g[2] + g[3]) * 0.5f, a12 = (g[4] + g[5]) * 0.5f, a13 = (g[6] + g[7]) * 0.5f;
        const float a20 = (a10 + a11) * 0.5f, a21 = (a12 + a13) * 0.5f;
        l3[m] = (a20 + a21) * 0.5f;                         // L3[q-4+m]
        if (m >= 2 && m <= 7) {                              // L2[2q-8+2m+j] -> index 2(m-2)+j of l2
          l2[2 * (m - 2)] = a20;
          if (2 * (m - 2) + 1 < 11) l2[2 * (m - 2) + 1] = a21;
        }
        if (m >= 3 && m <= 6) {                              // L1[4q-16+4m+j] -> index 4(m-3)+j of l1
          const float a1v[4] = {a10, a11, a12, a13};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (4 * (m - 3) + j < 13) l1[4 * (m - 3) + j] = a1v[j];
        }
        if (m == 3) {                                        // L0[8q-8+j], j = 4..7 -> l0[0..3]
#pragma unroll
          for (int j = 4; j < 8; ++j) l0[j - 4] = g[j];
        } else if (m == 4) {                                 // L0[8q+j] -> l0[4..11]
#pragma unroll
          for (int j = 0; j < 8; ++j) l0[4 + j] = g[j];
        } else if (m == 5) {                                 // L0[8q+8+j], j = 0..4 -> l0[12..16]
#pragma unroll
          for (int j = 0; j < 5; ++j) l0[12 + j] = g[j];
        }
      }
    }
  } else {
    if (H0 != 0 && v == 0)
      line_levels_half<H0 ? H0 : 1>(buf + tid * 32, p, l0, l1, l2, l3);
    else
      line_levels(buf + tid * 32, p, l0, l1, l2, l3);  // unit index == tid
  }
  __syncthreads();  // staging is dead: `buf` becomes the [channel][pixel] output tile

  blend_windows(l0, l1, l2, l3, s_x[p], [&](int c, float val) { buf[(v * NC + c) * SP + p] = val; });

  // ---- [channel][pixel] tile -> NCHW.  One TMA tensor store per volume (box {TILE pixels, 36 channels, 1} of the
  // map {HW, 36, B}; pixels beyond HW are clipped by the map) instead of 9 LDS.128 + 9 STG.128 per thread: the
  // kernel is bound by L1 / shared-memory wavefronts (ncu: l1tex 59 % busy, 31 % of the stall samples on the
  // shared-memory scoreboard / MIO queue), and those 18 instructions were 27 % of its wavefronts.
  if (TMA) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < NV) {
      // evict_first: the 69 MB a launch writes would otherwise sit dirty in the L2 and push out the packed lines and
      // right-normal rows that the next iteration reads again (coordinates move by a fraction of a pixel per
      // iteration) - 22.7 -> 21.1 us per dual lookup at KITTI size (profiles/r2/lookup_l2_policy_sweep.txt).  Only
      // while the output of a launch fits the L2 with room to spare: beyond that the hint costs 1-2 %.
      if (a.out_evict_first)
        tma_store_3d_hint(tid ? &map_o1 : &map_o0, buf + tid * NC * SP, (int)hw0, 0, (int)b, l2_evict_first_policy());
      else
        tma_store_3d(tid ? &map_o1 : &map_o0, buf + tid * NC * SP, (int)hw0, 0, (int)b);
      tma_commit();
      tma_wait_read<0>();   // the tile must outlive the store's read of it
    }
    return;
  }
  __syncthreads();
  if ((a.HW & 3) == 0) {
    // thread tid always stores pixels t .. t+3 of the channels c0 + 4 NV k, k < NC/4: one 32-bit float4 index per
    // volume, advanced by a constant stride
    constexpr int T4 = TILE / 4;
    static_assert(THREADS == 4 * NV * T4 && NC % 4 == 0, "store mapping");
    const int c0 = tid / T4, t = (tid % T4) * 4;
    if (t < npx) {
      const unsigned hw4 = (unsigned)a.HW >> 2;
      float4* o0 = reinterpret_cast<float4*>(a.out[0]);
      float4* o1 = reinterpret_cast<float4*>(NV == 2 ? a.out[1] : a.out[0]);
      const unsigned pix4 = b * NC * hw4 + ((hw0 + t) >> 2);
#pragma unroll
      for (int k = 0; k < NC / 4; ++k) {
        const int c = c0 + 4 * NV * k;          // 0 .. NV*NC-1
        const int vv = (NV == 2 && c >= NC) ? 1 : 0;
        const int cc = c - vv * NC;
        const float4 val = *reinterpret_cast<const float4*>(buf + c * SP + t);
        st_stream_v4(reinterpret_cast<float*>((vv ? o1 : o0) + (pix4 + cc * hw4)), val);
      }
    }
  } else {
    for (int idx = tid; idx < NV * NC * TILE; idx += THREADS) {
      const int c = idx / TILE, t = idx % TILE;
      if (t < npx) {
        const int vv = c / NC, cc = c - vv * NC;
        (vv ? a.out[1] : a.out[0])[((long long)b * NC + cc) * a.HW + hw0 + t] = buf[c * SP + t];
      }
    }
  }
}

// Output tensor map {HW, 36, B} with box {TILE, 36, 1}; encoding one costs ~1 us of host time, so the last few
// (pointer, geometry) combinations are kept per thread - torch's caching allocator hands the same blocks back.
static int output_map(CUtensorMap* m, const float* out, int B, int HW, int tile) {
  struct Entry { const float* p; int B, HW, tile; CUtensorMap m; };
  static thread_local Entry cache[16];
  static thread_local int next = 0;
  for (int i = 0; i < 16; ++i)
    if (cache[i].p == out && cache[i].B == B && cache[i].HW == HW && cache[i].tile == tile) {
      *m = cache[i].m;
      return 0;
    }
  EncodeTiledFn fn = encode_fn();
  SA_REQUIRE(fn != nullptr, SA_E_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable (no driver?)");
  cuuint64_t dims[3] = {(cuuint64_t)HW, 36, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 36 * 4};
  cuuint32_t box[3] = {(cuuint32_t)tile, 36, 1};
  cuuint32_t ones[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(out), dims, str, box, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SA_REQUIRE(r == CUDA_SUCCESS, SA_E_INVALID, "cuTensorMapEncodeTiled(lookup output) failed with CUresult %d", (int)r);
  cache[next] = Entry{out, B, HW, tile, *m};
  next = (next + 1) & 15;
  return 0;
}

// Consecutive lookups of one volume walk its pixel tiles in OPPOSITE directions.  The GRU's coordinates move by a
// fraction of a pixel per iteration, so a launch re-reads mostly the lines (and right-normal rows) the previous one
// read; walking forwards every time, the lines a launch needs first are the ones the previous launch read longest
// ago - the first to have left the L2 (the cyclic-scan worst case of an LRU-like cache).  Walking back and forth, a
// launch starts where its predecessor stopped: 21.2 -> 19.9 us per dual lookup at KITTI size (two packed volumes:
// 23.3 -> 21.7).  Results do not depend on the direction.  Parity per volume (keyed by its packed array) and per
// thread; under CUDA-graph capture every captured launch keeps the direction it was captured with.
// SA_B200_LOOKUP_ALTERNATE=0 walks forwards always.
static int next_direction(const void* key) {
  static const bool on = !(getenv("SA_B200_LOOKUP_ALTERNATE") && atoi(getenv("SA_B200_LOOKUP_ALTERNATE")) == 0);
  if (!on) return 0;
  struct Entry { const void* p; unsigned n; };
  static thread_local Entry tab[8];
  static thread_local int next = 0;
  for (int i = 0; i < 8; ++i)
    if (tab[i].p == key) return (int)(tab[i].n++ & 1u);
  tab[next] = Entry{key, 1u};
  next = (next + 1) & 7;
  return 0;
}

int lookup_next_direction(const void* key) { return next_direction(key); }  // for csrc/lookup_conv.cu

template <int NV, int TILE, int OTF, int FV, bool TMA, int H0>
static int launch_packed_tt(PLookupArgs a, int B, cudaStream_t st) {
  constexpr int NC = 36, SP = TMA ? TILE : TILE + 4;
  constexpr int buf_floats = (NV * TILE * 32 > NV * NC * SP) ? NV * TILE * 32 : NV * NC * SP;
  const size_t smem = (size_t)(buf_floats + (FV >= 0 ? 6 : 2) * TILE) * sizeof(float)  /* s_x, s_line [, float4 s_n] */;
  auto kern = lookup_packed_kernel<NV, TILE, OTF, FV, TMA, H0>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) SA_FAIL((int)e, "sa_lookup_packed: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  CUtensorMap m0, m1;
  memset(&m0, 0, sizeof(m0));
  memset(&m1, 0, sizeof(m1));
  if (TMA) {
    int rc = output_map(&m0, a.out[0], B, a.HW, TILE);
    if (rc) return rc;
    if (NV == 2) {
      rc = output_map(&m1, a.out[1], B, a.HW, TILE);
      if (rc) return rc;
    }
  }
  dim3 grid((a.HW + TILE - 1) / TILE, B);
  // output bytes of this launch against the L2: the hint pays up to ~0.6 of the cache (KITTI batch 8: 69 MB of 126)
  a.out_evict_first = (long long)NV * NC * 4 * a.HW * B * 10 <= l2_bytes() * 6;
  {
    static const int env = getenv("SA_B200_LOOKUP_EVICT") ? atoi(getenv("SA_B200_LOOKUP_EVICT")) : -1;
    if (env >= 0) a.out_evict_first = env;
  }
  a.reverse = next_direction(a.packed[0] ? a.packed[0] : a.packed[1]);
  kern<<<grid, NV * TILE, smem, st>>>(a, m0, m1);
  return finish_launch("sa_lookup_packed");
}

// Output path.  Dual lookups: one TMA tensor store per volume with the L2 evict_first hint (factored mono form 20.0 us
// per dual lookup at KITTI size against 21.3 with the transposed 128-bit stores; two packed volumes 21.45 against
// 21.75 - before the hint and the alternating direction the TMA stores lost there, 24.4 against 23.2).  Single
// lookups keep the 128-bit stores (12.3 against 12.6 us).  SA_B200_LOOKUP_TMA=0/1 overrides.
template <int NV, int TILE, int OTF, int FV, int H0>
static int launch_packed_t(const PLookupArgs& a, int B, cudaStream_t st) {
  static const int env = getenv("SA_B200_LOOKUP_TMA") ? atoi(getenv("SA_B200_LOOKUP_TMA")) : -1;
  const bool want = env >= 0 ? env != 0 : (FV >= 0 || (NV == 2 && OTF < 0));
  if (want && (a.HW & 3) == 0 && encode_fn() != nullptr) return launch_packed_tt<NV, TILE, OTF, FV, true, H0>(a, B, st);
  return launch_packed_tt<NV, TILE, OTF, FV, false, H0>(a, B, st);
}

template <int NV, int OTF, int FV = -1, int H0 = 0>
static int launch_packed(PLookupArgs a, int B, cudaStream_t st) {
  // (an L2 prefetch of the lines of the CTA k launches ahead, issued by the warp that idles in the prologue, was
  // measured neutral at every distance - profiles/r2/lookup_l2_prefetch_sweep.txt - and removed)
  // (factored mono volume: 32 - 23.5 us against 24.4 at 64 and 27.3 at 128; its staging is latency-bound)
  static const int tile = getenv("SA_B200_LOOKUP_TILE") ? atoi(getenv("SA_B200_LOOKUP_TILE")) : (FV >= 0 ? 32 : 64);
  // pixels per CTA: 64 measured best at c2 (23.2 us per dual lookup; 128: 24.2, 32: 23.5) - smaller CTAs
  // shorten the tail of the last wave and raise the number of lines in flight per SM
  if (H0 != 0) {   // 16-bit storage: 32 or 64 pixels per CTA only (fewer instantiations)
    if (tile == 32) return launch_packed_t<NV, 32, OTF, FV, H0>(a, B, st);
    return launch_packed_t<NV, 64, OTF, FV, H0>(a, B, st);
  }
  if (tile == 32) return launch_packed_t<NV, 32, OTF, FV, H0>(a, B, st);
  if (tile == 128) return launch_packed_t<NV, 128, OTF, FV, H0>(a, B, st);
  return launch_packed_t<NV, 64, OTF, FV, H0>(a, B, st);
}

}  // namespace sa

extern "C" int64_t sa_packed_row_floats(int W) { return (int64_t)sa::packed_blocks(W) * 32; }

namespace sa {
static int launch_pack(PackArgs& a, bool trunc, bool normals, cudaStream_t st) {
  const int W = a.W;
  const int warps = 8;
  const size_t smem = (size_t)warps * ((W + W / 2 + W / 4 + W / 8 + 3) & ~3) * sizeof(float);
  SA_REQUIRE(smem <= 200 * 1024, SA_E_UNSUPPORTED, "sa_pack_pyramid: W = %d too wide", W);
  void (*kern)(const PackArgs) = normals ? pack_kernel<false, true>
                                         : (trunc ? pack_kernel<true, false> : pack_kernel<false, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) SA_FAIL((int)e, "sa_pack_pyramid: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const long long want = (a.rows + warps - 1) / warps;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  kern<<<grid, warps * 32, smem, st>>>(a);
  return finish_launch("sa_pack_pyramid");
}
}  // namespace sa

extern "C" int sa_pack_pyramid(const float* src, int64_t rows, int W, const float* trunc_disp, const float* trunc_conf,
                               double trunc_gain, int w2_size, float* packed, void* stream) {
  using namespace sa;
  SA_REQUIRE(src && packed && rows > 0, SA_E_INVALID, "sa_pack_pyramid: null pointer / no rows");
  SA_REQUIRE(W >= 8 && W % 8 == 0, SA_E_UNSUPPORTED, "sa_pack_pyramid: W must be a positive multiple of 8 (got %d)", W);
  SA_REQUIRE(aligned16(src) && aligned16(packed), SA_E_ALIGN, "sa_pack_pyramid: pointers must be 16-byte aligned");
  const bool trunc = trunc_disp != nullptr;
  if (trunc)
    SA_REQUIRE(trunc_conf && w2_size > 0 && rows % w2_size == 0, SA_E_INVALID,
               "sa_pack_pyramid: truncation needs conf and rows %% w2_size == 0");
  PackArgs a = {};
  a.src = src; a.packed = packed; a.rows = rows; a.W = W;
  a.disp = trunc_disp; a.conf = trunc_conf;
  a.gain = (float)trunc_gain; a.one_minus_gain = (float)(1.0 - trunc_gain);
  a.w2_size = w2_size;
  return launch_pack(a, trunc, false, (cudaStream_t)stream);
}

extern "C" int sa_pack_pyramid_normals(const float* normals_l, const float* normals_r, float divisor, float post_scale,
                                       int B, int H, int W2, int W3, float* packed, void* stream) {
  using namespace sa;
  SA_REQUIRE(normals_l && normals_r && packed, SA_E_INVALID, "sa_pack_pyramid_normals: null pointer");
  SA_REQUIRE(B > 0 && H > 0 && W2 > 0 && divisor != 0.f, SA_E_INVALID, "sa_pack_pyramid_normals: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_pack_pyramid_normals: W3 must be a positive multiple of 8");
  SA_REQUIRE(aligned16(normals_r) && aligned16(packed), SA_E_ALIGN, "sa_pack_pyramid_normals: pointers must be 16-byte aligned");
  PackArgs a = {};
  a.nl = normals_l; a.nr = normals_r; a.H = H; a.W2 = W2;
  a.divisor = kernel_divisor(divisor); a.inv_divisor = kernel_inv_divisor(divisor); a.post_scale = post_scale;
  a.packed = packed; a.rows = (long long)B * H * W2; a.W = W3;
  return launch_pack(a, false, true, (cudaStream_t)stream);
}

extern "C" int sa_lookup_packed(const float* packed_a, const float* packed_b, int W3, const float* coords,
                                int64_t coords_bstride, float* out_a, float* out_b, int B, int H, int W, void* stream) {
  using namespace sa;
  SA_REQUIRE(packed_a && coords && out_a, SA_E_INVALID, "sa_lookup_packed: null pointer");
  SA_REQUIRE((packed_b == nullptr) == (out_b == nullptr), SA_E_INVALID, "sa_lookup_packed: packed_b / out_b must come together");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 31), SA_E_INVALID, "sa_lookup_packed: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_lookup_packed: W3 must be a multiple of 8");
  SA_REQUIRE((long long)B * H * W * packed_blocks(W3) < (1ll << 29), SA_E_UNSUPPORTED,
             "sa_lookup_packed: packed array of 64 GB or more (32-bit chunk indices)");
  SA_REQUIRE(aligned16(packed_a) && aligned16(out_a) && (!packed_b || (aligned16(packed_b) && aligned16(out_b))), SA_E_ALIGN,
             "sa_lookup_packed: pointers must be 16-byte aligned");
  (void)num_sms();
  PLookupArgs a = {};
  a.packed[0] = packed_a; a.packed[1] = packed_b;
  a.out[0] = out_a; a.out[1] = out_b;
  a.coords = coords; a.coords_bstride = coords_bstride;
  a.HW = H * W; a.W3 = W3; a.nblk = packed_blocks(W3);
  return packed_b ? launch_packed<2, -1>(a, B, (cudaStream_t)stream) : launch_packed<1, -1>(a, B, (cudaStream_t)stream);
}

extern "C" int sa_lookup_packed_normals(const float* packed_a, const float* normals_l, const float* normals_r, float divisor,
                                        float post_scale, int W3, const float* coords, int64_t coords_bstride,
                                        float* out_a, float* out_mono, int B, int H, int W, void* stream) {
  using namespace sa;
  SA_REQUIRE(normals_l && normals_r && coords && out_mono, SA_E_INVALID, "sa_lookup_packed_normals: null pointer");
  SA_REQUIRE((packed_a == nullptr) == (out_a == nullptr), SA_E_INVALID, "sa_lookup_packed_normals: packed_a / out_a must come together");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 31) && divisor != 0.f, SA_E_INVALID,
             "sa_lookup_packed_normals: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_lookup_packed_normals: W3 must be a multiple of 8");
  SA_REQUIRE((long long)B * H * W * packed_blocks(W3) < (1ll << 29), SA_E_UNSUPPORTED,
             "sa_lookup_packed_normals: packed array of 64 GB or more (32-bit chunk indices)");
  SA_REQUIRE(aligned16(normals_r) && aligned16(out_mono) && (!packed_a || (aligned16(packed_a) && aligned16(out_a))), SA_E_ALIGN,
             "sa_lookup_packed_normals: pointers must be 16-byte aligned");
  (void)num_sms();
  PLookupArgs a = {};
  a.coords = coords; a.coords_bstride = coords_bstride;
  a.HW = H * W; a.W3 = W3; a.nblk = packed_blocks(W3);
  a.nl = normals_l; a.nr = normals_r; a.H = H; a.Wimg = W;
  a.divisor = kernel_divisor(divisor); a.inv_divisor = kernel_inv_divisor(divisor); a.post_scale = post_scale;
  if (packed_a) {
    a.packed[0] = packed_a; a.out[0] = out_a; a.out[1] = out_mono;
    return launch_packed<2, 1>(a, B, (cudaStream_t)stream);
  }
  a.out[0] = out_mono;
  return launch_packed<1, 0>(a, B, (cudaStream_t)stream);
}

extern "C" int sa_lookup_packed_factored(const float* packed_a, const float* packed_normals_r, const float* normals_l,
                                         float divisor, float post_scale, int W3, const float* coords,
                                         int64_t coords_bstride, float* out_a, float* out_mono, int B, int H, int W,
                                         void* stream) {
  using namespace sa;
  SA_REQUIRE(packed_normals_r && normals_l && coords && out_mono, SA_E_INVALID, "sa_lookup_packed_factored: null pointer");
  SA_REQUIRE((packed_a == nullptr) == (out_a == nullptr), SA_E_INVALID, "sa_lookup_packed_factored: packed_a / out_a must come together");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 31) && divisor != 0.f, SA_E_INVALID,
             "sa_lookup_packed_factored: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_lookup_packed_factored: W3 must be a multiple of 8");
  SA_REQUIRE((long long)B * H * W * packed_blocks(W3) < (1ll << 29), SA_E_UNSUPPORTED,
             "sa_lookup_packed_factored: packed array of 64 GB or more (32-bit chunk indices)");
  SA_REQUIRE(aligned16(packed_normals_r) && aligned16(out_mono) && (!packed_a || (aligned16(packed_a) && aligned16(out_a))),
             SA_E_ALIGN, "sa_lookup_packed_factored: pointers must be 16-byte aligned");
  (void)num_sms();
  PLookupArgs a = {};
  a.coords = coords; a.coords_bstride = coords_bstride;
  a.HW = H * W; a.W3 = W3; a.nblk = packed_blocks(W3);
  a.nl = normals_l; a.H = H; a.Wimg = W;
  a.divisor = kernel_divisor(divisor); a.inv_divisor = kernel_inv_divisor(divisor); a.post_scale = post_scale;
  if (packed_a) {
    a.packed[0] = packed_a; a.packed[1] = packed_normals_r; a.out[0] = out_a; a.out[1] = out_mono;
    return launch_packed<2, -1, 1>(a, B, (cudaStream_t)stream);
  }
  a.packed[0] = packed_normals_r; a.out[0] = out_mono;
  return launch_packed<1, -1, 0>(a, B, (cudaStream_t)stream);
}

extern "C" int sa_lookup_packed_half(const void* packed_h_a, int half_kind, int mode_b, const float* packed_b,
                                     const float* normals_l, float divisor, float post_scale, int W3, const float* coords,
                                     int64_t coords_bstride, float* out_a, float* out_b, int B, int H, int W, void* stream) {
  using namespace sa;
  SA_REQUIRE(packed_h_a && coords && out_a, SA_E_INVALID, "sa_lookup_packed_half: null pointer");
  SA_REQUIRE(half_kind == 1 || half_kind == 2, SA_E_INVALID, "sa_lookup_packed_half: half_kind must be 1 (fp16) or 2 (bf16)");
  SA_REQUIRE(mode_b >= 0 && mode_b <= 2, SA_E_INVALID, "sa_lookup_packed_half: mode_b must be 0 (none), 1 (packed fp32) or 2 (factored)");
  SA_REQUIRE((mode_b == 0) == (out_b == nullptr) && (mode_b == 0 || packed_b) && (mode_b != 2 || (normals_l && divisor != 0.f)),
             SA_E_INVALID, "sa_lookup_packed_half: second volume arguments do not match mode_b");
  SA_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 31), SA_E_INVALID, "sa_lookup_packed_half: bad sizes");
  SA_REQUIRE(W3 >= 8 && W3 % 8 == 0, SA_E_UNSUPPORTED, "sa_lookup_packed_half: W3 must be a multiple of 8");
  SA_REQUIRE((long long)B * H * W * packed_blocks(W3) < (1ll << 29), SA_E_UNSUPPORTED,
             "sa_lookup_packed_half: packed array of 64 GB or more (32-bit chunk indices)");
  SA_REQUIRE(aligned16(packed_h_a) && aligned16(out_a) && (!packed_b || aligned16(packed_b)) && (!out_b || aligned16(out_b)),
             SA_E_ALIGN, "sa_lookup_packed_half: pointers must be 16-byte aligned");
  (void)num_sms();
  PLookupArgs a = {};
  a.packed[0] = reinterpret_cast<const float*>(packed_h_a); a.packed[1] = packed_b;
  a.out[0] = out_a; a.out[1] = out_b;
  a.coords = coords; a.coords_bstride = coords_bstride;
  a.HW = H * W; a.W3 = W3; a.nblk = packed_blocks(W3);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode_b == 2) {
    a.nl = normals_l; a.H = H; a.Wimg = W;
    a.divisor = kernel_divisor(divisor); a.inv_divisor = kernel_inv_divisor(divisor); a.post_scale = post_scale;
    return half_kind == 1 ? launch_packed<2, -1, 1, 1>(a, B, st) : launch_packed<2, -1, 1, 2>(a, B, st);
  }
  if (mode_b == 1) return half_kind == 1 ? launch_packed<2, -1, -1, 1>(a, B, st) : launch_packed<2, -1, -1, 2>(a, B, st);
  return half_kind == 1 ? launch_packed<1, -1, -1, 1>(a, B, st) : launch_packed<1, -1, -1, 2>(a, B, st);
}
