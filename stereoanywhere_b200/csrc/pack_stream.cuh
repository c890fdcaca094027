// Streaming line packer shared by the kernels that write the line-packed pyramid (csrc/packed.cu layout) with
// one thread per volume row: the row is consumed left to right in chunks of 32 level-0 values and every step
// emits four complete 128-byte lines from a register window - all indices compile-time constants.
//   step cc (chunk = columns 32cc .. 32cc+31) emits the lines q = 4cc-5 .. 4cc-2, i.e. line index 4cc + j:
//     slots  0..16  L0[8q-4 .. 8q+12]
//     slots 17..21  L1[4q-4], L1[4q-3], L1[4q+6], L1[4q+7], L1[4q+8]
//     slots 22..26  L2[2q-4], L2[2q-3], L2[2q+4], L2[2q+5], L2[2q+6]
//     slots 27..31  L3[ q-4], L3[ q-3], L3[ q+3], L3[ q+4], L3[ q+5]
//   window (relative to the chunk of the current step):
//     P [i] = L0[32cc - 44 + i]  i < 44
//     H1[i] = L1[16cc - 24 + i]  i < 40 (24.. are this step's)
//     H2[i] = L2[ 8cc - 14 + i]  i < 22 (14..)
//     H3[i] = L3[ 4cc -  9 + i]  i < 13 ( 9..)
// Columns left of 0 / right of W3 enter as zeros, which is the layout's zero padding.  Pooling is the
// pyramid's own 0.5 * (a + b) (reference corr.py:88-91), so the lines equal those of pack_kernel bit for bit.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sa_common.cuh"

namespace sa {

// two floats -> one word of 16-bit storage (lo in bits 0..15), and back
template <int HK>
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  if (HK == 1) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <int HK>
__device__ __forceinline__ void unpack_half2(uint32_t w, float& lo, float& hi) {
  if (HK == 1) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x; hi = f.y;
  } else {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
  }
}

struct PackWindow {
  float P[44], H1[40], H2[22], H3[13];

  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 44; ++i) P[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) H1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 14; ++i) H2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) H3[i] = 0.f;
  }
  // pooled values of the chunk v = L0[32cc .. 32cc+31]
  __device__ __forceinline__ void pool(const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) H1[24 + i] = (v[2 * i] + v[2 * i + 1]) * 0.5f;
#pragma unroll
    for (int i = 0; i < 8; ++i) H2[14 + i] = (H1[24 + 2 * i] + H1[24 + 2 * i + 1]) * 0.5f;
#pragma unroll
    for (int i = 0; i < 4; ++i) H3[9 + i] = (H2[14 + 2 * i] + H2[14 + 2 * i + 1]) * 0.5f;
  }
  // line j (0..3) of the current step as 8 float4, written 128B-swizzled for a TMA store: chunk k lands at
  // chunk k ^ swz of the 128-byte row `dst_row`
  template <int J>
  __device__ __forceinline__ void store_line(uint8_t* dst_row, int swz) const {
    float ln[32];
#pragma unroll
    for (int s = 0; s < 17; ++s) ln[s] = P[8 * J + s];
    ln[17] = H1[4 * J]; ln[18] = H1[4 * J + 1]; ln[19] = H1[4 * J + 10]; ln[20] = H1[4 * J + 11]; ln[21] = H1[4 * J + 12];
    ln[22] = H2[2 * J]; ln[23] = H2[2 * J + 1]; ln[24] = H2[2 * J + 8]; ln[25] = H2[2 * J + 9]; ln[26] = H2[2 * J + 10];
    ln[27] = H3[J]; ln[28] = H3[J + 1]; ln[29] = H3[J + 7]; ln[30] = H3[J + 8]; ln[31] = H3[J + 9];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      *reinterpret_cast<float4*>(dst_row + ((k ^ swz) << 4)) = make_float4(ln[4 * k], ln[4 * k + 1], ln[4 * k + 2], ln[4 * k + 3]);
  }
  // the same line in 16-bit storage (HK = 1: fp16, 2: bf16; round to nearest even): 64 bytes, element i in the low /
  // high half of word i >> 1, written 64B-swizzled for a TMA store (chunk k of row r lands at chunk k ^ ((r >> 1) & 3))
  template <int J, int HK>
  __device__ __forceinline__ void store_line_half(uint8_t* dst_row, int row) const {
    float ln[32];
#pragma unroll
    for (int s = 0; s < 17; ++s) ln[s] = P[8 * J + s];
    ln[17] = H1[4 * J]; ln[18] = H1[4 * J + 1]; ln[19] = H1[4 * J + 10]; ln[20] = H1[4 * J + 11]; ln[21] = H1[4 * J + 12];
    ln[22] = H2[2 * J]; ln[23] = H2[2 * J + 1]; ln[24] = H2[2 * J + 8]; ln[25] = H2[2 * J + 9]; ln[26] = H2[2 * J + 10];
    ln[27] = H3[J]; ln[28] = H3[J + 1]; ln[29] = H3[J + 7]; ln[30] = H3[J + 8]; ln[31] = H3[J + 9];
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = pack_half2<HK>(ln[2 * i], ln[2 * i + 1]);
    const int sw = (row >> 1) & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<uint4*>(dst_row + ((k ^ sw) << 4)) = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
  }
  // slide the window by one chunk
  __device__ __forceinline__ void advance(const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) P[i] = P[32 + i];
#pragma unroll
    for (int i = 0; i < 32; ++i) P[12 + i] = v[i];
#pragma unroll
    for (int i = 0; i < 24; ++i) H1[i] = H1[16 + i];
#pragma unroll
    for (int i = 0; i < 14; ++i) H2[i] = H2[8 + i];
#pragma unroll
    for (int i = 0; i < 9; ++i) H3[i] = H3[4 + i];
  }
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

}  // namespace sa
