// Tile stitch of the MapReduce full-resolution inference (BASELINE config 4), reference:
// mapreduce_v2/tile_wrapper.py:172-185 (loop + normalise), :206 (negate), :226-247 (pad / un-pad),
// :340-362 (`stitched[..] += disp * weight`, `weight_sum[..] += weight`).
//
// Two kernels, no collective:
//   sa_stitch_tile    one launch per tile, right after the tile's model run: un-pad, (optionally) nearest-upsample,
//                     scale (the reference negates the model output), multiply by the cosine blend window and the
//                     tile's multiplicity, and store the weighted tile into its own SLOT.  The slot pointer may be a
//                     peer pointer of the gathering GPU (NVLink / NVSwitch, e.g. a torch symmetric-memory buffer):
//                     the "gather" of a tile-sharded run is these plain coalesced 128-bit peer stores - every tile
//                     leaves its GPU exactly once, as soon as it is finished.
//   sa_stitch_finish  one launch per step on the gathering GPU: out[img, y, x] = (sum over the tiles covering the
//                     pixel, in the reference's enumeration order, of their slot values) / den[y, x].
// The sum order is fixed by the tile table, so the stitched image does not depend on how the tiles were sharded:
// an N-GPU run is bit-identical to the one-GPU run.
#include "sa_common.cuh"

namespace sa {

struct StitchTileArgs {
  const float* src;     // [src_h, src_w] tile result (quarter resolution when up == 4, else padded full resolution)
  int src_w;
  int up;               // nearest-neighbour upsampling factor of src (1, 2 or 4 ...)
  float scale;          // -1 for a model output (tile_wrapper.py:206), +up for a quarter-resolution disparity
  int pad_top, pad_left;
  int th, tw;           // un-padded tile size
  const float* weight;  // [th, tw] blend window
  float mult;           // times the reference emits this tile
  float* slot;          // [th, tw] destination (local or peer)
};

__global__ void __launch_bounds__(256) stitch_tile_kernel(const StitchTileArgs a) {
  const int tw4 = a.tw >> 2;
  const long long n4 = (long long)a.th * tw4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / tw4), x = (int)(i - (long long)y * tw4) * 4;
    const float* srow = a.src + (long long)((y + a.pad_top) / a.up) * a.src_w;
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.weight + (long long)y * a.tw + x));
    float4 o;
    o.x = (__ldg(srow + (x + a.pad_left) / a.up) * a.scale) * w.x * a.mult;
    o.y = (__ldg(srow + (x + 1 + a.pad_left) / a.up) * a.scale) * w.y * a.mult;
    o.z = (__ldg(srow + (x + 2 + a.pad_left) / a.up) * a.scale) * w.z * a.mult;
    o.w = (__ldg(srow + (x + 3 + a.pad_left) / a.up) * a.scale) * w.w * a.mult;
    *reinterpret_cast<float4*>(a.slot + (long long)y * a.tw + x) = o;
  }
}

// tile table entry: {image, y0, y1, x0, x1, slot offset / 4 (in float4 units would overflow nothing; floats / 4)}
constexpr int kTileFields = 6;

__global__ void __launch_bounds__(256) stitch_finish_kernel(const float* __restrict__ slots, const int* __restrict__ table,
                                                            int n_tiles, const float* __restrict__ den,
                                                            float* __restrict__ out, int images, int H, int W) {
  extern __shared__ int s_tab[];
  for (int i = threadIdx.x; i < n_tiles * kTileFields; i += blockDim.x) s_tab[i] = table[i];
  __syncthreads();
  const int W4 = W >> 2;
  const long long n4 = (long long)images * H * W4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W4) * 4;
    const long long r = i / W4;
    const int y = (int)(r % H), img = (int)(r / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < n_tiles; ++t) {
      const int* e = s_tab + t * kTileFields;
      if (e[0] != img || y < e[1] || y >= e[2] || x < e[3] || x >= e[4]) continue;   // x0, x1 are multiples of 4
      const float4 v = *reinterpret_cast<const float4*>(slots + (long long)e[5] * 4 + (long long)(y - e[1]) * (e[4] - e[3]) + (x - e[3]));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float4 d = __ldg(reinterpret_cast<const float4*>(den + (long long)y * W + x));
    acc.x /= d.x; acc.y /= d.y; acc.z /= d.z; acc.w /= d.w;
    st_stream_v4(out + ((long long)img * H + y) * W + x, acc);
  }
}

}  // namespace sa

extern "C" int sa_stitch_tile(const float* src, int src_h, int src_w, int up, float scale, int pad_top, int pad_left,
                              int th, int tw, const float* weight, float mult, float* slot, void* stream) {
  using namespace sa;
  SA_REQUIRE(src && weight && slot && th > 0 && tw > 0 && up >= 1, SA_E_INVALID, "sa_stitch_tile: null pointer / bad sizes");
  SA_REQUIRE(tw % 4 == 0, SA_E_UNSUPPORTED, "sa_stitch_tile: tile width must be a multiple of 4 (got %d)", tw);
  SA_REQUIRE(pad_top >= 0 && pad_left >= 0 && (th + pad_top + up - 1) / up <= src_h && (tw + pad_left + up - 1) / up <= src_w,
             SA_E_INVALID, "sa_stitch_tile: the un-padded tile does not fit the source");
  SA_REQUIRE(aligned16(weight) && aligned16(slot), SA_E_ALIGN, "sa_stitch_tile: weight / slot must be 16-byte aligned");
  StitchTileArgs a = {src, src_w, up, scale, pad_top, pad_left, th, tw, weight, mult, slot};
  const long long n4 = (long long)th * (tw / 4);
  const long long want = (n4 + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  stitch_tile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return finish_launch("sa_stitch_tile");
}

extern "C" int sa_stitch_finish(const float* slots, const int* tile_table, int n_tiles, const float* den, float* out,
                                int images, int H, int W, void* stream) {
  using namespace sa;
  SA_REQUIRE(slots && tile_table && den && out && n_tiles > 0 && images > 0 && H > 0 && W > 0, SA_E_INVALID,
             "sa_stitch_finish: null pointer / bad sizes");
  SA_REQUIRE(W % 4 == 0 && n_tiles <= 2048, SA_E_UNSUPPORTED, "sa_stitch_finish: W %% 4 == 0 and <= 2048 tiles required");
  SA_REQUIRE(aligned16(slots) && aligned16(den) && aligned16(out), SA_E_ALIGN, "sa_stitch_finish: pointers must be 16-byte aligned");
  const long long n4 = (long long)images * H * (W / 4);
  const long long want = (n4 + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  const size_t smem = (size_t)n_tiles * kTileFields * sizeof(int);
  stitch_finish_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(slots, tile_table, n_tiles, den, out, images, H, W);
  return finish_launch("sa_stitch_finish");
}
