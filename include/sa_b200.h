/*
 * sa_b200.h - C ABI of the B200-native Stereo Anywhere cost-volume path.
 *
 * One shared library (stereoanywhere_b200/lib/libsa_b200.so), plain pointers and sizes, no
 * torch / C++ types.  Every pointer is a DEVICE pointer unless the parameter name starts with
 * `h_` (host array, read synchronously before the call returns).  `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = legacy default stream).  Launches are asynchronous;
 * nothing here synchronises the device or the host.
 *
 * Return value: 0 on success; > 0 is a `cudaError_t` from the launch; < 0 is one of the
 * SA_E_* argument errors below.  `sa_last_error()` returns a thread-local description.
 * The functions never throw and never fall back to a CPU path.
 *
 * The reference (kei312/stereoanywhere) has no FFI of its own - its "operator API" for this
 * path is a Python class protocol (`models/stereoanywhere/corr.py:75-132`).  Each entry point
 * cites the reference lines it replaces; `INTEGRATION.md` shows the ctypes binding and the
 * three-line patch that selects this block inside `StereoAnywhere.forward`
 * (`models/stereoanywhere/stereoanywhere.py:128-133`).
 *
 * Tensor layouts are the reference's: feature maps NCHW fp32, volumes [B,H,W2,1,W3] fp32
 * (flattened here to rows = B*H*W2 of W3 floats), lookup output [B, L*(2r+1), H, W] fp32.
 */
#ifndef SA_B200_H
#define SA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SA_ABI_VERSION 1
#define SA_MAX_LEVELS 8   /* pyramid levels a lookup can address */
#define SA_MAX_BINS 32    /* depth bins of sa_masked_volume */

#define SA_E_INVALID (-1)     /* bad size / null pointer */
#define SA_E_ALIGN (-2)       /* pointer or pitch not aligned as documented */
#define SA_E_UNSUPPORTED (-3) /* shape outside what the kernel family covers */

int sa_abi_version(void);
const char* sa_last_error(void);

/* ---------------------------------------------------------------- A1 / A2: correlation volume
 * vol[b,h,w2,w3] = (sum_c L[b,c,h,w2] * R[b,c,h,w3]) / divisor * post_scale
 * Replaces `CorrBlock1D.corr` (corr.py:117-132: einsum, / sqrt(C)) and the `1.73 *` of
 * stereoanywhere.py:136 (post_scale).  fmap_l is [B,C,H,W2], fmap_r is [B,C,H,W3], both
 * contiguous fp32; vol is rows = B*H*W2 by W3 floats, contiguous.
 *
 * divisor > 0: correctly rounded division (the reference on the CPU).  divisor < 0: vol = acc * (1.0f / |divisor|) -
 * what the reference computes on a CUDA device, where ATen multiplies by the fp32 reciprocal of a scalar divisor
 * (one ulp apart for sqrt(3), identical for sqrt(256)).  The same convention holds for every `divisor` below.
 *
 * sa_corr_fp32  : SIMT fp32 FMA, any C >= 1 (the mono C=3 volume is a pure streaming write).
 * sa_corr_tf32  : tcgen05 kind::tf32, operands TMA-staged straight from NCHW (MN-major), fp32
 *                 accumulate in TMEM.  Needs C % 8 == 0, W2 % 4 == 0, W3 % 4 == 0, W3 <= 1024,
 *                 16-byte aligned pointers.  Normwise error <= 1e-3 of max|vol|.
 * If pyr1..pyr3 are non-NULL (all three or none) the avg-pooled levels 1..3 of
 * `CorrBlock1D.__init__` (corr.py:88-91) are written by the same kernel (row pitches pitchN in
 * floats, multiples of 4; requires W3 % 8 == 0).
 * If trunc_disp/trunc_conf are non-NULL ([B,1,H,W2] each) the stored volume (and its pyramid)
 * is T * vol with T of `truncate_corr_volume_v2` (utils/utils.py:231-236, stereoanywhere.py:
 * 253-255), trunc_gain = attenuation gain.
 */
int sa_corr_fp32(const float* fmap_l, const float* fmap_r, float* vol, int B, int C, int H, int W2, int W3,
                 float divisor, float post_scale, void* stream);

int sa_corr_tf32(const float* fmap_l, const float* fmap_r, float* vol, int B, int C, int H, int W2, int W3,
                 float divisor, float post_scale, const float* trunc_disp, const float* trunc_conf,
                 double trunc_gain, float* pyr1, float* pyr2, float* pyr3, int64_t pitch1, int64_t pitch2,
                 int64_t pitch3, void* stream);

/* ---------------------------------------------------------------- A3: avg-pooled pyramid
 * dst_k[row, j] = 0.5 * (prev[row, 2j] + prev[row, 2j+1]),  j < floor(W_prev / 2)
 * Replaces the avg_pool2d chain of `CorrBlock1D.__init__` (corr.py:76-91).  Builds n_out (1..3)
 * further levels from `src` (rows x W floats, row pitch src_pitch floats) in ONE pass over
 * `src`.  dst pitches are in floats.  The reference's dead (L+1)-th level is not built.
 * Optional truncation (A5): if trunc_disp != NULL, `masked0` (rows x W, pitch = W) receives
 * T * src and the levels are pooled from the masked values; trunc_disp / trunc_conf hold one
 * float per row (the [B,1,H,W2] maps flattened), `w2_size` = W2 (left-image width).
 * Fast path: W % 8 == 0, pitches % 4 == 0, 16-byte aligned pointers; anything else takes the
 * generic (scalar) kernel - still on the GPU.
 */
int sa_pyramid(const float* src, int64_t rows, int W, int64_t src_pitch, int n_out, float* dst1, float* dst2,
               float* dst3, int64_t pitch1, int64_t pitch2, int64_t pitch3, const float* trunc_disp,
               const float* trunc_conf, double trunc_gain, int w2_size, float* masked0, void* stream);

/* ---------------------------------------------------------------- A4: multi-level lookup
 * For level i < L, tap k in [-r, r]:  xs = (coords[b,0,h,w] + pad0) / 2^i + k,
 *   out[b, i*(2r+1) + k + r, h, w - pad0] = (1-f) * tap(P_i[b,h,w,:], x0) + f * tap(.., x0+1),
 *   x0 = floor(xs), f = xs - x0, tap = 0 outside [0, W_i).
 * Replaces `CorrBlock1D.__call__` + `bilinear_sampler` (corr.py:93-115, utils/utils.py:19-35).
 * h_levels / h_widths / h_pitches are HOST arrays of length num_levels (device pointers,
 * valid widths W_i, row pitches in floats).  coords is [B,2,H,W] (only channel 0 is read;
 * coords_bstride = floats between batches, 2*H*W for a contiguous tensor).  out is
 * [B, L*(2r+1), H, W - pad0 - pad1] contiguous.
 * Fast path (warp-cooperative 128-bit loads): radius <= 5, L <= 4, pad = 0, all pitches % 4 == 0,
 * 16-byte aligned level pointers.  Otherwise a generic kernel runs.
 */
int sa_lookup(const float* const* h_levels, const int* h_widths, const int64_t* h_pitches, int num_levels,
              int radius, const float* coords, int64_t coords_bstride, float* out, int B, int H, int W,
              int pad0, int pad1, void* stream);

/* Same, two volumes (stereo + mono) with one read of coords and one launch
 * (the two calls of stereoanywhere.py:270-271).  Both volumes share L, r and geometry. */
int sa_lookup2(const float* const* h_levels_a, const float* const* h_levels_b, const int* h_widths,
               const int64_t* h_pitches_a, const int64_t* h_pitches_b, int num_levels, int radius,
               const float* coords, int64_t coords_bstride, float* out_a, float* out_b, int B, int H, int W,
               void* stream);

/* ---------------------------------------------------------------- A3 + A4, line-packed (fast path)
 * For num_levels == 4, radius == 4, W3 % 8 == 0 (the model's configuration) the pyramid can be
 * stored "line-packed": per volume row and per block of 8 level-0 columns ONE 128-byte line that
 * holds every value a lookup with floor(x) in that block needs (see csrc/packed.cu).  HBM serves
 * L2 misses in whole 128-byte lines, so a lookup costs 1 line per (pixel, volume) instead of >= 4.
 * Results are bit-identical to sa_pyramid + sa_lookup.
 *   sa_packed_row_floats(W3)  floats per volume row of the packed array ((W3/8 + 9) * 32)
 *   sa_pack_pyramid           src: rows x W3 fp32 -> packed: rows x sa_packed_row_floats(W3);
 *                             optional truncation as in sa_pyramid (the masked level 0 is not
 *                             materialised)
 *   sa_pack_pyramid_normals   the same for the mono volume of A2 (C = 3 unit normals, divisor, post_scale
 *                             as in sa_corr_fp32) computed on the fly: the volume itself is never written
 *   sa_lookup_packed          `CorrBlock1D.__call__` for one (packed_b == out_b == NULL) or two
 *                             volumes; coords / out as in sa_lookup, pad = 0.
 */
int64_t sa_packed_row_floats(int W3);
int sa_pack_pyramid(const float* src, int64_t rows, int W3, const float* trunc_disp, const float* trunc_conf,
                    double trunc_gain, int w2_size, float* packed, void* stream);
int sa_pack_pyramid_normals(const float* normals_l, const float* normals_r, float divisor, float post_scale, int B,
                            int H, int W2, int W3, float* packed, void* stream);
int sa_lookup_packed(const float* packed_a, const float* packed_b, int W3, const float* coords,
                     int64_t coords_bstride, float* out_a, float* out_b, int B, int H, int W, void* stream);

/* Lookup with the MONO volume computed on the fly from the C = 3 normal maps (stereoanywhere.py:136 when the
 * volume is looked up as it is): the thread of a (pixel, mono) pair forms the 80 level-0 values its packed line
 * is made of and pools them - bit-identical to sa_pack_pyramid_normals + sa_lookup_packed, with no packed mono
 * array at all.  packed_a / out_a: an ordinary packed volume looked up in the same launch (both NULL: mono only).
 * normals_l is [B,3,H,W], normals_r [B,3,H,W3]; divisor / post_scale as in sa_corr_fp32. */
int sa_lookup_packed_normals(const float* packed_a, const float* normals_l, const float* normals_r, float divisor,
                             float post_scale, int W3, const float* coords, int64_t coords_bstride, float* out_a,
                             float* out_mono, int B, int H, int W, void* stream);

/* Lookup with the MONO volume in FACTORED form.  The volume of stereoanywhere.py:136 has rank 3 (V = g * nL^T nR /
 * sqrt 3) and both the avg-pool pyramid (corr.py:88-91) and the packed layout are linear in it, so the packed line
 * of pixel (b,h,w2) is the combination, with nL[b,:,h,w2] as coefficients, of the packed lines of the three
 * right-normal rows.  packed_normals_r = sa_pack_pyramid(normals_r viewed as B*3*H rows of W3) - 14 MB at KITTI
 * size, L2-resident, written in ~10 us - replaces the 1.47 GB packed mono volume.  Differences from
 * sa_pack_pyramid_normals + sa_lookup_packed are fp32 rounding only (post_scale / divisor is applied to the
 * coefficients; stored border entries of levels 1..3 are pooled before the contraction): |delta| <= 1e-6 for unit
 * normals.  packed_a / out_a: an
 * ordinary packed volume looked up in the same launch (both NULL: mono only).  normals_l is [B,3,H,W]. */
int sa_lookup_packed_factored(const float* packed_a, const float* packed_normals_r, const float* normals_l,
                              float divisor, float post_scale, int W3, const float* coords, int64_t coords_bstride,
                              float* out_a, float* out_mono, int B, int H, int W, void* stream);

/* ---------------------------------------------------------------- A1 + A5 + A3 fused (tensor cores -> packed pyramid)
 * packed = sa_pack_pyramid(T * sa_corr_tf32(L, R)) in ONE kernel: the truncation product and the avg-pooled
 * pyramid are formed in the GEMM epilogue (TMEM -> registers -> packed lines -> TMA store); the fp32 volume
 * is never written.  Replaces corr.py:117-132 + utils/utils.py:216-238 / stereoanywhere.py:253-255 +
 * corr.py:76-91 for the stereo block.  Bit-identical to the two-step path.  trunc_disp / trunc_conf may both
 * be NULL (no truncation).  Needs C % 32 == 0, W2 % 4 == 0, W3 % 8 == 0, 16-byte aligned pointers;
 * packed is rows = B*H*W2 by sa_packed_row_floats(W3) floats. */
int sa_corr_pack_tf32(const float* fmap_l, const float* fmap_r, int B, int C, int H, int W2, int W3, float divisor,
                      float post_scale, const float* trunc_disp, const float* trunc_conf, double trunc_gain,
                      float* packed, void* stream);
/* The same kernel with the packed pyramid stored in 16 bits (half_kind 1 = fp16, 2 = bf16; round to nearest even):
 * lines of 64 bytes, element i of a line in the low (i even) / high (i odd) half of 32-bit word i / 2; packed_h is
 * rows x (W3/8 + 9) x 64 bytes.  The arithmetic up to the rounding of the stored values is that of
 * sa_corr_pack_tf32; read it with sa_lookup_packed_half.  Opt-in storage mode: fp16 adds at most 2^-11 = 4.9e-4 of
 * max|vol| (inside the TF32 tolerance of 1e-3 together with the product's own ~3e-4), bf16 3.9e-3 (the 1e-2 class). */
int sa_corr_pack_tf32_half(const float* fmap_l, const float* fmap_r, int B, int C, int H, int W2, int W3, float divisor,
                           float post_scale, const float* trunc_disp, const float* trunc_conf, double trunc_gain,
                           int half_kind, void* packed_h, void* stream);

/* ---------------------------------------------------------------- lookup + motion-encoder front end
 * out_v[b,n,h,w] = relu(bias[n] + sum_k weight[n,k] * lookup_v[b,k,h,w]) for the stereo and the mono volume
 * with the SAME 1x1 convolution (models/stereoanywhere/update.py:74,80-84: `relu(convc1(corr))`,
 * `relu(convc1(corr_mono))`, convc1 = Conv2d(36, 64, 1)); lookups as in sa_lookup_packed
 * (stereoanywhere.py:270-271).  weight is [64][36] (the conv weight with its 1x1 dims dropped), bias [64];
 * out_a / out_b are [B,64,H,W].  The 36-channel lookups are never written to memory: the taps feed a TF32
 * tcgen05 MMA out of shared memory.  Normwise error vs the fp32 convolution <= 1e-3. */
int sa_lookup_packed_conv(const float* packed_a, const float* packed_b, int W3, const float* coords,
                          int64_t coords_bstride, const float* weight, const float* bias, float* out_a, float* out_b,
                          int B, int H, int W, void* stream);
/* The same with the mono volume in factored form (packed_normals_r / normals_l / divisor / post_scale as in
 * sa_lookup_packed_factored): the mono line of a pixel is combined from three right-normal lines while staging. */
int sa_lookup_factored_conv(const float* packed_a, const float* packed_normals_r, const float* normals_l,
                            float divisor, float post_scale, int W3, const float* coords, int64_t coords_bstride,
                            const float* weight, const float* bias, float* out_a, float* out_mono, int B, int H, int W,
                            void* stream);

/* ---------------------------------------------------------------- soft-argmax / entropy reductions
 * The four reductions the model runs over the aggregated mono volume (stereoanywhere.py:174-177), two per
 * launch with ONE read of the volume (the reference makes four separate softmax passes).  vol is
 * [BH = B*H][W2][W3] fp32 contiguous (the [B,1,H,W2,W3] tensor).
 *   sa_volume_softargmax:   disp_left [BH,W2] = w2 - sum_w3 softmax_w3(vol) * w3   (utils/utils.py:112-131)
 *                           disp_right[BH,W3] = sum_w2 softmax_w2(vol) * w2 - w3   (utils/utils.py:133-152)
 *   sa_volume_entropy_conf: conf_left [BH,W2] = 1 + sum_w3 p*log2(p+1e-6)/log2(W3) (utils/utils.py:154-161)
 *                           conf_right[BH,W3] = 1 + sum_w2 p*log2(p+1e-6)/log2(W2) (utils/utils.py:163-170)
 * W3 <= 1024.  Agreement with the reference's fp32 ATen sequence: <= 2e-4 px / <= 2e-5 (tests). */
int sa_volume_softargmax(const float* vol, int64_t BH, int W2, int W3, float* disp_left, float* disp_right,
                         void* stream);
int sa_volume_entropy_conf(const float* vol, int64_t BH, int W2, int W3, float* conf_left, float* conf_right,
                           void* stream);

/* ---------------------------------------------------------------- backward of lookup and pyramid
 * Adjoints of sa_lookup / sa_pyramid for the reference's training step (train.py:277,383: autograd through
 * grid_sample and avg_pool2d).  Coordinates are detached before every lookup (stereoanywhere.py:268), so only
 * the volume receives a gradient.
 *   sa_lookup_backward   h_dlevels[i][row, c] += adjoint of the taps of level i for grad_out [B, L*(2r+1), H, W]
 *                        (same coords / pad0 as the forward call; level-gradient buffers are rows x pitch_i
 *                        floats, zero-initialised by the caller, accumulated over the GRU iterations)
 *   sa_pyramid_backward  folds the level gradients into level 0 in place: d0 <- [T *] (dP_0 + pooled adjoints);
 *                        with trunc_* the result is the gradient w.r.t. V of the block built from T * V
 *                        (T detached as in the reference, stereoanywhere.py:203). */
/*   sa_corr_backward_tf32  adjoint of A1 on the tensor cores (corr.py:130-132 under autograd): with G = grad_vol
 *                        [B,H,W2,W3] and s = post_scale / divisor,
 *                          grad_l[b,c,h,w2] = s * sum_w3 G[b,h,w2,w3] * fmap_r[b,c,h,w3]
 *                          grad_r[b,c,h,w3] = s * sum_w2 G[b,h,w2,w3] * fmap_l[b,c,h,w2]
 *                        (either output may be NULL).  tcgen05 kind::tf32, operands TMA-staged from NCHW (K-major)
 *                        and from the volume gradient, fp32 accumulate in TMEM; W2 % 4 == 0, W3 % 4 == 0, 16-byte
 *                        aligned pointers; normwise error <= 1e-3 of max|grad|. */
int sa_corr_backward_tf32(const float* grad_vol, const float* fmap_l, const float* fmap_r, float* grad_l, float* grad_r,
                          int B, int C, int H, int W2, int W3, float divisor, float post_scale, void* stream);
int sa_lookup_backward(const float* grad_out, const float* coords, int64_t coords_bstride, float* const* h_dlevels,
                       const int* h_widths, const int64_t* h_pitches, int num_levels, int radius, int B, int H, int W,
                       int pad0, void* stream);
int sa_pyramid_backward(float* d0, const float* const* h_dlevels, const int* h_widths, const int64_t* h_pitches,
                        int num_levels, int64_t rows, const float* trunc_disp, const float* trunc_conf, double trunc_gain,
                        int w2_size, void* stream);

/* ---------------------------------------------------------------- lookups from a 16-bit packed pyramid
 * The lookup of sa_lookup_packed with volume A stored by sa_corr_pack_tf32_half (half_kind 1 = fp16, 2 = bf16): the
 * stored values are widened to fp32 and everything after that is sa_lookup_packed's arithmetic - the result equals,
 * bit for bit, a lookup from the fp32 packed array holding the same (rounded) values.  mode_b selects the second
 * volume of a dual lookup: 0 none (out_b NULL), 1 an fp32 packed pyramid (packed_b), 2 the factored mono volume
 * (packed_b = packed right normals, normals_l / divisor / post_scale as in sa_lookup_packed_factored). */
int sa_lookup_packed_half(const void* packed_h_a, int half_kind, int mode_b, const float* packed_b, const float* normals_l,
                          float divisor, float post_scale, int W3, const float* coords, int64_t coords_bstride,
                          float* out_a, float* out_b, int B, int H, int W, void* stream);

/* ---------------------------------------------------------------- config 4: tile stitch (no collective)
 * Replaces the accumulate / normalise of `TileWrapper` (mapreduce_v2/tile_wrapper.py:172-185, :206, :226-247,
 * :340-362) for tile-sharded inference on one node.
 *   sa_stitch_tile    slot[y, x] = src[(y + pad_top) / up, (x + pad_left) / up] * scale * weight[y, x] * mult for the
 *                     un-padded th x tw tile (tw % 4 == 0).  `src` is the tile's result, src_h x src_w floats: the
 *                     padded full-resolution model output (up = 1, scale = -1: the reference negates it) or a
 *                     quarter-resolution disparity (up = 4, scale = 4).  `weight` is the cosine blend window
 *                     (tile_wrapper.py:36-49), `mult` the number of times the reference emits this tile.  `slot`
 *                     (th * tw floats, 16-byte aligned) may live on ANOTHER GPU of the node (NVLink peer pointer,
 *                     e.g. a torch symmetric-memory buffer): the stores are the gather.
 *   sa_stitch_finish  out[img, y, x] = (sum of the slots of the tiles covering (img, y, x), in table order) /
 *                     den[y, x].  tile_table: n_tiles x 6 DEVICE ints {image, y0, y1, x0, x1, slot offset in floats
 *                     / 4}, x0 and x1 multiples of 4, listed in the reference's enumeration order; den: [H, W],
 *                     already clamped (tile_wrapper.py:185); out: [images, H, W].  Table order fixes the summation
 *                     order: the result is independent of which GPU produced which tile. */
int sa_stitch_tile(const float* src, int src_h, int src_w, int up, float scale, int pad_top, int pad_left, int th, int tw,
                   const float* weight, float mult, float* slot, void* stream);
int sa_stitch_finish(const float* slots, const int* tile_table, int n_tiles, const float* den, float* out, int images,
                     int H, int W, void* stream);

/* ---------------------------------------------------------------- A5: truncation mask (standalone)
 * mask[b,h,w2,w3] = (1-c) + c * (sigmoid((w2 - d) - w3) * (1-g) + g); writes `out` = mask * vol
 * when vol != NULL, else the mask itself.  Replaces `truncate_corr_volume_v2`
 * (utils/utils.py:216-238) and the product of stereoanywhere.py:253-255. */
int sa_truncate(const float* vol, const float* disp, const float* conf, double gain, float* out, int64_t rows,
                int W2, int W3, void* stream);

/* ---------------------------------------------------------------- A6: depth-bin masked volume
 * out[b,n,h,w2,w3] = vol[b,h,w2,w3] if bin(mde_l[b,h,w2]) == bin(mde_r[b,h,w3]) == n else 0,
 * bin(x) = n iff h_edges[n] <= x < h_edges[n+1]  (h_edges: HOST array of n_bins+1 floats;
 * the reference's edges are float32(i / N)).  Replaces `generate_masks` + the two broadcast
 * products (utils/utils.py:48-54, stereoanywhere.py:138-139,161).
 * If vol == NULL the volume is computed on the fly from unit normals normals_l / normals_r
 * ([B,3,H,W]) as in A2 (divisor, post_scale). */
int sa_masked_volume(const float* vol, const float* normals_l, const float* normals_r, float divisor,
                     float post_scale, const float* mde_l, const float* mde_r, const float* h_edges, int n_bins,
                     float* out, int B, int H, int W2, int W3, void* stream);

/* ---------------------------------------------------------------- producers of the path's inputs
 *   sa_mono_inputs   mde [B,1,H,W] -> lowres [B,1,Hl,Wl] (Hl = floor(H / 2^n), bilinear, align_corners=True:
 *                    stereoanywhere.py:109-110), normals [B,3,Hl,Wl] = normalise(-d/dx (g d), -d/dy (g d), 1) with the
 *                    replicate-padded central difference of kornia's spatial_gradient(mode="diff")
 *                    (utils/utils.py:73-77; g = normal_gain) and, when masks_f16 != NULL, the one-hot depth bins
 *                    [B,n_bins,Hl,Wl] as fp16 (utils/utils.py:48-54; h_edges as in sa_masked_volume) - one launch.
 *   sa_weighted_lsq  scale[b], shift[b] of weighted_lsq (utils/utils.py:345-384) for B samples of n values each
 *                    (mono / disp / conf: [B, n] fp32): quantile window [min_q, max_q] of relu(disp) by exact radix
 *                    select (torch.quantile's linear interpolation), weights 0.9 |conf| + 0.1, normal equations in
 *                    double.  One CTA per sample, no host sync. */
int sa_mono_inputs(const float* mde, int B, int H, int W, int n_downsample, float normal_gain, const float* h_edges,
                   int n_bins, float* lowres, float* normals, void* masks_f16, void* stream);
int sa_weighted_lsq(const float* mono, const float* disp, const float* conf, int B, int n, float min_quantile,
                    float max_quantile, float* scale, float* shift, void* stream);

/* ---------------------------------------------------------------- A7: training-only corruption
 * mode 0 (roll):  out = vol*(1-m) + roll(vol, shift, dim=W2)*m
 * mode 1 (noise): out = vol*(1-m) + vol*noise[b,h,w2]*m
 * mode 2 (gauss): out = vol*(1-m) + vol*(k*exp(-(w2-w3)^2/2))*m
 * m = bin_mask[b,h,w2] in {0,1}.  Replaces stereoanywhere.py:214-251 (+ utils/utils.py:200-214). */
int sa_corrupt(const float* vol, const float* bin_mask, int mode, int shift, const float* noise, float gauss_k,
               float* out, int B, int H, int W2, int W3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SA_B200_H */
