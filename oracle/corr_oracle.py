"""CPU oracle for the Stereo Anywhere cost-volume hot path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as
the checker / the timed CPU baseline.  `stereoanywhere_b200` never imports it.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against
fixtures under tests/golden/ that were produced by running the *unmodified* reference
(`/root/reference`, imported through oracle/ref_shim.py) on seeded inputs; the generator is
tests/golden/make_golden.py.  The only unpinned boundary is kornia's `spatial_gradient`
(absent from this image, unpinned in the reference's requirements.txt:3) which produces the
*inputs* of the mono correlation and is outside the path.

Two restatements are kept side by side:

* ``aten_*``   - the reference's own sequence of library calls (einsum, avg_pool2d,
  grid_sample ...) restated as free functions.  This is what the reference executes on a CPU,
  so it is also the timed "reference CPU path" (kind = "port").
* ``closed_*`` - the same mathematics written as explicit float64 numpy index arithmetic
  (SURVEY.md Appendix A).  Independent of ATen; used to bound the ATen path's own noise.

All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# A1 / A2 - all-pairs 1-D correlation volume
# --------------------------------------------------------------------------------------


def aten_corr_volume(fmap_l: torch.Tensor, fmap_r: torch.Tensor) -> torch.Tensor:
    """V[b,h,w2,0,w3] = sum_c L[b,c,h,w2] R[b,c,h,w3] / sqrt(C).

    Follows models/stereoanywhere/corr.py:117-132 (`CorrBlock1D.corr`): einsum over the
    channel axis, reshape to [B,H,W2,1,W3], divide by `torch.sqrt(torch.tensor(C))` (a CPU
    float32 scalar) and cast back to the input dtype.
    """
    b, c, h, w2 = fmap_l.shape
    w3 = fmap_r.shape[3]
    vol = torch.einsum("bchw,bchv->bhwv", fmap_l, fmap_r)
    vol = vol.reshape(b, h, w2, 1, w3).contiguous()
    denom = torch.sqrt(torch.tensor(c))
    return (vol / denom).to(fmap_l.dtype)


def aten_mono_corr_volume(normals_l: torch.Tensor, normals_r: torch.Tensor) -> torch.Tensor:
    """Mono volume as the model builds it: 1.73 * corr(nL, nR).

    models/stereoanywhere/stereoanywhere.py:136 (python float 1.73 times the C=3 volume).
    """
    return 1.73 * aten_corr_volume(normals_l, normals_r)


def closed_corr_volume(fmap_l: np.ndarray, fmap_r: np.ndarray, post_scale: float = 1.0) -> np.ndarray:
    """float64 closed form of A1/A2; returns [B,H,W2,W3]."""
    l64 = fmap_l.astype(np.float64)
    r64 = fmap_r.astype(np.float64)
    c = l64.shape[1]
    vol = np.einsum("bchw,bchv->bhwv", l64, r64)
    # the reference divides by float32(sqrt(C)) (corr.py:132)
    return vol / float(np.float32(math.sqrt(c))) * post_scale


# --------------------------------------------------------------------------------------
# A3 - average-pooled pyramid
# --------------------------------------------------------------------------------------


def aten_pyramid(fullcorr: torch.Tensor, num_levels: int = 4) -> List[torch.Tensor]:
    """Pyramid of `CorrBlock1D.__init__` (corr.py:76-91).

    fullcorr is [B,H,W2,1,W3]; returns num_levels+1 tensors [B*H*W2,1,1,W3_i] - the reference
    builds one level more than it reads (corr.py:88-91 vs :101); the dead level is kept here so
    the timed CPU baseline does the reference's work.
    """
    b, h, w2, d, w3 = fullcorr.shape
    lvl = fullcorr.reshape(b * h * w2, d, 1, w3)
    out = [lvl]
    for _ in range(num_levels):
        lvl = F.avg_pool2d(lvl, [1, 2], stride=[1, 2])
        out.append(lvl)
    return out


def closed_pyramid(vol: np.ndarray, num_levels: int = 4) -> List[np.ndarray]:
    """Closed form: P_{i+1}[..., j] = 0.5 (P_i[..., 2j] + P_i[..., 2j+1]), j < floor(W_i/2).

    `vol` is [..., W3]; returns num_levels arrays (the dead level is not produced).  Computed
    in the dtype of `vol` so that a float32 input reproduces the reference bit for bit
    (halving is exact in binary floating point).
    """
    out = [vol]
    cur = vol
    for _ in range(num_levels - 1):
        wi = cur.shape[-1] // 2
        cur = (cur[..., 0 : 2 * wi : 2] + cur[..., 1 : 2 * wi : 2]) * cur.dtype.type(0.5)
        out.append(cur)
    return out


# --------------------------------------------------------------------------------------
# A4 - multi-level radius-r linear lookup
# --------------------------------------------------------------------------------------


def _aten_sample_row(img: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """`bilinear_sampler` (utils/utils.py:19-35) for an H==1 image.

    Pixel x -> normalised 2x/(W-1)-1, y passed through, grid_sample(align_corners=True,
    zeros padding).  The reference's `torch.unique(ygrid)` assert is a host sync with no
    numerical effect and is not restated.
    """
    w = img.shape[-1]
    xg, yg = coords.split([1, 1], dim=-1)
    xg = 2 * xg / (w - 1) - 1
    grid = torch.cat([xg, yg], dim=-1)
    return F.grid_sample(img.float(), grid, align_corners=True).to(img.dtype)


def aten_lookup(
    pyramid: Sequence[torch.Tensor],
    coords: torch.Tensor,
    radius: int = 4,
    num_levels: int = 4,
    pad: Sequence[int] = (0, 0),
) -> torch.Tensor:
    """`CorrBlock1D.__call__` (corr.py:93-115): [B,2,H,W] coords -> [B, L*(2r+1), H, W']."""
    r = radius
    cx = coords[:, :1].permute(0, 2, 3, 1) + pad[0]
    b, h, w, _ = cx.shape
    taps = torch.linspace(-r, r, 2 * r + 1).view(1, 1, 2 * r + 1, 1).to(cx.device)
    per_level = []
    for i in range(num_levels):
        x = taps + cx.reshape(b * h * w, 1, 1, 1) / 2**i
        xy = torch.cat([x, torch.zeros_like(x)], dim=-1)
        s = _aten_sample_row(pyramid[i], xy).view(b, h, w, -1)
        per_level.append(s[:, :, pad[0] : w - pad[1], :])
    out = torch.cat(per_level, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().to(coords.dtype)


def closed_lookup(
    levels: Sequence[np.ndarray],
    coords_x: np.ndarray,
    radius: int = 4,
    pad: Sequence[int] = (0, 0),
) -> np.ndarray:
    """float64 closed form of A4.

    levels[i] is [B,H,W,W_i]; coords_x is [B,H,W] (channel 0 of the model's coords).
    out[b, i*(2r+1)+(k+r), h, w] = (1-f) tap(x0) + f tap(x0+1), x0 = floor(xs), f = xs - x0,
    xs = (x + pad0)/2^i + k, tap() = 0 outside [0, W_i).
    """
    b, h, w = coords_x.shape
    r = radius
    nl = len(levels)
    out = np.zeros((b, nl * (2 * r + 1), h, w), dtype=np.float64)
    x = coords_x.astype(np.float64) + pad[0]
    for i, lv in enumerate(levels):
        lv64 = lv.astype(np.float64)
        wi = lv64.shape[-1]
        for k in range(-r, r + 1):
            xs = x / (2.0**i) + k
            x0 = np.floor(xs)
            f = xs - x0
            i0 = x0.astype(np.int64)
            i1 = i0 + 1
            v0 = np.take_along_axis(lv64, np.clip(i0, 0, wi - 1)[..., None], axis=-1)[..., 0]
            v1 = np.take_along_axis(lv64, np.clip(i1, 0, wi - 1)[..., None], axis=-1)[..., 0]
            v0 = np.where((i0 >= 0) & (i0 < wi), v0, 0.0)
            v1 = np.where((i1 >= 0) & (i1 < wi), v1, 0.0)
            out[:, i * (2 * r + 1) + (k + r)] = (1.0 - f) * v0 + f * v1
    return out[:, :, :, pad[0] : w - pad[1]]


# --------------------------------------------------------------------------------------
# A5 - truncation mask,  A6 - depth-bin masks,  A7 - training-only corruption pieces
# --------------------------------------------------------------------------------------


def aten_truncation_mask(disp: torch.Tensor, conf: torch.Tensor, gain: float, conf_th=None) -> torch.Tensor:
    """T[b,0,h,w2,w3] = (1-c) + c (sigmoid((w2 - d) - w3)(1-g) + g).

    utils/utils.py:216-238 (`truncate_corr_volume_v2`); the model calls it with conf_th=None
    and attenuation_gain = args.mirror_attenuation (stereoanywhere.py:203).
    """
    b, _, h, w = disp.shape
    cols = torch.arange(w, dtype=disp.dtype, device=disp.device)
    if conf_th is not None:
        conf = (conf > conf_th).to(disp.dtype)
    c = conf.unsqueeze(4)
    centre = cols.view(1, 1, 1, w, 1) - disp.unsqueeze(4)
    arg = centre - cols.view(1, 1, 1, 1, w)
    return 1 * (1 - c) + c * (torch.sigmoid(arg) * (1 - gain) + gain)


def closed_truncation_mask(disp: np.ndarray, conf: np.ndarray, gain: float) -> np.ndarray:
    """float64 closed form; disp/conf [B,H,W] -> [B,H,W,W]."""
    w = disp.shape[-1]
    cols = np.arange(w, dtype=np.float64)
    arg = (cols[None, None, :, None] - disp.astype(np.float64)[..., None]) - cols[None, None, None, :]
    c = conf.astype(np.float64)[..., None]
    return (1.0 - c) + c * ((1.0 / (1.0 + np.exp(-arg))) * (1.0 - gain) + gain)


def aten_depth_bin_masks(mde: torch.Tensor, n: int) -> torch.Tensor:
    """One-hot depth bins [i/N,(i+1)/N) stored as fp16 (utils/utils.py:48-54)."""
    b, _, h, w = mde.shape
    masks = torch.zeros(b, n, h, w, dtype=torch.float16, device=mde.device)
    for i in range(n):
        masks[:, i] = ((mde >= i / n) & (mde < (i + 1) / n)).squeeze(1)
    return masks


def aten_masked_volume(vol: torch.Tensor, masks_l: torch.Tensor, masks_r: torch.Tensor) -> torch.Tensor:
    """vol [B,1,H,W2,W3] x mL[B,N,H,W2,1] x mR[B,N,H,1,W3] (stereoanywhere.py:161)."""
    return vol * masks_l.unsqueeze(4) * masks_r.unsqueeze(3)


def aten_gauss_volume(disp: torch.Tensor, gauss_k: float = 10.0, gauss_c: float = 1.0) -> torch.Tensor:
    """k exp(-((w2 - d) - w3)^2 / (2 c^2)) -> [B,1,H,W,W] (utils/utils.py:200-214)."""
    b, _, h, w = disp.shape
    cols = torch.arange(w, dtype=disp.dtype, device=disp.device)
    centre = cols.view(1, 1, 1, w, 1) - disp.unsqueeze(4)
    arg = centre - cols.view(1, 1, 1, 1, w)
    return gauss_k * torch.exp(-(arg**2) / (2 * gauss_c**2))


def aten_corrupt_roll(vol: torch.Tensor, bin_mask: torch.Tensor, shift: int) -> torch.Tensor:
    """Volume-rolling corruption inside one depth bin (stereoanywhere.py:218-221).

    vol [B,1,H,W2,W3], bin_mask [B,1,H,W2] -> blend of vol and vol rolled along W2.
    """
    m = bin_mask.unsqueeze(4).to(vol.dtype)
    return vol * (1 - m) + torch.roll(vol, shifts=shift, dims=3) * m


def aten_corrupt_scale(vol: torch.Tensor, bin_mask: torch.Tensor, curve: torch.Tensor) -> torch.Tensor:
    """Noise / gaussian corruption: vol (1-m) + vol curve m (stereoanywhere.py:224-233).

    `curve` broadcasts against [B,1,H,W2,W3]: per-left-pixel noise [B,1,H,W2,1] or the gaussian
    volume of `aten_gauss_volume`.
    """
    m = bin_mask.unsqueeze(4).to(vol.dtype)
    return vol * (1 - m) + vol * curve * m


# --------------------------------------------------------------------------------------
# The whole path, as the model drives it (used for the timed CPU baseline)
# --------------------------------------------------------------------------------------


class OracleCorrBlock:
    """Restatement of the reference block protocol on top of the aten_* functions.

    Same constructor / call / static `corr` protocol as corr.py:75-132 so that tests can
    drive the oracle and the B200 block through identical code.
    """

    def __init__(self, fullcorr, num_levels=4, radius=4, pad=(0, 0)):
        self.num_levels = num_levels
        self.radius = radius
        self.pad = list(pad)
        self.fullcorr = fullcorr
        self.corr_pyramid = aten_pyramid(fullcorr, num_levels)

    def __call__(self, coords):
        return aten_lookup(self.corr_pyramid, coords, self.radius, self.num_levels, self.pad)

    corr = staticmethod(aten_corr_volume)


# --------------------------------------------------------------------------------------
# SURVEY 8f-4 - closed-form adjoint of lookup + pyramid (what autograd computes through
# grid_sample / avg_pool2d in the reference's training step, train.py:277,383)
# --------------------------------------------------------------------------------------


def closed_lookup_backward(vol_shape, coords_list, weights_list, num_levels: int = 4, radius: int = 4) -> np.ndarray:
    """d/dV of sum_k <lookup(V, coords_k), weights_k> in float64; returns [B,H,W2,W3].

    Lookup taps (corr.py:93-115): out[b, i*(2r+1)+k] = (1-f) P_i[x0+k-r] + f P_i[x0+k-r+1] with zero padding; the
    pyramid is P_{i+1}[j] = 0.5 (P_i[2j] + P_i[2j+1]) (corr.py:88-91), so dP_i[m] += 0.5 dP_{i+1}[m >> 1]."""
    b, h, w2, _, w3 = vol_shape
    widths = [w3]
    for _ in range(num_levels - 1):
        widths.append(widths[-1] // 2)
    dl = [np.zeros((b, h, w2, w), dtype=np.float64) for w in widths]
    nt = 2 * radius + 1
    for coords, wts in zip(coords_list, weights_list):
        x = coords[:, 0].astype(np.float64)                     # [B,H,W2]
        for i, w in enumerate(widths):
            xs = x / (2 ** i)
            x0 = np.floor(xs)
            f = xs - x0
            for k in range(nt):
                g = wts[:, i * nt + k].astype(np.float64)       # [B,H,W2]
                for off, wgt in ((0, 1.0 - f), (1, f)):
                    c = (x0 + k - radius + off).astype(np.int64)
                    ok = (c >= 0) & (c < w)
                    bi, hi, wi = np.nonzero(ok)
                    np.add.at(dl[i], (bi, hi, wi, c[ok]), (wgt * g)[ok])
    for i in range(num_levels - 1, 0, -1):
        n = widths[i]
        dl[i - 1][..., : 2 * n] += 0.5 * np.repeat(dl[i], 2, axis=-1)
    return dl[0]


# --------------------------------------------------------------------------------------
# SURVEY 8f-2 - soft-argmax disparities / entropy confidences of an aggregated volume
# --------------------------------------------------------------------------------------


def aten_estimate_left_disparity(vol: torch.Tensor, vol_pad: Sequence[int] = (0, 0)) -> torch.Tensor:
    """x - E_{softmax over w3}[w3], `[B,1,H,W2]` (utils/utils.py:112-131; called at stereoanywhere.py:174)."""
    b, _, h, w2, w3 = vol.shape
    idx = torch.arange(0, w3, dtype=vol.dtype).view(1, 1, 1, -1).repeat(b, 1, 1, 1)
    prob = F.softmax(vol.squeeze(1) * 1.0, dim=3)
    expect = torch.sum(prob * idx, 3, keepdim=False)
    xs = torch.arange(w2, dtype=vol.dtype).view(1, 1, w2).repeat(b, h, 1)
    return (xs - expect).unsqueeze(1)[:, :, :, vol_pad[0]: w2 - vol_pad[1]]


def aten_estimate_right_disparity(vol: torch.Tensor, vol_pad: Sequence[int] = (0, 0)) -> torch.Tensor:
    """E_{softmax over w2}[w2] - x, `[B,1,H,W3]` (utils/utils.py:133-152; stereoanywhere.py:175)."""
    b, _, h, w2, w3 = vol.shape
    idx = torch.arange(0, w2, dtype=vol.dtype).view(1, 1, -1, 1).repeat(b, 1, 1, 1)
    prob = F.softmax(vol.squeeze(1) * 1.0, dim=2)
    expect = torch.sum(prob * idx, 2, keepdim=False)
    xs = torch.arange(w3, dtype=vol.dtype).view(1, 1, w3).repeat(b, h, 1)
    return (expect - xs).unsqueeze(1)[:, :, :, vol_pad[0]: w3 - vol_pad[1]]


def aten_estimate_left_confidence(vol: torch.Tensor) -> torch.Tensor:
    """1 - H(softmax over w3) / log2(W3), `[B,1,H,W2]` (utils/utils.py:154-161; stereoanywhere.py:176)."""
    w3 = vol.shape[4]
    prob = F.softmax(vol.squeeze(1), dim=3)
    ent = -torch.sum(prob * torch.log2(prob + 1e-6), dim=3, keepdim=False) / math.log2(w3)
    return (1 - ent).unsqueeze(1)


def aten_estimate_right_confidence(vol: torch.Tensor) -> torch.Tensor:
    """1 - H(softmax over w2) / log2(W2), `[B,1,H,W3]` (utils/utils.py:163-170; stereoanywhere.py:177)."""
    w2 = vol.shape[3]
    prob = F.softmax(vol.squeeze(1), dim=2)
    ent = -torch.sum(prob * torch.log2(prob + 1e-6), dim=2, keepdim=False) / math.log2(w2)
    return (1 - ent).unsqueeze(1)


def closed_volume_reductions(vol: np.ndarray):
    """float64 closed forms of the four reductions; vol [B,1,H,W2,W3] -> (dl, dr, cl, cr)."""
    v = vol.astype(np.float64)[:, 0]
    b, h, w2, w3 = v.shape

    def softmax(x, axis):
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return e / e.sum(axis=axis, keepdims=True)

    p3, p2 = softmax(v, 3), softmax(v, 2)
    dl = np.arange(w2).reshape(1, 1, w2) - (p3 * np.arange(w3).reshape(1, 1, 1, w3)).sum(3)
    dr = (p2 * np.arange(w2).reshape(1, 1, w2, 1)).sum(2) - np.arange(w3).reshape(1, 1, w3)
    cl = 1 + (p3 * np.log2(p3 + 1e-6)).sum(3) / math.log2(w3)
    cr = 1 + (p2 * np.log2(p2 + 1e-6)).sum(2) / math.log2(w2)
    return dl[:, None], dr[:, None], cl[:, None], cr[:, None]


def run_path_cpu(
    fmap_l: torch.Tensor,
    fmap_r: torch.Tensor,
    normals_l: torch.Tensor,
    normals_r: torch.Tensor,
    coords_seq: Sequence[torch.Tensor],
    trunc: Tuple[torch.Tensor, torch.Tensor, float] | None = None,
    radius: int = 4,
    num_levels: int = 4,
):
    """2 x corr + (truncation) + 2 x pyramid + len(coords_seq) x 2 lookups, reference op sequence.

    Mirrors the call sites stereoanywhere.py:135-136, :203, :253-259, :270-271 for the
    `use_aggregate_mono_vol=False` wiring (the mono lookup volume is the raw mono volume), which
    is the wiring every path-only number in BASELINE.md uses.
    """
    v_s = aten_corr_volume(fmap_l, fmap_r).squeeze(3).unsqueeze(1)
    v_m = aten_mono_corr_volume(normals_l, normals_r).squeeze(3).unsqueeze(1)
    if trunc is not None:
        t = aten_truncation_mask(trunc[0], trunc[1], trunc[2])
        v_s = t * v_s
    blk_s = OracleCorrBlock(v_s.squeeze(1).unsqueeze(3), num_levels=num_levels, radius=radius)
    blk_m = OracleCorrBlock(v_m.squeeze(1).unsqueeze(3), num_levels=num_levels, radius=radius)
    outs = None
    for c in coords_seq:
        outs = (blk_s(c), blk_m(c))
    return outs
