"""Producer chain of the path's inputs, without host syncs or per-sample Python loops.

Two forms.  `mono_inputs` and `weighted_lsq_b200` are the CUDA kernels (csrc/producers.cu, CUDA tensors only):
resize + normals + depth bins in one launch, and the per-sample scale / shift fit in one launch for the whole
batch (exact radix-select quantiles, normal equations in double).  The functions below them are the same
arithmetic as device-agnostic batched tensor ops (the host mirror the CPU tests pin against the reference).

The reference forms the inputs of the mono volume and of the truncation mask with a handful of small ops on
`[B,1,H/4,W/4]` maps (stereoanywhere.py:109-114, 138-139, 191): a bilinear 1/4 resize, `estimate_normals`
(utils/utils.py:73-77), `generate_masks` (:48-54) and the per-sample `weighted_lsq` (:345-384), whose Python loop
runs two `torch.quantile`s, boolean-mask gathers (a device->host sync each) and a `torch.linalg.lstsq` per
sample.  These maps are far below a megabyte per pair, so the cost is launches and syncs, not bytes: at batch 64 the
reference's `weighted_lsq` alone issues 5 762 launches and 576 device->host syncs in front of the hot path; the two
kernels are two launches and no sync (profiles/r2/producers_launches_syncs.txt).  The host mirror is the same arithmetic as batched tensor ops
that never leave the stream: plain PyTorch, any device.

Parity: `tests/golden/producers.npz` (generated from the reference).  `generate_masks` is bit-exact;
`weighted_lsq` solves the same 2-parameter weighted least-squares problem through its normal equations in
float64 instead of a QR `lstsq` (agreement ~1e-5 relative); `estimate_normals` sits on kornia's
`spatial_gradient(mode="diff")`, unpinned upstream (requirements.txt:3) - the fixture uses the stub of
oracle/ref_shim.py (replicate pad, central difference without the 1/2 factor).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def mono_inputs(mde: torch.Tensor, n_downsample: int = 2, normal_gain: Optional[float] = None, n_bins: int = 8,
                with_masks: bool = True):
    """`(mde_lowres [B,1,H/4,W/4], normals [B,3,H/4,W/4], masks [B,N,H/4,W/4] fp16 or None)` of a full-resolution mono
    depth `[B,1,H,W]` in ONE kernel (`sa_mono_inputs`): the resize of stereoanywhere.py:109-110, `estimate_normals`
    (:113-114; `normal_gain` defaults to the model's `W_lowres / 10`, :46,113) and `generate_masks` (:138-139)."""
    from . import _lib, ops

    ops._cuda_f32(mde, "mde")
    ops._req(mde.dim() == 4 and mde.shape[1] == 1, "mde must be [B,1,H,W]")
    b, _, h, w = mde.shape
    mde = mde.contiguous()
    hl, wl = int(h * (1.0 / 2 ** n_downsample)), int(w * (1.0 / 2 ** n_downsample))
    if normal_gain is None:
        normal_gain = (w // (2 ** n_downsample)) / 10
    low = torch.empty((b, 1, hl, wl), dtype=torch.float32, device=mde.device)
    normals = torch.empty((b, 3, hl, wl), dtype=torch.float32, device=mde.device)
    masks = torch.empty((b, n_bins, hl, wl), dtype=torch.float16, device=mde.device) if with_masks else None
    lib = _lib.load()
    with ops._on(mde.device):
        rc = lib.sa_mono_inputs(mde.data_ptr(), b, h, w, n_downsample, float(normal_gain), ops.bin_edges(n_bins), n_bins,
                                low.data_ptr(), normals.data_ptr(), masks.data_ptr() if with_masks else None,
                                ops._stream_ptr(mde))
    _lib.check(rc, "sa_mono_inputs")
    return low, normals, masks


def weighted_lsq_b200(mde: torch.Tensor, disp: torch.Tensor, conf: torch.Tensor, min_quantile: float = 0.2,
                      max_quantile: float = 0.9) -> Tuple[torch.Tensor, torch.Tensor]:
    """`weighted_lsq` (utils/utils.py:345-384) for the whole batch in ONE launch, no host sync (`sa_weighted_lsq`).
    Same call shape as the reference: `[B,C,H,W]` maps (the model passes left and right stacked on C,
    stereoanywhere.py:191), returns `(scale, shift)` as `[B,1,1,1]`."""
    from . import _lib, ops

    b = mde.shape[0]
    dt = mde.dtype
    mono, st, cf = (t.reshape(b, -1).float().contiguous() for t in (mde, disp, conf))
    for t, n in ((mono, "mde"), (st, "disp"), (cf, "conf")):
        ops._cuda_f32(t, n)
    ops._req(mono.shape == st.shape == cf.shape, "mde / disp / conf must have the same number of elements per sample")
    scale = torch.empty(b, dtype=torch.float32, device=mono.device)
    shift = torch.empty_like(scale)
    lib = _lib.load()
    with ops._on(mono.device):
        rc = lib.sa_weighted_lsq(mono.data_ptr(), st.data_ptr(), cf.data_ptr(), b, mono.shape[1], float(min_quantile),
                                 float(max_quantile), scale.data_ptr(), shift.data_ptr(), ops._stream_ptr(mono))
    _lib.check(rc, "sa_weighted_lsq")
    return scale.reshape(b, 1, 1, 1).to(dt), shift.reshape(b, 1, 1, 1).to(dt)


def generate_masks(mde: torch.Tensor, N: int = 16) -> torch.Tensor:
    """One-hot depth bins `[i/N, (i+1)/N)` as fp16 `[B,N,H,W]` (utils/utils.py:48-54) in one comparison pass.
    The thresholds are the reference's Python floats `i/N`, compared in the map's dtype like the scalar
    comparisons of the reference; a pixel with mde == 1.0 falls in no bin."""
    edges = torch.tensor([i / N for i in range(N + 1)], dtype=torch.float64, device=mde.device).to(mde.dtype)
    lo, hi = edges[:-1].view(1, N, 1, 1), edges[1:].view(1, N, 1, 1)
    return ((mde < hi) & (mde >= lo)).to(torch.float16)


def spatial_gradient_diff(x: torch.Tensor) -> torch.Tensor:
    """`kornia.filters.spatial_gradient(x, mode="diff", order=1, normalized=False)`: `[B,C,H,W] -> [B,C,2,H,W]`,
    replicate-padded central differences `x[i+1] - x[i-1]` (no 1/2 factor)."""
    p = F.pad(x, (1, 1, 1, 1), mode="replicate")
    gx = p[..., 1:-1, 2:] - p[..., 1:-1, :-2]
    gy = p[..., 2:, 1:-1] - p[..., :-2, 1:-1]
    return torch.stack([gx, gy], dim=2)


def estimate_normals(depth: torch.Tensor, normal_gain: float) -> torch.Tensor:
    """Unit normals `[B,3,H,W]` of a depth map (utils/utils.py:73-77)."""
    g = -spatial_gradient_diff(normal_gain * depth).squeeze(1)          # B 2 H W
    n = torch.cat([g, torch.ones_like(g[:, 0:1])], 1)
    return n / torch.linalg.norm(n, dim=1, keepdim=True)


def lowres(mde: torch.Tensor, n_downsample: int = 2) -> torch.Tensor:
    """The 1 / 2^n bilinear resize of the mono depth (stereoanywhere.py:109-110)."""
    return F.interpolate(mde, scale_factor=1 / (2 ** n_downsample), mode="bilinear", align_corners=True)


def weighted_lsq(mde: torch.Tensor, disp: torch.Tensor, conf: torch.Tensor, min_quantile: float = 0.2,
                 max_quantile: float = 0.9) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-sample scale / shift of the mono depth against the coarse disparity (utils/utils.py:345-384), batched:
    the quantile window, the confidence weights `sqrt(0.9 |c| + 0.1)` and the weighted least squares
    `min sum w^2 (s |mono| + t - |disp|)^2` over the pixels inside the window, for all samples at once.
    No Python loop, no boolean-mask gather, no host sync."""
    b = mde.shape[0]
    dt = mde.dtype
    mono = mde.reshape(b, -1).float().abs()
    stereo = F.relu(disp.reshape(b, -1).float())
    cf = conf.reshape(b, -1).float().abs() * (1 - 0.1) + 0.1
    q = torch.quantile(stereo, torch.tensor([min_quantile, max_quantile], device=stereo.device), dim=1)  # [2,B]
    inside = ((q[0].unsqueeze(1) <= stereo) & (stereo <= q[1].unsqueeze(1))).double()
    w2 = cf.double() * inside                     # (sqrt(conf))^2 restricted to the window
    m, s = mono.double(), stereo.abs().double()
    a00, a01, a11 = (w2 * m * m).sum(1), (w2 * m).sum(1), w2.sum(1)
    b0, b1 = (w2 * m * s).sum(1), (w2 * s).sum(1)
    det = a00 * a11 - a01 * a01
    scale = (a11 * b0 - a01 * b1) / det
    shift = (a00 * b1 - a01 * b0) / det
    return scale.reshape(b, 1, 1, 1).to(dt), shift.reshape(b, 1, 1, 1).to(dt)
