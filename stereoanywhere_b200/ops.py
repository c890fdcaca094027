"""torch custom ops (`torch.ops.sa_b200.*`) over the C-ABI library.

Host code is plumbing only: argument checks, output allocation from torch's caching allocator,
the current CUDA stream, one ctypes call.  No CPU implementation is registered - calling an op
with CPU tensors raises (the dispatcher has no CPU kernel), and a missing `libsa_b200.so` raises
at first use.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

MAX_LEVELS = 8
_PYR_ALIGN = 4  # floats: level pitches are multiples of 16 bytes so the lookup can use 128-bit loads


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(t: torch.Tensor) -> int:
    """cudaStream_t of torch's current stream on the tensor's device (raw handle: ~0.3 us instead of ~2 us for
    the Stream object - the lookup runs 64 times per forward and its kernel takes 3 us at batch 1)."""
    if _raw_stream is not None:
        idx = t.device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(t.device).cuda_stream


#: How `acc / sqrt(C)` (corr.py:132) is rounded.  "reciprocal": acc * (1.0f / sqrt(C)) - what the reference computes
#: when it runs ON A CUDA DEVICE (ATen's true-division kernel multiplies by the fp32 reciprocal of a scalar divisor),
#: so the C = 3 mono volume is bit-identical to the reference's on the same GPU.  "exact": correctly rounded division
#: - the reference on the CPU (the golden fixtures).  They differ by one ulp in ~1/4 of the entries for sqrt(3) and
#: not at all for the stereo volume (sqrt(256) = 16).  The one ulp matters: `weighted_lsq` (utils/utils.py:345-384,
#: out of scope) selects pixels by quantile thresholds and can amplify it to 0.1 px of final disparity.
DIVISION = os.environ.get("SA_B200_DIVISION", "reciprocal")


def _divisor(c: int) -> float:
    """sqrt(C) as the reference forms it (a float32 scalar, corr.py:132); negative = reciprocal convention of the C ABI."""
    if DIVISION not in ("reciprocal", "exact"):
        raise ValueError(f"stereoanywhere_b200.ops.DIVISION must be 'reciprocal' or 'exact' (got {DIVISION!r})")
    d = float(torch.sqrt(torch.tensor(c)))
    return -d if DIVISION == "reciprocal" else d


def packed_row_floats(w3: int) -> int:
    """Floats per volume row of the packed layout = sa_packed_row_floats(w3) (csrc/packed.cu), without the call."""
    return (w3 // 8 + 9) * 32


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NOGUARD = _NoGuard()


def _on(dev: torch.device):
    """Device guard that costs nothing in the common single-device-per-process case."""
    if dev.index is None or torch.cuda.current_device() == dev.index:
        return _NOGUARD
    return torch.cuda.device(dev)


def _req(cond: bool, msg: str):
    if not cond:
        raise ValueError(msg)


def _cuda_f32(t: torch.Tensor, name: str):
    _req(t.is_cuda, f"{name} must be a CUDA tensor (stereoanywhere_b200 has no CPU path)")
    _req(t.dtype == torch.float32, f"{name} must be float32, got {t.dtype}")


def level_widths(w: int, num_levels: int) -> List[int]:
    """Valid widths of the pyramid levels (floor halving, reference corr.py:88-91)."""
    out = [w]
    for _ in range(num_levels - 1):
        out.append(out[-1] // 2)
    return out


def level_pitch(w: int) -> int:
    return (w + _PYR_ALIGN - 1) // _PYR_ALIGN * _PYR_ALIGN


# ------------------------------------------------------------------------------------------
# implementations (CUDA only)
# ------------------------------------------------------------------------------------------


def _corr_volume(fmap_l: torch.Tensor, fmap_r: torch.Tensor, precision: str, post_scale: float) -> torch.Tensor:
    _cuda_f32(fmap_l, "fmap_l")
    _cuda_f32(fmap_r, "fmap_r")
    _req(fmap_l.dim() == 4 and fmap_r.dim() == 4, "feature maps must be [B,C,H,W]")
    b, c, h, w2 = fmap_l.shape
    _req(fmap_r.shape[:3] == (b, c, h), "left/right feature maps must share B, C, H")
    w3 = fmap_r.shape[3]
    fmap_l = fmap_l.contiguous()
    fmap_r = fmap_r.contiguous()
    vol = torch.empty((b, h, w2, 1, w3), dtype=torch.float32, device=fmap_l.device)
    # the reference divides by torch.sqrt(torch.tensor(C)): a float32 scalar (corr.py:132)
    divisor = _divisor(c)
    lib = _lib.load()
    with _on(fmap_l.device):
        if precision == "fp32":
            rc = lib.sa_corr_fp32(fmap_l.data_ptr(), fmap_r.data_ptr(), vol.data_ptr(), b, c, h, w2, w3, divisor,
                                  post_scale, _stream_ptr(vol))
        elif precision == "tf32":
            rc = lib.sa_corr_tf32(fmap_l.data_ptr(), fmap_r.data_ptr(), vol.data_ptr(), b, c, h, w2, w3, divisor,
                                  post_scale, None, None, 0.0, None, None, None, 0, 0, 0, _stream_ptr(vol))
        else:
            raise ValueError(f"unknown precision {precision!r} (use 'tf32' or 'fp32')")
    _lib.check(rc, f"sa_corr_{precision}")
    return vol


def _pyramid(vol: torch.Tensor, num_levels: int, trunc_disp: Optional[torch.Tensor],
             trunc_conf: Optional[torch.Tensor], trunc_gain: float) -> List[torch.Tensor]:
    """vol: [rows, W] (contiguous view of the block's volume). Returns the NEW tensors only (a functional op must
    not return its input): levels 1.. [rows, pitch_i], preceded by the truncated level 0 when trunc_* are given.
    Without truncation level 0 is `vol` itself - `pyramid_levels` below prepends it."""
    _cuda_f32(vol, "fullcorr")
    _req(vol.dim() == 2 and vol.is_contiguous(), "pyramid expects a contiguous [rows, W] view")
    _req(1 <= num_levels <= MAX_LEVELS, f"num_levels must be in 1..{MAX_LEVELS}")
    rows, w = vol.shape
    widths = level_widths(w, num_levels)
    _req(widths[-1] >= 1, f"volume width {w} is too small for {num_levels} levels")
    lib = _lib.load()
    dev = vol.device
    trunc = trunc_disp is not None
    levels = [vol]
    if trunc:
        _cuda_f32(trunc_disp, "trunc_disp")
        _cuda_f32(trunc_conf, "trunc_conf")
        trunc_disp = trunc_disp.contiguous()
        trunc_conf = trunc_conf.contiguous()
        _req(trunc_disp.numel() == rows and trunc_conf.numel() == rows,
             "truncation maps must be [B,1,H,W2] matching the volume")
        w2_size = trunc_disp.shape[-1]
        levels = [torch.empty_like(vol)]
    with _on(dev):
        st = _stream_ptr(vol)
        if num_levels == 1:
            if trunc:
                rc = lib.sa_truncate(vol.data_ptr(), trunc_disp.data_ptr(), trunc_conf.data_ptr(), trunc_gain,
                                     levels[0].data_ptr(), rows, w2_size, w, st)
                _lib.check(rc, "sa_truncate")
            return levels if trunc else []
        src, src_w, src_pitch = vol, w, w
        nxt = 1
        first = True
        while nxt < num_levels:
            n_out = min(3, num_levels - nxt)
            outs = [torch.empty((rows, level_pitch(widths[nxt + k])), dtype=torch.float32, device=dev)
                    for k in range(n_out)]
            ptr = [o.data_ptr() for o in outs] + [None] * (3 - n_out)
            pit = [o.shape[1] for o in outs] + [0] * (3 - n_out)
            if first and trunc:
                rc = lib.sa_pyramid(src.data_ptr(), rows, src_w, src_pitch, n_out, ptr[0], ptr[1], ptr[2], pit[0],
                                    pit[1], pit[2], trunc_disp.data_ptr(), trunc_conf.data_ptr(), trunc_gain,
                                    w2_size, levels[0].data_ptr(), st)
            else:
                rc = lib.sa_pyramid(src.data_ptr(), rows, src_w, src_pitch, n_out, ptr[0], ptr[1], ptr[2], pit[0],
                                    pit[1], pit[2], None, None, 0.0, 0, None, st)
            _lib.check(rc, "sa_pyramid")
            levels.extend(outs)
            nxt += n_out
            src, src_w, src_pitch = outs[-1], widths[nxt - 1], outs[-1].shape[1]
            first = False
    return levels if trunc else levels[1:]


def pyramid_levels(vol: torch.Tensor, num_levels: int, trunc_disp: Optional[torch.Tensor],
                   trunc_conf: Optional[torch.Tensor], trunc_gain: float) -> List[torch.Tensor]:
    """All `num_levels` levels of `vol` ([rows, W]); level 0 is `vol` itself unless a truncation is applied."""
    new = list(torch.ops.sa_b200.pyramid(vol, num_levels, trunc_disp, trunc_conf, trunc_gain))
    return new if trunc_disp is not None else [vol] + new


def _marshal(levels: Sequence[torch.Tensor], widths: Sequence[int]):
    n = len(levels)
    _req(1 <= n <= MAX_LEVELS and len(widths) == n, f"need 1..{MAX_LEVELS} levels with matching widths")
    rows = levels[0].shape[0]
    for t in levels:
        _cuda_f32(t, "pyramid level")
        _req(t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == rows, "pyramid levels must be [rows, pitch] row-major")
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in levels])
    wid = (C.c_int * n)(*widths)
    pit = (C.c_int64 * n)(*[t.stride(0) for t in levels])
    return n, ptrs, wid, pit


def _coords_view(coords: torch.Tensor):
    _cuda_f32(coords, "coords")
    _req(coords.dim() == 4 and coords.shape[1] >= 1, "coords must be [B,2,H,W]")
    b, _, h, w = coords.shape
    if not (coords.stride(3) == 1 and coords.stride(2) == w):
        coords = coords.contiguous()
    return coords, b, h, w


def _lookup(levels: Sequence[torch.Tensor], widths: Sequence[int], coords: torch.Tensor, radius: int, pad0: int,
            pad1: int) -> torch.Tensor:
    coords, b, h, w = _coords_view(coords)
    n, ptrs, wid, pit = _marshal(levels, widths)
    _req(levels[0].shape[0] == b * h * w, "coords do not match the volume this block was built from")
    _req(0 <= pad0 and 0 <= pad1 and pad0 + pad1 < w, "bad pad")
    out = torch.empty((b, n * (2 * radius + 1), h, w - pad0 - pad1), dtype=torch.float32, device=coords.device)
    lib = _lib.load()
    with _on(coords.device):
        rc = lib.sa_lookup(ptrs, wid, pit, n, radius, coords.data_ptr(), coords.stride(0), out.data_ptr(), b, h, w,
                           pad0, pad1, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup")
    return out


def _lookup2(levels_a: Sequence[torch.Tensor], levels_b: Sequence[torch.Tensor], widths: Sequence[int],
             coords: torch.Tensor, radius: int) -> Tuple[torch.Tensor, torch.Tensor]:
    coords, b, h, w = _coords_view(coords)
    n, ptrs_a, wid, pit_a = _marshal(levels_a, widths)
    nb, ptrs_b, _, pit_b = _marshal(levels_b, widths)
    _req(n == nb, "lookup2 needs two pyramids with the same number of levels")
    _req(levels_a[0].shape[0] == b * h * w and levels_b[0].shape[0] == b * h * w, "coords do not match the volumes")
    out_a = torch.empty((b, n * (2 * radius + 1), h, w), dtype=torch.float32, device=coords.device)
    out_b = torch.empty_like(out_a)
    lib = _lib.load()
    with _on(coords.device):
        rc = lib.sa_lookup2(ptrs_a, ptrs_b, wid, pit_a, pit_b, n, radius, coords.data_ptr(), coords.stride(0),
                            out_a.data_ptr(), out_b.data_ptr(), b, h, w, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup2")
    return out_a, out_b


def packable(num_levels: int, radius: int, w3: int, pad) -> bool:
    """The line-packed fast path covers the model's configuration (corr_levels=4, corr_radius=4,
    quarter-resolution width a multiple of 8, no pad)."""
    return num_levels == 4 and radius == 4 and w3 >= 8 and w3 % 8 == 0 and list(pad) == [0, 0]


def _pack_pyramid(vol: torch.Tensor, trunc_disp: Optional[torch.Tensor], trunc_conf: Optional[torch.Tensor],
                  trunc_gain: float) -> torch.Tensor:
    """vol [rows, W3] -> packed [rows, (W3/8 + 9) * 32] (csrc/packed.cu)."""
    _cuda_f32(vol, "fullcorr")
    _req(vol.dim() == 2 and vol.is_contiguous(), "pack expects a contiguous [rows, W] view")
    rows, w = vol.shape
    _req(w >= 8 and w % 8 == 0, "packed pyramid needs W3 % 8 == 0")
    lib = _lib.load()
    packed = torch.empty((rows, packed_row_floats(w)), dtype=torch.float32, device=vol.device)
    with _on(vol.device):
        if trunc_disp is not None:
            _cuda_f32(trunc_disp, "trunc_disp")
            _cuda_f32(trunc_conf, "trunc_conf")
            trunc_disp, trunc_conf = trunc_disp.contiguous(), trunc_conf.contiguous()
            _req(trunc_disp.numel() == rows and trunc_conf.numel() == rows,
                 "truncation maps must be [B,1,H,W2] matching the volume")
            rc = lib.sa_pack_pyramid(vol.data_ptr(), rows, w, trunc_disp.data_ptr(), trunc_conf.data_ptr(), trunc_gain,
                                     trunc_disp.shape[-1], packed.data_ptr(), _stream_ptr(vol))
        else:
            rc = lib.sa_pack_pyramid(vol.data_ptr(), rows, w, None, None, 0.0, 0, packed.data_ptr(), _stream_ptr(vol))
    _lib.check(rc, "sa_pack_pyramid")
    return packed


def _pack_pyramid_normals(normals_l: torch.Tensor, normals_r: torch.Tensor, post_scale: float) -> torch.Tensor:
    """A2 + A3 fused: packed pyramid of post_scale * corr(nL, nR) straight from the normal maps."""
    _cuda_f32(normals_l, "normals_l")
    _cuda_f32(normals_r, "normals_r")
    b, c, h, w2 = normals_l.shape
    w3 = normals_r.shape[3]
    _req(c == 3 and normals_r.shape[:3] == (b, 3, h), "normals must be [B,3,H,W]")
    _req(w3 >= 8 and w3 % 8 == 0, "packed pyramid needs W3 % 8 == 0")
    normals_l, normals_r = normals_l.contiguous(), normals_r.contiguous()
    lib = _lib.load()
    packed = torch.empty((b * h * w2, packed_row_floats(w3)), dtype=torch.float32, device=normals_l.device)
    divisor = _divisor(3)
    with _on(normals_l.device):
        rc = lib.sa_pack_pyramid_normals(normals_l.data_ptr(), normals_r.data_ptr(), divisor, post_scale, b, h, w2, w3,
                                         packed.data_ptr(), _stream_ptr(packed))
    _lib.check(rc, "sa_pack_pyramid_normals")
    return packed


def _corr_pack(fmap_l: torch.Tensor, fmap_r: torch.Tensor, trunc_disp: Optional[torch.Tensor],
               trunc_conf: Optional[torch.Tensor], trunc_gain: float) -> torch.Tensor:
    """A1 (+A5) + A3 fused: packed pyramid of [T *] corr(L, R) straight from the feature maps
    (csrc/corr_pack_tcgen05.cu); the volume is never written."""
    _cuda_f32(fmap_l, "fmap_l")
    _cuda_f32(fmap_r, "fmap_r")
    _req(fmap_l.dim() == 4 and fmap_r.dim() == 4, "feature maps must be [B,C,H,W]")
    b, c, h, w2 = fmap_l.shape
    _req(fmap_r.shape[:3] == (b, c, h), "left/right feature maps must share B, C, H")
    w3 = fmap_r.shape[3]
    _req(corr_packable(c, w2, w3), "corr_pack needs C % 32 == 0, W2 % 4 == 0, W3 % 8 == 0")
    fmap_l, fmap_r = fmap_l.contiguous(), fmap_r.contiguous()
    lib = _lib.load()
    rows = b * h * w2
    packed = torch.empty((rows, packed_row_floats(w3)), dtype=torch.float32, device=fmap_l.device)
    divisor = _divisor(c)
    td = tc = None
    if trunc_disp is not None:
        _cuda_f32(trunc_disp, "trunc_disp")
        _cuda_f32(trunc_conf, "trunc_conf")
        trunc_disp, trunc_conf = trunc_disp.contiguous(), trunc_conf.contiguous()
        _req(trunc_disp.numel() == rows and trunc_conf.numel() == rows,
             "truncation maps must be [B,1,H,W2] matching the left feature map")
        td, tc = trunc_disp.data_ptr(), trunc_conf.data_ptr()
    with _on(fmap_l.device):
        rc = lib.sa_corr_pack_tf32(fmap_l.data_ptr(), fmap_r.data_ptr(), b, c, h, w2, w3, divisor, 1.0, td, tc,
                                   float(trunc_gain), packed.data_ptr(), _stream_ptr(packed))
    _lib.check(rc, "sa_corr_pack_tf32")
    return packed


HALF_KINDS = {"fp16": (1, torch.float16), "bf16": (2, torch.bfloat16)}


def _corr_pack_half(fmap_l: torch.Tensor, fmap_r: torch.Tensor, trunc_disp: Optional[torch.Tensor],
                    trunc_conf: Optional[torch.Tensor], trunc_gain: float, kind: int) -> torch.Tensor:
    """`_corr_pack` with the packed pyramid stored in 16 bits (kind 1 = fp16, 2 = bf16): [rows, (W3/8 + 9) * 32] of
    that dtype (csrc/corr_pack_tcgen05.cu, sa_corr_pack_tf32_half)."""
    _cuda_f32(fmap_l, "fmap_l")
    _cuda_f32(fmap_r, "fmap_r")
    _req(kind in (1, 2), "half storage kind must be 1 (fp16) or 2 (bf16)")
    b, c, h, w2 = fmap_l.shape
    _req(fmap_r.shape[:3] == (b, c, h), "left/right feature maps must share B, C, H")
    w3 = fmap_r.shape[3]
    _req(corr_packable(c, w2, w3), "corr_pack needs C % 32 == 0, W2 % 4 == 0, W3 % 8 == 0")
    fmap_l, fmap_r = fmap_l.contiguous(), fmap_r.contiguous()
    rows = b * h * w2
    packed = torch.empty((rows, packed_row_floats(w3)), dtype=torch.float16 if kind == 1 else torch.bfloat16,
                         device=fmap_l.device)
    td = tc = None
    if trunc_disp is not None:
        _cuda_f32(trunc_disp, "trunc_disp")
        _cuda_f32(trunc_conf, "trunc_conf")
        trunc_disp, trunc_conf = trunc_disp.contiguous(), trunc_conf.contiguous()
        _req(trunc_disp.numel() == rows and trunc_conf.numel() == rows,
             "truncation maps must be [B,1,H,W2] matching the left feature map")
        td, tc = trunc_disp.data_ptr(), trunc_conf.data_ptr()
    lib = _lib.load()
    with _on(fmap_l.device):
        rc = lib.sa_corr_pack_tf32_half(fmap_l.data_ptr(), fmap_r.data_ptr(), b, c, h, w2, w3, _divisor(c), 1.0, td, tc,
                                        float(trunc_gain), kind, packed.data_ptr(), _stream_ptr(packed))
    _lib.check(rc, "sa_corr_pack_tf32_half")
    return packed


def _lookup_half(packed_h: torch.Tensor, kind: int, mode_b: int, packed_b: Optional[torch.Tensor],
                 normals_l: Optional[torch.Tensor], post_scale: float, w3: int, coords: torch.Tensor):
    """Lookup with volume A in 16-bit storage; mode_b 0 = alone, 1 = with an fp32 packed volume, 2 = with the
    factored mono volume (packed_b = packed right normals).  Returns (out_a, out_b or None)."""
    coords, b, h, w = _coords_view(coords)
    _req(packed_h.is_cuda and packed_h.dtype == (torch.float16 if kind == 1 else torch.bfloat16),
         "packed_h must be a CUDA tensor of the storage dtype")
    _req(packed_h.shape == (b * h * w, packed_row_floats(w3)) and packed_h.is_contiguous(),
         "coords do not match the volume this block was built from")
    out_a = torch.empty((b, 36, h, w), dtype=torch.float32, device=coords.device)
    out_b = torch.empty_like(out_a) if mode_b else None
    pb = nl = None
    divisor = 1.0
    if mode_b == 1:
        _cuda_f32(packed_b, "packed pyramid")
        _req(packed_b.shape == (b * h * w, packed_row_floats(w3)), "lookup of two volumes needs identical geometry")
        pb = packed_b.data_ptr()
    elif mode_b == 2:
        _cuda_f32(packed_b, "packed right normals")
        _cuda_f32(normals_l, "normals_l")
        _req(normals_l.shape == (b, 3, h, w), "normals_l must be [B,3,H,W] matching coords")
        _req(packed_b.dim() == 2 and packed_b.shape == (b * 3 * h, packed_row_floats(w3)) and packed_b.is_contiguous(),
             "packed right normals must be [B*3*H, row floats]")
        normals_l = normals_l.contiguous()
        pb, nl, divisor = packed_b.data_ptr(), normals_l.data_ptr(), _divisor(3)
    lib = _lib.load()
    with _on(coords.device):
        rc = lib.sa_lookup_packed_half(packed_h.data_ptr(), kind, mode_b, pb, nl, divisor, float(post_scale), w3,
                                       coords.data_ptr(), coords.stride(0), out_a.data_ptr(),
                                       out_b.data_ptr() if out_b is not None else None, b, h, w, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_packed_half")
    return out_a, out_b


def _lookup_half1(packed_h, kind, w3, coords):
    return _lookup_half(packed_h, kind, 0, None, None, 1.0, w3, coords)[0]


def _lookup_half2(packed_h, kind, mode_b, packed_b, normals_l, post_scale, w3, coords):
    return _lookup_half(packed_h, kind, mode_b, packed_b, normals_l, post_scale, w3, coords)


def corr_packable(c: int, w2: int, w3: int) -> bool:
    return c % 32 == 0 and c >= 32 and w2 % 4 == 0 and w3 >= 8 and w3 % 8 == 0


def _volume_reduce(vol: torch.Tensor, which: str):
    """[B,1,H,W2,W3] -> (left [B,1,H,W2], right [B,1,H,W3]); `which` = "softargmax" | "entropy_conf"
    (csrc/volume_reduce.cu: one read of the volume for both directions)."""
    _cuda_f32(vol, "corr_volume")
    _req(vol.dim() == 5 and vol.shape[1] == 1, "corr_volume must be [B,1,H,W2,W3]")
    b, _, h, w2, w3 = vol.shape
    _req(w3 <= 1024, "W3 > 1024 is not covered by the reduction kernels")
    vol = vol.contiguous()
    left = torch.empty((b, 1, h, w2), dtype=torch.float32, device=vol.device)
    right = torch.empty((b, 1, h, w3), dtype=torch.float32, device=vol.device)
    lib = _lib.load()
    fn = lib.sa_volume_softargmax if which == "softargmax" else lib.sa_volume_entropy_conf
    with _on(vol.device):
        rc = fn(vol.data_ptr(), b * h, w2, w3, left.data_ptr(), right.data_ptr(), _stream_ptr(vol))
    _lib.check(rc, f"sa_volume_{which}")
    return left, right


def _volume_softargmax(vol: torch.Tensor):
    return _volume_reduce(vol, "softargmax")


def _volume_entropy_conf(vol: torch.Tensor):
    return _volume_reduce(vol, "entropy_conf")


def corr_backward(grad_vol: torch.Tensor, fmap_l: torch.Tensor, fmap_r: torch.Tensor, post_scale: float, need_l: bool,
                  need_r: bool):
    """Adjoint of `corr_volume` on the tensor cores (csrc/corr_bwd_tcgen05.cu): grad_vol [B,H,W2,1,W3] ->
    (grad_l [B,C,H,W2] or None, grad_r [B,C,H,W3] or None).  TF32 operands, fp32 accumulate."""
    _cuda_f32(grad_vol, "grad_vol")
    _cuda_f32(fmap_l, "fmap_l")
    _cuda_f32(fmap_r, "fmap_r")
    b, c, h, w2 = fmap_l.shape
    w3 = fmap_r.shape[3]
    _req(grad_vol.numel() == b * h * w2 * w3, "grad_vol does not match the feature maps")
    grad_vol, fmap_l, fmap_r = grad_vol.contiguous(), fmap_l.contiguous(), fmap_r.contiguous()
    gl = torch.empty_like(fmap_l) if need_l else None
    gr = torch.empty_like(fmap_r) if need_r else None
    lib = _lib.load()
    with _on(grad_vol.device):
        rc = lib.sa_corr_backward_tf32(grad_vol.data_ptr(), fmap_l.data_ptr(), fmap_r.data_ptr(),
                                       gl.data_ptr() if need_l else None, gr.data_ptr() if need_r else None, b, c, h, w2, w3,
                                       _divisor(c), float(post_scale), _stream_ptr(grad_vol))
    _lib.check(rc, "sa_corr_backward_tf32")
    return gl, gr


def corr_backward_ok(c: int, w2: int, w3: int) -> bool:
    return w2 % 4 == 0 and w3 % 4 == 0 and c >= 8


def _lookup_backward(grad_out: torch.Tensor, coords: torch.Tensor, dlevels: List[torch.Tensor], widths: List[int],
                     radius: int, pad0: int) -> None:
    """Accumulate the adjoint of one lookup into the level-gradient buffers `dlevels` (csrc/backward.cu).
    grad_out is [B, L*(2r+1), H, W - pad]; pad != 0 is handled by the caller (it zero-pads grad_out)."""
    coords, b, h, w = _coords_view(coords)
    _cuda_f32(grad_out, "grad_out")
    n, ptrs, wid, pit = _marshal(dlevels, widths)
    _req(grad_out.shape == (b, n * (2 * radius + 1), h, w), "grad_out does not match coords / levels")
    _req(dlevels[0].shape[0] == b * h * w, "level-gradient buffers do not match coords")
    grad_out = grad_out.contiguous()
    lib = _lib.load()
    with _on(coords.device):
        rc = lib.sa_lookup_backward(grad_out.data_ptr(), coords.data_ptr(), coords.stride(0), ptrs, wid, pit, n, radius,
                                    b, h, w, pad0, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_backward")


def _pyramid_backward(dlevels: List[torch.Tensor], widths: List[int], trunc_disp: Optional[torch.Tensor],
                      trunc_conf: Optional[torch.Tensor], trunc_gain: float) -> torch.Tensor:
    """Fold the level gradients into level 0 in place and return it ([rows, W], the gradient w.r.t. the volume)."""
    n, ptrs, wid, pit = _marshal(dlevels, widths)
    rows = dlevels[0].shape[0]
    lib = _lib.load()
    td = tc = None
    w2 = 0
    if trunc_disp is not None:
        trunc_disp, trunc_conf = trunc_disp.contiguous(), trunc_conf.contiguous()
        _cuda_f32(trunc_disp, "trunc_disp")
        _cuda_f32(trunc_conf, "trunc_conf")
        td, tc, w2 = trunc_disp.data_ptr(), trunc_conf.data_ptr(), trunc_disp.shape[-1]
    with _on(dlevels[0].device):
        rc = lib.sa_pyramid_backward(dlevels[0].data_ptr(), ptrs, wid, pit, n, rows, td, tc, float(trunc_gain), w2,
                                     _stream_ptr(dlevels[0]))
    _lib.check(rc, "sa_pyramid_backward")
    return dlevels[0]


def _lookup_packed(packed_a: torch.Tensor, packed_b: Optional[torch.Tensor], w3: int, coords: torch.Tensor):
    coords, b, h, w = _coords_view(coords)
    _cuda_f32(packed_a, "packed pyramid")
    lib = _lib.load()
    rowf = packed_row_floats(w3)
    _req(packed_a.shape == (b * h * w, rowf), "coords do not match the volume this block was built from")
    out_a = torch.empty((b, 36, h, w), dtype=torch.float32, device=coords.device)
    out_b = None
    if packed_b is not None:
        _cuda_f32(packed_b, "packed pyramid")
        _req(packed_b.shape == packed_a.shape, "lookup of two volumes needs identical geometry")
        out_b = torch.empty_like(out_a)
    with _on(coords.device):
        rc = lib.sa_lookup_packed(packed_a.data_ptr(), packed_b.data_ptr() if packed_b is not None else None, w3,
                                  coords.data_ptr(), coords.stride(0), out_a.data_ptr(),
                                  out_b.data_ptr() if out_b is not None else None, b, h, w, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_packed")
    return out_a, out_b


def _lookup_normals(packed_a: Optional[torch.Tensor], normals_l: torch.Tensor, normals_r: torch.Tensor, post_scale: float,
                    coords: torch.Tensor):
    """Lookup with the mono volume computed on the fly from the normal maps (csrc/packed.cu, OTF path); with
    `packed_a` the stereo volume is looked up in the same launch.  Returns (out_a or None, out_mono)."""
    coords, b, h, w = _coords_view(coords)
    _cuda_f32(normals_l, "normals_l")
    _cuda_f32(normals_r, "normals_r")
    _req(normals_l.shape == (b, 3, h, w) and normals_r.shape[:3] == (b, 3, h), "normals must be [B,3,H,W] matching coords")
    w3 = normals_r.shape[3]
    _req(w3 >= 8 and w3 % 8 == 0, "on-the-fly mono lookup needs W3 % 8 == 0")
    normals_l, normals_r = normals_l.contiguous(), normals_r.contiguous()
    out_m = torch.empty((b, 36, h, w), dtype=torch.float32, device=coords.device)
    out_a = None
    if packed_a is not None:
        _cuda_f32(packed_a, "packed pyramid")
        _req(packed_a.shape == (b * h * w, packed_row_floats(w3)), "coords do not match the packed volume")
        out_a = torch.empty_like(out_m)
    lib = _lib.load()
    divisor = _divisor(3)
    with _on(coords.device):
        rc = lib.sa_lookup_packed_normals(packed_a.data_ptr() if packed_a is not None else None, normals_l.data_ptr(),
                                          normals_r.data_ptr(), divisor, float(post_scale), w3, coords.data_ptr(),
                                          coords.stride(0), out_a.data_ptr() if out_a is not None else None,
                                          out_m.data_ptr(), b, h, w, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_packed_normals")
    return out_a, out_m


def _lookup_factored(packed_a: Optional[torch.Tensor], packed_nr: torch.Tensor, normals_l: torch.Tensor,
                     post_scale: float, coords: torch.Tensor):
    """Lookup with the mono volume in factored form (csrc/packed.cu, FV path): `packed_nr` is the packed pyramid
    of the right normal map's B*3*H rows, `normals_l` the left normals.  Returns (out_a or None, out_mono)."""
    coords, b, h, w = _coords_view(coords)
    _cuda_f32(normals_l, "normals_l")
    _cuda_f32(packed_nr, "packed right normals")
    _req(normals_l.shape == (b, 3, h, w), "normals_l must be [B,3,H,W] matching coords")
    _req(packed_nr.dim() == 2 and packed_nr.shape[0] == b * 3 * h and packed_nr.is_contiguous(),
         "packed right normals must be [B*3*H, row floats]")
    w3 = (packed_nr.shape[1] // 32 - 9) * 8
    _req(w3 >= 8 and packed_row_floats(w3) == packed_nr.shape[1], "bad packed row size")
    normals_l = normals_l.contiguous()
    out_m = torch.empty((b, 36, h, w), dtype=torch.float32, device=coords.device)
    out_a = None
    if packed_a is not None:
        _cuda_f32(packed_a, "packed pyramid")
        _req(packed_a.shape == (b * h * w, packed_row_floats(w3)), "coords do not match the packed volume")
        out_a = torch.empty_like(out_m)
    lib = _lib.load()
    divisor = _divisor(3)
    with _on(coords.device):
        rc = lib.sa_lookup_packed_factored(packed_a.data_ptr() if packed_a is not None else None, packed_nr.data_ptr(),
                                           normals_l.data_ptr(), divisor, float(post_scale), w3, coords.data_ptr(),
                                           coords.stride(0), out_a.data_ptr() if out_a is not None else None,
                                           out_m.data_ptr(), b, h, w, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_packed_factored")
    return out_a, out_m


def _lookup_factored1(packed_nr, normals_l, post_scale, coords):
    return _lookup_factored(None, packed_nr, normals_l, post_scale, coords)[1]


def _lookup_packed_factored2(packed_a, packed_nr, normals_l, post_scale, coords):
    return _lookup_factored(packed_a, packed_nr, normals_l, post_scale, coords)


def _lookup_normals1(normals_l, normals_r, post_scale, coords):
    return _lookup_normals(None, normals_l, normals_r, post_scale, coords)[1]


def _lookup_packed_normals2(packed_a, normals_l, normals_r, post_scale, coords):
    return _lookup_normals(packed_a, normals_l, normals_r, post_scale, coords)


def _lookup_packed_conv(packed_a: torch.Tensor, packed_b: torch.Tensor, w3: int, coords: torch.Tensor,
                        weight: torch.Tensor, bias: torch.Tensor):
    """relu(convc1(lookup_a)), relu(convc1(lookup_b)) without materialising the lookups (csrc/lookup_conv.cu)."""
    coords, b, h, w = _coords_view(coords)
    for t, n in ((packed_a, "packed_a"), (packed_b, "packed_b"), (weight, "weight"), (bias, "bias")):
        _cuda_f32(t, n)
    lib = _lib.load()
    rowf = packed_row_floats(w3)
    _req(packed_a.shape == (b * h * w, rowf) and packed_b.shape == packed_a.shape, "coords do not match the volumes")
    _req(weight.numel() == 64 * 36 and bias.numel() == 64, "convc1 must be Conv2d(36, 64, 1): weight [64,36,1,1], bias [64]")
    weight = weight.reshape(64, 36).contiguous()
    bias = bias.contiguous()
    out_a = torch.empty((b, 64, h, w), dtype=torch.float32, device=coords.device)
    out_b = torch.empty_like(out_a)
    with _on(coords.device):
        rc = lib.sa_lookup_packed_conv(packed_a.data_ptr(), packed_b.data_ptr(), w3, coords.data_ptr(), coords.stride(0),
                                       weight.data_ptr(), bias.data_ptr(), out_a.data_ptr(), out_b.data_ptr(), b, h, w,
                                       _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_packed_conv")
    return out_a, out_b


def _lookup_factored_conv(packed_a: torch.Tensor, packed_nr: torch.Tensor, normals_l: torch.Tensor, post_scale: float,
                          coords: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor):
    """`_lookup_packed_conv` with the mono volume in factored form (packed right normals + left normals)."""
    coords, b, h, w = _coords_view(coords)
    for t, n in ((packed_a, "packed_a"), (packed_nr, "packed right normals"), (normals_l, "normals_l"), (weight, "weight"),
                 (bias, "bias")):
        _cuda_f32(t, n)
    _req(normals_l.shape == (b, 3, h, w), "normals_l must be [B,3,H,W] matching coords")
    _req(packed_nr.dim() == 2 and packed_nr.shape[0] == b * 3 * h and packed_nr.is_contiguous(),
         "packed right normals must be [B*3*H, row floats]")
    w3 = (packed_nr.shape[1] // 32 - 9) * 8
    _req(w3 >= 8 and packed_row_floats(w3) == packed_nr.shape[1], "bad packed row size")
    _req(packed_a.shape == (b * h * w, packed_row_floats(w3)), "coords do not match the packed volume")
    _req(weight.numel() == 64 * 36 and bias.numel() == 64, "convc1 must be Conv2d(36, 64, 1): weight [64,36,1,1], bias [64]")
    weight = weight.reshape(64, 36).contiguous()
    bias = bias.contiguous()
    normals_l = normals_l.contiguous()
    out_a = torch.empty((b, 64, h, w), dtype=torch.float32, device=coords.device)
    out_m = torch.empty_like(out_a)
    lib = _lib.load()
    divisor = _divisor(3)
    with _on(coords.device):
        rc = lib.sa_lookup_factored_conv(packed_a.data_ptr(), packed_nr.data_ptr(), normals_l.data_ptr(), divisor,
                                         float(post_scale), w3, coords.data_ptr(), coords.stride(0), weight.data_ptr(),
                                         bias.data_ptr(), out_a.data_ptr(), out_m.data_ptr(), b, h, w, _stream_ptr(coords))
    _lib.check(rc, "sa_lookup_factored_conv")
    return out_a, out_m


def _lookup_packed1(packed: torch.Tensor, w3: int, coords: torch.Tensor) -> torch.Tensor:
    return _lookup_packed(packed, None, w3, coords)[0]


def _lookup_packed2(packed_a: torch.Tensor, packed_b: torch.Tensor, w3: int, coords: torch.Tensor):
    return _lookup_packed(packed_a, packed_b, w3, coords)


def _truncate(vol: Optional[torch.Tensor], disp: torch.Tensor, conf: torch.Tensor, gain: float) -> torch.Tensor:
    _cuda_f32(disp, "disp")
    _cuda_f32(conf, "conf")
    b, _, h, w2 = disp.shape
    disp, conf = disp.contiguous(), conf.contiguous()
    if vol is not None:
        _cuda_f32(vol, "vol")
        vol = vol.contiguous()
        w3 = vol.shape[-1]
        _req(vol.numel() == b * h * w2 * w3, "vol does not match disp")
        out = torch.empty_like(vol)
    else:
        w3 = w2
        out = torch.empty((b, 1, h, w2, w3), dtype=torch.float32, device=disp.device)
    lib = _lib.load()
    with _on(disp.device):
        rc = lib.sa_truncate(vol.data_ptr() if vol is not None else None, disp.data_ptr(), conf.data_ptr(), gain,
                             out.data_ptr(), b * h * w2, w2, w3, _stream_ptr(disp))
    _lib.check(rc, "sa_truncate")
    return out


def bin_edges(n_bins: int):
    """float32(i / N) for i = 0..N, as the reference's comparisons see them (utils/utils.py:51-53)."""
    return (C.c_float * (n_bins + 1))(*[i / n_bins for i in range(n_bins + 1)])


def _masked_volume(vol: Optional[torch.Tensor], normals_l: Optional[torch.Tensor], normals_r: Optional[torch.Tensor],
                   post_scale: float, mde_l: torch.Tensor, mde_r: torch.Tensor, n_bins: int) -> torch.Tensor:
    _cuda_f32(mde_l, "mde_l")
    _cuda_f32(mde_r, "mde_r")
    b, _, h, w2 = mde_l.shape
    w3 = mde_r.shape[-1]
    mde_l, mde_r = mde_l.contiguous(), mde_r.contiguous()
    out = torch.empty((b, n_bins, h, w2, w3), dtype=torch.float32, device=mde_l.device)
    lib = _lib.load()
    with _on(mde_l.device):
        if vol is not None:
            _cuda_f32(vol, "vol")
            vol = vol.contiguous()
            _req(vol.numel() == b * h * w2 * w3, "vol does not match the depth maps")
            rc = lib.sa_masked_volume(vol.data_ptr(), None, None, 1.0, 1.0, mde_l.data_ptr(), mde_r.data_ptr(),
                                      bin_edges(n_bins), n_bins, out.data_ptr(), b, h, w2, w3, _stream_ptr(out))
        else:
            _cuda_f32(normals_l, "normals_l")
            _cuda_f32(normals_r, "normals_r")
            _req(normals_l.shape == (b, 3, h, w2) and normals_r.shape == (b, 3, h, w3), "normals must be [B,3,H,W]")
            normals_l, normals_r = normals_l.contiguous(), normals_r.contiguous()
            divisor = _divisor(3)
            rc = lib.sa_masked_volume(None, normals_l.data_ptr(), normals_r.data_ptr(), divisor, post_scale,
                                      mde_l.data_ptr(), mde_r.data_ptr(), bin_edges(n_bins), n_bins, out.data_ptr(),
                                      b, h, w2, w3, _stream_ptr(out))
    _lib.check(rc, "sa_masked_volume")
    return out


def _corrupt(vol: torch.Tensor, bin_mask: torch.Tensor, mode: int, shift: int, noise: Optional[torch.Tensor],
             gauss_k: float) -> torch.Tensor:
    _cuda_f32(vol, "vol")
    _req(vol.dim() == 5 and vol.shape[1] == 1, "vol must be [B,1,H,W2,W3]")
    b, _, h, w2, w3 = vol.shape
    vol = vol.contiguous()
    bin_mask = bin_mask.to(torch.float32).contiguous()
    _req(bin_mask.numel() == b * h * w2, "bin_mask must be [B,1,H,W2]")
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
        _req(noise.numel() == b * h * w2, "noise must be [B,1,H,W2,1]")
    out = torch.empty_like(vol)
    lib = _lib.load()
    with _on(vol.device):
        rc = lib.sa_corrupt(vol.data_ptr(), bin_mask.data_ptr(), mode, shift,
                            noise.data_ptr() if noise is not None else None, gauss_k, out.data_ptr(), b, h, w2, w3,
                            _stream_ptr(vol))
    _lib.check(rc, "sa_corrupt")
    return out


# ------------------------------------------------------------------------------------------
# torch.library registration: functional ops, CUDA backend only
# ------------------------------------------------------------------------------------------

_LIBDEF = torch.library.Library("sa_b200", "DEF")
_LIBDEF.define("corr_volume(Tensor fmap_l, Tensor fmap_r, str precision, float post_scale) -> Tensor")
_LIBDEF.define("pyramid(Tensor vol_rows, int num_levels, Tensor? trunc_disp, Tensor? trunc_conf, float trunc_gain) -> Tensor[]")
_LIBDEF.define("lookup(Tensor[] levels, int[] widths, Tensor coords, int radius, int pad0, int pad1) -> Tensor")
_LIBDEF.define("lookup2(Tensor[] levels_a, Tensor[] levels_b, int[] widths, Tensor coords, int radius) -> (Tensor, Tensor)")
_LIBDEF.define("pack_pyramid(Tensor vol_rows, Tensor? trunc_disp, Tensor? trunc_conf, float trunc_gain) -> Tensor")
_LIBDEF.define("pack_pyramid_normals(Tensor normals_l, Tensor normals_r, float post_scale) -> Tensor")
_LIBDEF.define("corr_pack(Tensor fmap_l, Tensor fmap_r, Tensor? trunc_disp, Tensor? trunc_conf, float trunc_gain) -> Tensor")
_LIBDEF.define("corr_pack_half(Tensor fmap_l, Tensor fmap_r, Tensor? trunc_disp, Tensor? trunc_conf, float trunc_gain, int kind) -> Tensor")
_LIBDEF.define("lookup_half(Tensor packed_h, int kind, int w3, Tensor coords) -> Tensor")
_LIBDEF.define("lookup_half2(Tensor packed_h, int kind, int mode_b, Tensor packed_b, Tensor? normals_l, float post_scale, int w3, Tensor coords) -> (Tensor, Tensor)")
_LIBDEF.define("volume_softargmax(Tensor vol) -> (Tensor, Tensor)")
_LIBDEF.define("volume_entropy_conf(Tensor vol) -> (Tensor, Tensor)")
_LIBDEF.define("lookup_normals(Tensor normals_l, Tensor normals_r, float post_scale, Tensor coords) -> Tensor")
_LIBDEF.define("lookup_packed_normals2(Tensor packed_a, Tensor normals_l, Tensor normals_r, float post_scale, Tensor coords) -> (Tensor, Tensor)")
_LIBDEF.define("lookup_factored(Tensor packed_nr, Tensor normals_l, float post_scale, Tensor coords) -> Tensor")
_LIBDEF.define("lookup_packed_factored2(Tensor packed_a, Tensor packed_nr, Tensor normals_l, float post_scale, Tensor coords) -> (Tensor, Tensor)")
_LIBDEF.define("lookup_packed(Tensor packed, int w3, Tensor coords) -> Tensor")
_LIBDEF.define("lookup_packed2(Tensor packed_a, Tensor packed_b, int w3, Tensor coords) -> (Tensor, Tensor)")
_LIBDEF.define("lookup_packed_conv(Tensor packed_a, Tensor packed_b, int w3, Tensor coords, Tensor weight, Tensor bias) -> (Tensor, Tensor)")
_LIBDEF.define("lookup_factored_conv(Tensor packed_a, Tensor packed_nr, Tensor normals_l, float post_scale, Tensor coords, Tensor weight, Tensor bias) -> (Tensor, Tensor)")
_LIBDEF.define("truncate(Tensor? vol, Tensor disp, Tensor conf, float gain) -> Tensor")
_LIBDEF.define("masked_volume(Tensor? vol, Tensor? normals_l, Tensor? normals_r, float post_scale, Tensor mde_l, Tensor mde_r, int n_bins) -> Tensor")
_LIBDEF.define("corrupt(Tensor vol, Tensor bin_mask, int mode, int shift, Tensor? noise, float gauss_k) -> Tensor")

_LIBDEF.impl("corr_volume", _corr_volume, "CUDA")
_LIBDEF.impl("pyramid", _pyramid, "CUDA")
_LIBDEF.impl("lookup", _lookup, "CUDA")
_LIBDEF.impl("lookup2", _lookup2, "CUDA")
_LIBDEF.impl("pack_pyramid", _pack_pyramid, "CUDA")
_LIBDEF.impl("pack_pyramid_normals", _pack_pyramid_normals, "CUDA")
_LIBDEF.impl("corr_pack", _corr_pack, "CUDA")
_LIBDEF.impl("corr_pack_half", _corr_pack_half, "CUDA")
_LIBDEF.impl("lookup_half", _lookup_half1, "CUDA")
_LIBDEF.impl("lookup_half2", _lookup_half2, "CUDA")
_LIBDEF.impl("volume_softargmax", _volume_softargmax, "CUDA")
_LIBDEF.impl("volume_entropy_conf", _volume_entropy_conf, "CUDA")
_LIBDEF.impl("lookup_normals", _lookup_normals1, "CUDA")
_LIBDEF.impl("lookup_packed_normals2", _lookup_packed_normals2, "CUDA")
_LIBDEF.impl("lookup_factored", _lookup_factored1, "CUDA")
_LIBDEF.impl("lookup_packed_factored2", _lookup_packed_factored2, "CUDA")
_LIBDEF.impl("lookup_packed", _lookup_packed1, "CUDA")
_LIBDEF.impl("lookup_packed2", _lookup_packed2, "CUDA")
_LIBDEF.impl("lookup_packed_conv", _lookup_packed_conv, "CUDA")
_LIBDEF.impl("lookup_factored_conv", _lookup_factored_conv, "CUDA")
_LIBDEF.impl("truncate", _truncate, "CUDA")
_LIBDEF.impl("masked_volume", _masked_volume, "CUDA")
_LIBDEF.impl("corrupt", _corrupt, "CUDA")

OP_NAMES = ["corr_pack_half", "lookup_half", "lookup_half2", "corr_volume", "pyramid", "lookup", "lookup2", "pack_pyramid", "pack_pyramid_normals", "corr_pack", "lookup_normals", "lookup_packed_normals2", "lookup_factored", "lookup_packed_factored2", "volume_softargmax", "volume_entropy_conf", "lookup_packed", "lookup_packed2", "lookup_packed_conv", "lookup_factored_conv",
            "truncate", "masked_volume", "corrupt"]
