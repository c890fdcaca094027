"""`CorrBlockB200` - drop-in for the reference's correlation block.

Same protocol as `CorrBlock1D` (reference models/stereoanywhere/corr.py:75-132): a static
`corr(fmapL, fmapR)`, a constructor taking the `[B,H,W2,1,W3]` volume with `num_levels`, `radius`,
`pad`, and `__call__(coords)` returning `[B, num_levels*(2*radius+1), H, W]`.  Selected inside the
reference model by adding one branch at stereoanywhere.py:128-133, e.g.

    elif self.args.corr_implementation == "b200":
        corr_block = CorrBlockB200

All arithmetic runs in the sm_100a kernels behind `torch.ops.sa_b200.*`; there is no CPU path.
Training: a volume / feature map that requires grad makes `corr()`, the constructor and
`__call__` autograd nodes whose backward runs `sa_lookup_backward` / `sa_pyramid_backward` / `sa_corr_backward_tf32`
(both products of the adjoint of `corr()` on the tensor cores; the C = 3 mono volume and "fp32" precision take two
library GEMMs); coords are detached before every lookup in the reference (stereoanywhere.py:268) and must not
require grad here.

Beyond the strict protocol (used when the caller's wiring allows, reference call sites in
brackets):
  * `CorrBlockB200(vol, truncate=(disp, conf, gain))`  - builds the pyramid of T*vol in the same
    pass that forms the product  [stereoanywhere.py:203, 253-255];
  * `CorrBlockB200.mono_corr(nL, nR)`                  - A2 with the `1.73 *` folded in  [:136];
  * `CorrBlockB200.from_normals(nL, nR)`               - A2 + pyramid in one pass, volume never written
    [:136, :257-259];
  * `CorrBlockB200.from_features(fL, fR, truncate=...)` - A1 + A5 + pyramid in one tensor-core kernel, volume never
    written  [:135, :203, :253-255];
  * `CorrBlockB200.lookup_pair(stereo_fn, mono_fn, coords)` - both per-iteration lookups in one
    launch  [:270-271];
  * `masked_volume(...)`, `truncation_mask(...)`, `corrupt_volume(...)` - A6 / A5 / A7.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops

_OPS = torch.ops.sa_b200


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _truncate_args(truncate):
    """(disp, conf, gain) of `truncate=` as the kernels take them: fp32 maps (half-precision maps arrive under
    autocast), python-float gain."""
    if truncate is None:
        return None
    return (truncate[0].detach().float(), truncate[1].detach().float(), float(truncate[2]))


def _no_grad_check(*tensors):
    if _needs_grad(*tensors):
        raise NotImplementedError(
            "this stereoanywhere_b200 entry point is forward-only for these inputs: they must not require grad "
            "(gradients flow to the volume / feature maps only; coords are detached before every lookup in the "
            "reference, stereoanywhere.py:268)")


class _CorrFn(torch.autograd.Function):
    """`corr()` with a backward: dL = dV . R / sqrt(C), dR = dV^T . L / sqrt(C).  In "tf32" precision both products
    run on the tensor cores (`sa_corr_backward_tf32`: the feature maps are K-major operands straight from NCHW);
    "fp32" (and the C = 3 mono volume) takes two batched fp32 library GEMMs."""

    @staticmethod
    def forward(ctx, f2, f3, prec, post_scale):
        ctx.save_for_backward(f2, f3)
        ctx.post_scale = post_scale
        ctx.prec = prec
        return _OPS.corr_volume(f2, f3, prec, post_scale)

    @staticmethod
    def backward(ctx, gvol):
        f2, f3 = ctx.saved_tensors
        need2, need3 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        c, w2, w3 = f2.shape[1], f2.shape[3], f3.shape[3]
        if ctx.prec == "tf32" and gvol.is_cuda and ops.corr_backward_ok(c, w2, w3):
            d2, d3 = ops.corr_backward(gvol.float(), f2, f3, ctx.post_scale, need2, need3)
            return d2, d3, None, None
        scale = ctx.post_scale / float(torch.sqrt(torch.tensor(c)))
        g = gvol.squeeze(3) * scale                                   # [B,H,W2,W3]
        d2 = torch.einsum("bhwv,bchv->bchw", g, f3) if need2 else None
        d3 = torch.einsum("bhwv,bchw->bchv", g, f2) if need3 else None
        return d2, d3, None, None


class _GradState:
    """What the backward of a block needs, and nothing that points back at the block or its autograd handle (a
    block -> handle -> grad_fn -> ctx -> block cycle would keep the packed pyramid - 1.47 GB per volume at KITTI
    size, batch 8 - alive until Python's cyclic GC happens to run)."""

    __slots__ = ("shape", "widths", "truncate", "radius", "pad", "dlevels")

    def __init__(self, shape, widths, truncate, radius, pad):
        self.shape, self.widths, self.truncate, self.radius, self.pad = shape, list(widths), truncate, radius, list(pad)
        self.dlevels: Optional[List[torch.Tensor]] = None   # level-gradient accumulators, filled by the lookups' backward


class _PyramidFn(torch.autograd.Function):
    """Ties the block to the volume it was built from: returns a 1-element handle; its backward runs after every
    lookup's backward has accumulated into the state's level-gradient buffers and folds them into dV."""

    @staticmethod
    def forward(ctx, fullcorr, state):
        ctx.state = state
        return fullcorr.new_zeros(1)

    @staticmethod
    def backward(ctx, _gh):
        st = ctx.state
        b, h, w2, w3 = st.shape
        if st.dlevels is None:  # no lookup took part in the loss
            return torch.zeros((b, h, w2, 1, w3), dtype=torch.float32, device=_gh.device), None
        t = st.truncate
        d0 = ops._pyramid_backward(st.dlevels, st.widths, t[0] if t else None, t[1] if t else None, t[2] if t else 0.0)
        st.dlevels = None
        return d0.view(b, h, w2, 1, w3), None


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, handle, coords, block):
        ctx.state = block._grad          # not the block: see _GradState
        ctx.save_for_backward(coords)
        return block._lookup_nograd(coords)

    @staticmethod
    def backward(ctx, gout):
        st = ctx.state
        (coords,) = ctx.saved_tensors
        b, h, w2, w3 = st.shape
        if st.dlevels is None:
            rows = b * h * w2
            st.dlevels = [torch.zeros((rows, w3 if i == 0 else ops.level_pitch(w)), dtype=torch.float32, device=gout.device)
                          for i, w in enumerate(st.widths)]
        p0, p1 = st.pad
        if p0 or p1:
            gout = torch.nn.functional.pad(gout, (p0, p1))
        ops._lookup_backward(gout.float(), coords, st.dlevels, st.widths, st.radius, p0)
        return torch.zeros(1, dtype=torch.float32, device=gout.device), None, None


class CorrBlockB200:
    """B200-native correlation block; see module docstring."""

    #: precision of `corr()` for the stereo volume: "tf32" (tcgen05) or "fp32" (SIMT FMA);
    #: SA_B200_PRECISION overrides the default at import time.
    precision = os.environ.get("SA_B200_PRECISION", "tf32")

    def __init__(self, fullcorr: torch.Tensor, num_levels: int = 4, radius: int = 4, pad: Sequence[int] = (0, 0), *,
                 truncate: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None):
        if truncate is not None:
            _no_grad_check(truncate[0], truncate[1])  # the mask is detached in the reference (stereoanywhere.py:203)
        if fullcorr.dim() != 5 or fullcorr.shape[3] != 1:
            raise ValueError("fullcorr must be [B, H, W2, 1, W3] (reference corr.py:86)")
        b, h, w2, _, w3 = fullcorr.shape
        if not fullcorr.is_contiguous():
            fullcorr = fullcorr.contiguous()
        self._reset(num_levels, radius, pad, (b, h, w2, w3))
        grad_src = fullcorr if _needs_grad(fullcorr) else None
        # any dtype is accepted, like the reference (bilinear_sampler casts to float and back, utils/utils.py:19-35):
        # under `--mixed_precision` the hourglass / classifier volumes arrive in fp16 (test.py:63,189)
        fullcorr = fullcorr.detach().float()
        self._src = fullcorr            # the volume the lookups see (fp32; the tensor handed in when it already was)
        self._truncate = _truncate_args(truncate)
        rows = fullcorr.view(b * h * w2, w3)
        if self.layout == "packed" and ops.packable(num_levels, radius, w3, self.pad):
            # one 128-byte line per (pixel, volume) lookup; levels are materialised only on request
            t = self._truncate
            self._packed = _OPS.pack_pyramid(rows, t[0] if t else None, t[1] if t else None, t[2] if t else 0.0)
        else:
            self._build_levels()
        if grad_src is not None:  # training: lookups go through autograd Functions
            self._grad = _GradState(self._shape, self._widths, self._truncate, self.radius, self.pad)
            self._handle = _PyramidFn.apply(grad_src.float() if grad_src.dtype != torch.float32 else grad_src, self._grad)

    def _reset(self, num_levels: int, radius: int, pad: Sequence[int], shape: Tuple[int, int, int, int]) -> None:
        """Every field of a block, in its empty state; the constructors fill in what they build."""
        self.num_levels, self.radius, self.pad = num_levels, radius, list(pad)
        self._shape = shape                                   # (B, H, W2, W3)
        self._widths: List[int] = ops.level_widths(shape[3], num_levels)
        self._src: Optional[torch.Tensor] = None              # the un-truncated volume, if it exists
        self._features: Optional[Tuple[torch.Tensor, torch.Tensor]] = None         # from_features: (fL, fR)
        self._normals: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None   # from_normals: (nL, nR, gain)
        self._truncate: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None
        self._levels: Optional[List[torch.Tensor]] = None     # 16-byte-pitched levels (general layout)
        self._packed: Optional[torch.Tensor] = None           # line-packed pyramid of the volume
        self._packed_h: Optional[torch.Tensor] = None         # line-packed pyramid in 16-bit storage (from_features, `storage`)
        self._half_kind = 0                                   # 1 = fp16, 2 = bf16 (of _packed_h)
        self._packed_nr: Optional[torch.Tensor] = None        # mono_mode "factored": packed rows of the right normals
        self._otf = False                                     # mono_mode "otf": lookups computed from the normals
        self._grad: Optional[_GradState] = None               # backward state (training only)
        self._handle: Optional[torch.Tensor] = None           # autograd tie to the volume (training only)

    #: "packed" (default; used whenever num_levels=4, radius=4, W3 % 8 == 0, pad=[0,0]) or "levels"
    layout = os.environ.get("SA_B200_LAYOUT", "packed")

    @classmethod
    def from_normals(cls, normals2: torch.Tensor, normals3: torch.Tensor, num_levels: int = 4, radius: int = 4,
                     gain: float = 1.73) -> "CorrBlockB200":
        """The mono block of stereoanywhere.py:136 + :257-259 in one pass: the lookup structure of
        `gain * corr(nL, nR)` is written straight from the normal maps; the volume itself is only formed
        if `fullcorr` / `corr_pyramid` are read.  Same values as `cls(cls.mono_corr(nL, nR))` - bit for bit with
        `mono_mode` "packed" / "otf", to fp32 rounding with the default "factored" (see `mono_mode`)."""
        if cls.mono_mode not in ("factored", "packed", "otf"):
            raise ValueError(f"CorrBlockB200.mono_mode must be 'factored', 'packed' or 'otf' (got {cls.mono_mode!r})")
        b, c, h, w2 = normals2.shape
        w3 = normals3.shape[3]
        if _needs_grad(normals2, normals3) or not (cls.layout == "packed" and ops.packable(num_levels, radius, w3, [0, 0])):
            return cls(cls.mono_corr(normals2, normals3, gain), num_levels=num_levels, radius=radius)
        self = cls.__new__(cls)
        self._reset(num_levels, radius, (0, 0), (b, h, w2, w3))
        self._normals = (normals2.float(), normals3.float(), float(gain))
        self._otf = cls.mono_mode == "otf"  # on-the-fly lookups (see mono_mode)
        if cls.mono_mode == "factored":     # packed pyramid of the right normal map's B*3*H rows (see mono_mode)
            self._packed_nr = _OPS.pack_pyramid(self._normals[1].contiguous().view(b * 3 * h, w3), None, None, 0.0)
        elif not self._otf:
            self._packed = _OPS.pack_pyramid_normals(self._normals[0], self._normals[1], float(gain))
        return self

    #: how a block built by `from_normals` serves its lookups.
    #: "factored" (default) uses the volume's rank: V = gain * nL^T nR / sqrt 3, and the pyramid and the packed
    #: layout are linear in V, so only the right normal map's B*3*H rows are packed (14 MB instead of 1.47 GB at
    #: KITTI size, batch 8; L2-resident, 10 us) and the lookup kernel combines three of their lines with the pixel's
    #: left normal while staging: no mono pack pass (241 us) and half the lookup's DRAM reads; within fp32 rounding
    #: (<= 1e-6 abs) of the packed mode.  "packed" writes the packed pyramid of the volume itself, bit-identical to
    #: `cls(cls.mono_corr(nL, nR))`.  "otf" keeps only the normal maps and forms 80 level-0 values per pixel inside
    #: the lookup kernel - bit-identical to "packed" too, but 47 us instead of 20 us per dual lookup (a memory-saving mode).
    #: SA_B200_MONO overrides the default.
    mono_mode = os.environ.get("SA_B200_MONO", "factored")

    #: storage of the packed pyramid written by `from_features`: "fp32" (default: 128-byte lines), "fp16" or "bf16"
    #: (64-byte lines: `sa_corr_pack_tf32` writes half the bytes, 293 instead of 370 us at KITTI size, batch 8).  Opt-in:
    #: the stored values are the fp32 ones rounded to nearest - fp16 adds at most 2^-11 = 4.9e-4 of max|V| (measured
    #: 3.2e-4; together with the TF32 product still inside the 1e-3 tolerance), bf16 3.9e-3 (the 1e-2 class); the
    #: lookup widens them and is otherwise unchanged.  SA_B200_STORAGE overrides the default.
    storage = os.environ.get("SA_B200_STORAGE", "fp32")

    def _ensure_packed(self) -> torch.Tensor:
        """The fp32 packed pyramid of this block (built on demand for on-the-fly / factored mono blocks and for
        stereo blocks held in 16-bit storage)."""
        if self._packed is None and self._packed_h is not None:
            t = self._truncate
            self._packed = _OPS.corr_pack(self._features[0], self._features[1], t[0] if t else None, t[1] if t else None,
                                          t[2] if t else 0.0)
        if self._packed is None and (self._otf or self._packed_nr is not None):
            self._packed = _OPS.pack_pyramid_normals(self._normals[0], self._normals[1], self._normals[2])
            self._otf = False
            self._packed_nr = None
        return self._packed

    @classmethod
    def from_features(cls, fmap2: torch.Tensor, fmap3: torch.Tensor, num_levels: int = 4, radius: int = 4, *,
                      truncate: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None,
                      storage: Optional[str] = None) -> "CorrBlockB200":
        """The stereo block of stereoanywhere.py:135 + :203 + :253-255 in ONE kernel: correlation on the tensor
        cores, truncation product and pyramid in the GEMM epilogue, written once as the packed pyramid
        (csrc/corr_pack_tcgen05.cu).  Same values, bit for bit, as
        `cls(cls.corr(fmap2, fmap3), truncate=truncate)` in tf32 precision; the volume is only formed if
        `fullcorr` / `corr_pyramid` are read.  Falls back to that two-step form for shapes the fused kernel
        does not cover, or when `precision == "fp32"`."""
        b, c, h, w2 = fmap2.shape
        w3 = fmap3.shape[3]
        if _needs_grad(fmap2, fmap3) or not (cls.layout == "packed" and cls.precision == "tf32"
                                             and ops.packable(num_levels, radius, w3, [0, 0]) and ops.corr_packable(c, w2, w3)):
            return cls(cls.corr(fmap2, fmap3), num_levels=num_levels, radius=radius, truncate=truncate)
        self = cls.__new__(cls)
        self._reset(num_levels, radius, (0, 0), (b, h, w2, w3))
        self._features = (fmap2.float(), fmap3.float())
        self._truncate = _truncate_args(truncate)
        t = self._truncate
        storage = cls.storage if storage is None else storage
        if storage == "fp32":
            self._packed = _OPS.corr_pack(self._features[0], self._features[1], t[0] if t else None, t[1] if t else None,
                                          t[2] if t else 0.0)
        elif storage in ops.HALF_KINDS:   # opt-in 16-bit storage of the packed pyramid (see `storage`)
            self._half_kind = ops.HALF_KINDS[storage][0]
            self._packed_h = _OPS.corr_pack_half(self._features[0], self._features[1], t[0] if t else None,
                                                 t[1] if t else None, t[2] if t else 0.0, self._half_kind)
        else:
            raise ValueError(f"storage must be 'fp32', 'fp16' or 'bf16' (got {storage!r})")
        return self

    def _source(self) -> torch.Tensor:
        """The un-truncated volume; formed on demand for blocks built by from_normals / from_features."""
        if self._src is None:
            if self._features is not None:
                self._src = CorrBlockB200.corr(*self._features)
            else:
                self._src = CorrBlockB200.mono_corr(*self._normals)
        return self._src

    def _build_levels(self):
        if self._levels is None:
            b, h, w2, w3 = self._shape
            rows = self._source().view(b * h * w2, w3)
            t = self._truncate
            self._levels = ops.pyramid_levels(rows, self.num_levels, t[0] if t else None, t[1] if t else None,
                                              t[2] if t else 0.0)
        return self._levels

    @property
    def fullcorr(self) -> torch.Tensor:
        """The volume the lookups see, `[B,H,W2,1,W3]` (reference attribute, corr.py:83).  With
        `truncate=` this is the product T*V (formed on first access when the packed layout is in use)."""
        if self._truncate is None:
            return self._source()
        b, h, w2, w3 = self._shape
        return self._build_levels()[0].view(b, h, w2, 1, w3)

    @property
    def corr_pyramid(self) -> List[torch.Tensor]:
        """Levels as `[B*H*W2, 1, 1, W3_i]` views like the reference attribute (corr.py:87-91).
        (Levels >= 1 are views into 16-byte-pitched rows; the reference's dead extra level is absent.
        With the packed layout they are built on first access.)"""
        return [lv[:, :w].unsqueeze(1).unsqueeze(1) for lv, w in zip(self._build_levels(), self._widths)]

    def _lookup_nograd(self, coords: torch.Tensor) -> torch.Tensor:
        if self._otf:
            return _OPS.lookup_normals(self._normals[0], self._normals[1], self._normals[2], coords)
        if self._packed_nr is not None:
            return _OPS.lookup_factored(self._packed_nr, self._normals[0], self._normals[2], coords)
        if self._packed_h is not None:
            return _OPS.lookup_half(self._packed_h, self._half_kind, self._shape[3], coords)
        if self._packed is not None:
            return _OPS.lookup_packed(self._packed, self._shape[3], coords)
        return _OPS.lookup(self._levels, self._widths, coords, self.radius, self.pad[0], self.pad[1])

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        _no_grad_check(coords)
        dt = coords.dtype
        if dt != torch.float32:
            coords = coords.float()
        if self._handle is not None and torch.is_grad_enabled():
            out = _LookupFn.apply(self._handle, coords, self)
        else:
            out = self._lookup_nograd(coords)
        return out if dt == torch.float32 else out.to(dt)

    # ---- protocol: static corr --------------------------------------------------------------
    @staticmethod
    def corr(fmap2: torch.Tensor, fmap3: torch.Tensor) -> torch.Tensor:
        """[B,C,H,W2] x [B,C,H,W3] -> [B,H,W2,1,W3], divided by sqrt(C) (reference corr.py:117-132).

        C % 32 == 0 (the kernel's channel slab) and 4-aligned widths take the tensor-core kernel in `CorrBlockB200.precision`;
        anything else (the C=3 normals volume in particular) takes the fp32 SIMT kernel."""
        dt = fmap2.dtype
        f2, f3 = fmap2.float(), fmap3.float()
        prec = CorrBlockB200.precision
        c, w2, w3 = f2.shape[1], f2.shape[3], f3.shape[3]
        if prec != "fp32" and not (c % 32 == 0 and w2 % 4 == 0 and w3 % 4 == 0 and w3 <= 1024):
            prec = "fp32"
        if _needs_grad(f2, f3):
            vol = _CorrFn.apply(f2, f3, prec, 1.0)
        else:
            vol = _OPS.corr_volume(f2, f3, prec, 1.0)
        return vol if dt == torch.float32 else vol.to(dt)

    # ---- beyond the protocol ------------------------------------------------------------------
    @staticmethod
    def mono_corr(normals2: torch.Tensor, normals3: torch.Tensor, gain: float = 1.73) -> torch.Tensor:
        """`1.73 * corr(nL, nR)` in one pass (stereoanywhere.py:136)."""
        n2, n3 = normals2.float(), normals3.float()
        if _needs_grad(n2, n3):
            return _CorrFn.apply(n2, n3, "fp32", float(gain))
        return _OPS.corr_volume(n2, n3, "fp32", float(gain))

    @staticmethod
    def _pairable(block_a: "CorrBlockB200", block_b: "CorrBlockB200") -> bool:
        """Whether `lookup_pair` can serve both blocks from one launch (same geometry, no pad, block_a holding a
        packed or levels pyramid of its own, neither block under autograd)."""
        if torch.is_grad_enabled() and (block_a._handle is not None or block_b._handle is not None):
            return False  # training: each lookup is its own autograd node
        otf_b = block_b._otf or block_b._packed_nr is not None  # block_b holds no packed volume of its own
        if block_a._packed_h is not None:   # 16-bit stereo block: with a factored or an fp32-packed partner
            return (block_a._shape == block_b._shape and block_b.pad == [0, 0] and block_b._packed_h is None
                    and not block_b._otf and (block_b._packed_nr is not None or block_b._packed is not None))
        if block_b._packed_h is not None:
            return False
        return not (block_a.radius != block_b.radius or block_a._widths != block_b._widths
                    or block_a._shape != block_b._shape
                    or block_a.pad != [0, 0] or block_b.pad != [0, 0] or block_a._otf
                    or block_a._packed_nr is not None
                    or (block_a._packed is None) != (block_b._packed is None and not otf_b))

    @staticmethod
    def lookup_pair(block_a: "CorrBlockB200", block_b: "CorrBlockB200", coords: torch.Tensor):
        """`(block_a(coords), block_b(coords))` with one launch (stereoanywhere.py:270-271)."""
        _no_grad_check(coords)
        if not CorrBlockB200._pairable(block_a, block_b):
            return CorrBlockB200.__call__(block_a, coords), CorrBlockB200.__call__(block_b, coords)
        fact_b = block_b._packed_nr is not None
        dt = coords.dtype
        if dt != torch.float32:
            coords = coords.float()
        if block_a._packed_h is not None:
            if fact_b:
                oa, ob = _OPS.lookup_half2(block_a._packed_h, block_a._half_kind, 2, block_b._packed_nr, block_b._normals[0],
                                           block_b._normals[2], block_a._shape[3], coords)
            else:
                oa, ob = _OPS.lookup_half2(block_a._packed_h, block_a._half_kind, 1, block_b._packed, None, 1.0,
                                           block_a._shape[3], coords)
        elif block_a._packed is not None and fact_b:
            oa, ob = _OPS.lookup_packed_factored2(block_a._packed, block_b._packed_nr, block_b._normals[0],
                                                  block_b._normals[2], coords)
        elif block_a._packed is not None and block_b._otf:
            oa, ob = _OPS.lookup_packed_normals2(block_a._packed, block_b._normals[0], block_b._normals[1],
                                                 block_b._normals[2], coords)
        elif block_a._packed is not None:
            oa, ob = _OPS.lookup_packed2(block_a._packed, block_b._packed, block_a._shape[3], coords)
        else:
            oa, ob = _OPS.lookup2(block_a._levels, block_b._levels, block_a._widths, coords, block_a.radius)
        return (oa, ob) if dt == torch.float32 else (oa.to(dt), ob.to(dt))


def lookup_pair_convc1(block_a: CorrBlockB200, block_b: CorrBlockB200, coords: torch.Tensor, weight: torch.Tensor,
                       bias: torch.Tensor):
    """`(relu(convc1(block_a(coords))), relu(convc1(block_b(coords))))` in one kernel - the front end of
    `BasicMotionEncoder.forward` (update.py:80-84; convc1 = Conv2d(36, 64, 1), shared by both volumes) fused
    into the lookups of stereoanywhere.py:270-271.  TF32 tensor-core product, fp32 accumulate.
    Falls back to two lookups + torch convolutions when the blocks are not in the packed layout."""
    _no_grad_check(coords, weight, bias)
    block_a._ensure_packed()
    convc1 = weight.shape[0] == 64 and weight.shape[1] == 36
    if block_b._packed_nr is not None and block_a._packed is not None and block_a._shape == block_b._shape and convc1:
        # factored mono block: its line is combined from the packed right normals inside the kernel
        return _OPS.lookup_factored_conv(block_a._packed, block_b._packed_nr, block_b._normals[0], block_b._normals[2],
                                         coords.float(), weight.float(), bias.float())
    block_b._ensure_packed()  # an on-the-fly mono block is packed on first use here
    if (block_a._packed is None or block_b._packed is None or block_a._shape != block_b._shape or not convc1):
        sa_, sb_ = CorrBlockB200.lookup_pair(block_a, block_b, coords)
        conv = torch.nn.functional.conv2d
        return torch.relu(conv(sa_, weight, bias)), torch.relu(conv(sb_, weight, bias))
    return _OPS.lookup_packed_conv(block_a._packed, block_b._packed, block_a._shape[3], coords.float(),
                                   weight.float(), bias.float())


def truncation_mask(disp: torch.Tensor, conf: torch.Tensor, attenuation_gain: float,
                    vol: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`truncate_corr_volume_v2(disp, conf, conf_th=None, attenuation_gain)` (utils/utils.py:216-238);
    with `vol` ([B,1,H,W2,W3]) returns the product `mask * vol` of stereoanywhere.py:253-255."""
    _no_grad_check(disp, conf, vol)
    return _OPS.truncate(vol, disp, conf, float(attenuation_gain))


def masked_volume(vol: torch.Tensor, mde_l: torch.Tensor, mde_r: torch.Tensor, n_bins: int = 8) -> torch.Tensor:
    """`vol * generate_masks(mdeL,N).unsqueeze(4) * generate_masks(mdeR,N).unsqueeze(3)`:
    [B,1,H,W2,W3] -> [B,N,H,W2,W3] (utils/utils.py:48-54, stereoanywhere.py:138-139,161)."""
    _no_grad_check(vol)
    return _OPS.masked_volume(vol, None, None, 1.0, mde_l, mde_r, n_bins)


def masked_mono_volume(normals_l: torch.Tensor, normals_r: torch.Tensor, mde_l: torch.Tensor, mde_r: torch.Tensor,
                       n_bins: int = 8, gain: float = 1.73) -> torch.Tensor:
    """A2 + A6 fused: the hourglass input straight from the normal maps (the mono volume is never
    materialised on its own)."""
    _no_grad_check(normals_l, normals_r)
    return _OPS.masked_volume(None, normals_l, normals_r, float(gain), mde_l, mde_r, n_bins)


def corrupt_volume(vol: torch.Tensor, bin_mask: torch.Tensor, mode: str, *, shift: int = 0,
                   noise: Optional[torch.Tensor] = None, gauss_k: float = 0.0) -> torch.Tensor:
    """Training-only corruption blends of stereoanywhere.py:214-251 (`mode` = roll | noise | gauss)."""
    code = {"roll": 0, "noise": 1, "gauss": 2}[mode]
    return _OPS.corrupt(vol, bin_mask, code, int(shift), noise, float(gauss_k))
