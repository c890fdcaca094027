"""Plugging the B200 block into an UNMODIFIED reference model.

`install(sa_mod)` with `sa_mod = models.stereoanywhere.stereoanywhere` (the reference module) makes the
`"reg"` branch of stereoanywhere.py:128-133 resolve to a B200 block - no edit of the reference tree.

Two modes:

* `fused=False` - `CorrBlock1D := CorrBlockB200`.  The reference's call sequence runs op for op (`corr()` x 2,
  `1.73 *`, truncation mask, product, two constructors, 64 lookups): `corr()`, the constructors (pyramid) and the
  lookups are kernels of this package; the `1.73 *`, the truncation mask and its product stay the reference's own
  PyTorch ops on the dense volume.
* `fused=True`  - the same calls, but the full-volume intermediates are never formed when nobody reads them.
  `corr()` returns a `LazyVolume` (it remembers the two maps), `truncate_corr_volume_v2(..., conf_th=None)` a
  `LazyTruncation` (it remembers disp / conf / gain); their product, the reshapes of stereoanywhere.py:135-136 /
  :253-258 and the `1.73 *` stay symbolic, and the block constructor then builds
  `CorrBlockB200.from_features(fL, fR, truncate=...)` (correlation + truncation + pyramid in one tensor-core kernel)
  or `CorrBlockB200.from_normals(nL, nR)` (factored mono block).  Any other use of a lazy object (the depth-bin
  products that feed the hourglass, stereoanywhere.py:161) materialises it with the plain kernels.  The two
  lookups of an iteration (stereoanywhere.py:270-271, same `coords1` tensor) are served by ONE `lookup_pair`
  launch: the stereo block computes both and hands the mono result to its partner.

Under autograd (training) nothing is lazy: `corr()` returns the dense, differentiable volume.
"""
from __future__ import annotations

import threading
import weakref
from typing import Optional

import torch

from . import ops
from .corr import CorrBlockB200, _needs_grad

_OPS = torch.ops.sa_b200
_tls = threading.local()   # one model replica per thread (nn.DataParallel): pairing state is per thread


class LazyTruncation:
    """`truncate_corr_volume_v2(disp, conf, conf_th=None, attenuation_gain)` (utils/utils.py:216-238), not yet formed."""

    def __init__(self, disp: torch.Tensor, conf: torch.Tensor, gain: float):
        self.disp, self.conf, self.gain = disp, conf, float(gain)

    def detach(self):   # stereoanywhere.py:203
        return self

    def materialize(self) -> torch.Tensor:
        return _OPS.truncate(None, self.disp.float(), self.conf.float(), self.gain)   # [B,1,H,W2,W3]

    def __mul__(self, other):
        if isinstance(other, LazyVolume):
            return other._with_truncation(self)
        return self.materialize() * other

    __rmul__ = __mul__


class LazyVolume:
    """`corr(a, b)` (corr.py:117-132) not yet formed: remembers the two maps, a scalar gain, an optional truncation
    and the symbolic shape ([B,H,W2,1,W3] and its squeeze / unsqueeze images)."""

    def __init__(self, a: torch.Tensor, b: torch.Tensor, gain: float = 1.0, truncation: Optional[LazyTruncation] = None,
                 shape=None):
        self.a, self.b, self.gain, self.truncation = a, b, float(gain), truncation
        bsz, _, h, w2 = a.shape
        self.shape = list(shape) if shape is not None else [bsz, h, w2, 1, b.shape[3]]
        self.dtype, self.device = a.dtype, a.device

    def _like(self, **kw):
        args = dict(a=self.a, b=self.b, gain=self.gain, truncation=self.truncation, shape=self.shape)
        args.update(kw)
        return LazyVolume(**args)

    def _with_truncation(self, t: LazyTruncation):
        if self.truncation is not None:
            return self.materialize() * t.materialize()
        return self._like(truncation=t)

    # ---- the reshapes of stereoanywhere.py:135-136, 253-258 stay symbolic -----------------------------
    def squeeze(self, d):
        d = d % len(self.shape)
        return self._like(shape=self.shape[:d] + self.shape[d + 1:]) if self.shape[d] == 1 else self

    def unsqueeze(self, d):
        d = d % (len(self.shape) + 1)
        return self._like(shape=self.shape[:d] + [1] + self.shape[d:])

    def dim(self):
        return len(self.shape)

    def size(self, d=None):
        return torch.Size(self.shape) if d is None else self.shape[d]

    def detach(self):
        return self

    def float(self):
        return self

    # ---- arithmetic ------------------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return self._like(gain=self.gain * float(other))
        if isinstance(other, LazyTruncation):
            return self._with_truncation(other)
        return self.materialize() * other

    __rmul__ = __mul__

    def materialize(self) -> torch.Tensor:
        """The dense volume in the current symbolic shape."""
        a, b = self.a.float(), self.b.float()
        if a.shape[1] == 3:
            vol = CorrBlockB200.mono_corr(a, b, self.gain)
        else:
            vol = CorrBlockB200.corr(a, b)
            if self.gain != 1.0:
                vol = vol * self.gain
        if self.truncation is not None:
            t = self.truncation
            vol = _OPS.truncate(vol.view(vol.shape[0], 1, vol.shape[1], vol.shape[2], vol.shape[4]), t.disp.float(),
                                t.conf.float(), t.gain)
        return vol.view(self.shape)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):   # tensor * volume, torch.max(volume) ...: form it
        kwargs = kwargs or {}
        args = [x.materialize() if isinstance(x, LazyVolume) else x for x in args]
        return func(*args, **kwargs)


class FusedCorrBlock(CorrBlockB200):
    """`CorrBlockB200` whose `corr()` is lazy and whose two per-iteration lookups share one launch (module doc)."""

    def __init__(self, fullcorr, num_levels: int = 4, radius: int = 4, pad=(0, 0)):
        self._partner = None     # stereo block: weakref to the mono block built right after it
        self._handoff = None     # mono block: (coords, version, result) left by the stereo block's lookup
        pending = getattr(_tls, "pending", None)
        _tls.pending = None
        if isinstance(fullcorr, LazyVolume):
            lv = fullcorr
            stereo = lv.a.shape[1] != 3
            can_fuse = list(pad) == [0, 0] and lv.shape == [lv.shape[0], lv.a.shape[2], lv.a.shape[3], 1, lv.b.shape[3]]
            t = lv.truncation
            trunc = None if t is None else (t.disp, t.conf, t.gain)
            if can_fuse and stereo and lv.gain == 1.0:
                blk = CorrBlockB200.from_features(lv.a, lv.b, num_levels=num_levels, radius=radius, truncate=trunc)
            elif can_fuse and not stereo and t is None:
                blk = CorrBlockB200.from_normals(lv.a, lv.b, num_levels=num_levels, radius=radius, gain=lv.gain)
            else:
                blk = CorrBlockB200(lv.materialize(), num_levels=num_levels, radius=radius, pad=pad)
            self.__dict__.update(blk.__dict__)
            self._partner, self._handoff = None, None
        else:
            super().__init__(fullcorr, num_levels=num_levels, radius=radius, pad=pad)
        # stereoanywhere.py:253-259 builds the stereo block, then the mono block: pair them
        mate = pending() if pending is not None else None
        if mate is not None and mate._shape == self._shape and mate._partner is None:
            mate._partner = weakref.ref(self)
        else:
            _tls.pending = weakref.ref(self)

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        h = self._handoff
        if h is not None:
            self._handoff = None
            if h[0] is coords and h[1] == coords._version:
                return h[2]
        mate = self._partner() if self._partner is not None else None
        if mate is None or _needs_grad(coords) or not CorrBlockB200._pairable(self, mate):
            return super().__call__(coords)
        s, m = CorrBlockB200.lookup_pair(self, mate, coords)
        mate._handoff = (coords, coords._version, m)
        return s

    @staticmethod
    def corr(fmap2: torch.Tensor, fmap3: torch.Tensor):
        b, c, h, w2 = fmap2.shape
        w3 = fmap3.shape[3]
        lazy_ok = (fmap2.is_cuda and not _needs_grad(fmap2, fmap3) and CorrBlockB200.layout == "packed"
                   and ops.packable(4, 4, w3, [0, 0])
                   and (c == 3 or (CorrBlockB200.precision == "tf32" and ops.corr_packable(c, w2, w3))))
        if not lazy_ok:
            return CorrBlockB200.corr(fmap2, fmap3)
        return LazyVolume(fmap2, fmap3)


_SAVED = "_sa_b200_saved"


def install(sa_mod, fused: bool = False):
    """Point the reference module's `"reg"` correlation block at the B200 implementation (module doc).
    `sa_mod` is `models.stereoanywhere.stereoanywhere` of the reference, imported unchanged."""
    if not hasattr(sa_mod, _SAVED):
        setattr(sa_mod, _SAVED, (sa_mod.CorrBlock1D, sa_mod.truncate_corr_volume_v2))
    ref_block, ref_trunc = getattr(sa_mod, _SAVED)
    if not fused:
        sa_mod.CorrBlock1D = CorrBlockB200
        sa_mod.truncate_corr_volume_v2 = ref_trunc
        return

    def truncate_corr_volume_v2(disp_left, conf_left, conf_th=0.5, attenuation_gain=0.1):
        if conf_th is None and disp_left.is_cuda and not _needs_grad(disp_left, conf_left):
            return LazyTruncation(disp_left, conf_left, attenuation_gain)
        return ref_trunc(disp_left, conf_left, conf_th=conf_th, attenuation_gain=attenuation_gain)

    sa_mod.CorrBlock1D = FusedCorrBlock
    sa_mod.truncate_corr_volume_v2 = truncate_corr_volume_v2


def uninstall(sa_mod):
    """Restore the reference's own block and truncation function."""
    if hasattr(sa_mod, _SAVED):
        sa_mod.CorrBlock1D, sa_mod.truncate_corr_volume_v2 = getattr(sa_mod, _SAVED)
        delattr(sa_mod, _SAVED)
