"""Import the UNMODIFIED reference (kei312/stereoanywhere) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists so that `tests/golden/make_golden.py`
can run the real reference code in the build container and freeze its outputs as
fixtures, and so that the GPU tests / `bench.py --impl reference` can run the reference's
own classes on the GPU box.  `/root/reference` does not exist there: the unmodified files
of the path travel in the git-ignored `oracle/_ref/` (recipe: `oracle/make_ref.py`, run by
`__graft_entry__.build()`), and `REFERENCE_ROOT` falls back to it.  The product package
never imports this file.

The reference needs four third-party packages that are absent from this image and
that the hot path never executes:

* ``matplotlib`` (+ ``.pyplot``, ``.cm``, ``.colors``) - imported for visualisation only
  (`models/stereoanywhere/stereoanywhere.py:9`, `utils/utils.py:6-7`);
* ``opt_einsum.contract`` - imported, never called (`update.py:4`);
* ``timm`` - only used by never-instantiated ``Feature*`` classes (`submodule.py:6`);
* ``kornia.filters.spatial_gradient`` - used by ``estimate_normals`` (`utils/utils.py:74`).
  Its stand-in below is a functional re-implementation (replicate pad + central
  difference), *unpinned* by any reference test: parity at the kornia boundary is
  "unpinned"; both sides of every comparison use the same stand-in, and the A2
  parity tests feed unit normals directly.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn.functional as F

_PREBUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _root() -> str:
    env = os.environ.get("SA_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/models/stereoanywhere"):
        return "/root/reference"
    return _PREBUILT


REFERENCE_ROOT = _root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "stereoanywhere"))


def _spatial_gradient(x, mode="diff", order=1, normalized=False):
    """Stand-in for kornia.filters.spatial_gradient(mode='diff', order=1, normalized=False).

    Returns [B, C, 2, H, W]: d/dx then d/dy, 3x3 central difference [-1, 0, 1] on a
    replicate-padded image.
    """
    assert mode == "diff" and order == 1 and not normalized
    b, c, h, w = x.shape
    kx = torch.tensor([[0.0, 0.0, 0.0], [-1.0, 0.0, 1.0], [0.0, 0.0, 0.0]], dtype=x.dtype, device=x.device)
    ker = torch.stack([kx, kx.t()])[:, None]  # [2,1,3,3]
    xp = F.pad(x.reshape(b * c, 1, h, w), (1, 1, 1, 1), mode="replicate")
    return F.conv2d(xp, ker).view(b, c, 2, h, w)


def install_stubs() -> None:
    def _mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    mpl = _mod("matplotlib")
    for sub in ("pyplot", "cm", "colors"):
        setattr(mpl, sub, _mod("matplotlib." + sub))
    _mod("timm")
    oe = _mod("opt_einsum")
    oe.contract = torch.einsum
    kornia = _mod("kornia")
    kf = _mod("kornia.filters")
    kf.spatial_gradient = _spatial_gradient
    kornia.filters = kf


def import_reference():
    """Return the reference's `models.stereoanywhere` package, imported unchanged."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("models.stereoanywhere")


def import_reference_corr():
    """(CorrBlock1D, utils module) of the reference."""
    pkg = import_reference()
    import importlib

    corr = importlib.import_module("models.stereoanywhere.corr")
    utils = importlib.import_module("models.stereoanywhere.utils.utils")
    return corr.CorrBlock1D, utils


def import_reference_tiles():
    """The reference's mapreduce_v2.tile_wrapper module (pure torch, no stubs needed)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("mapreduce_v2.tile_wrapper")
