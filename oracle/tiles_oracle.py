"""CPU oracle for the tile geometry of the reference's MapReduce tiled inference (config 4).

TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle/corr_oracle.py header for the import rule).
Pinned against fixtures generated from /root/reference/mapreduce_v2/tile_wrapper.py by
tests/golden/make_golden.py.  Citations are relative to /root/reference.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.nn.functional as F

Tile = Tuple[int, int, int, int]  # (y0, y1, x0, x1)


def enumerate_tiles(height: int, width: int, tile_h: int, tile_w: int, overlap: int) -> List[Tile]:
    """Tile list of `TileWrapper._enumerate_tiles` (mapreduce_v2/tile_wrapper.py:101-120).

    Stride = tile - overlap; a tile that would cross the border is shifted back so that it ends
    on the border.  The loop runs while the *unclamped* origin is inside the image, so the
    clamped last row/column can be emitted more than once - duplicates are kept, in order.
    """
    tiles: List[Tile] = []
    sy, sx = tile_h - overlap, tile_w - overlap
    for y in range(0, height, sy):
        y1 = min(y + tile_h, height)
        y0 = max(0, y1 - tile_h)
        for x in range(0, width, sx):
            x1 = min(x + tile_w, width)
            x0 = max(0, x1 - tile_w)
            tiles.append((y0, y1, x0, x1))
    return tiles


def blend_weight(h: int, w: int) -> torch.Tensor:
    """clamp(sin(pi y) sin(pi x), 1e-4) on linspace(0,1) (tile_wrapper.py:36-49)."""
    y = torch.linspace(0, 1, h)
    x = torch.linspace(0, 1, w)
    gy, gx = torch.meshgrid(y, x, indexing="ij")
    return torch.clamp(torch.sin(torch.pi * gy) * torch.sin(torch.pi * gx), min=1e-4)


def pad32(h: int, w: int) -> List[int]:
    """[left, right, top, bottom] replicate pad to a multiple of 32 (tile_wrapper.py:226-229)."""
    ph = (((h // 32) + 1) * 32 - h) % 32
    pw = (((w // 32) + 1) * 32 - w) % 32
    return [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]


def tiled_forward(
    model: Callable[..., torch.Tensor],
    left: torch.Tensor,
    right: torch.Tensor,
    mono_l: torch.Tensor,
    mono_r: torch.Tensor,
    tile_h: int,
    tile_w: int,
    overlap: int,
) -> torch.Tensor:
    """Serial tiled inference + cosine-blend stitch (tile_wrapper.py:122-186, 208-247, 328-362).

    `model(l, r, ml, mr)` returns the model's raw output (negative disparity, [1,1,h,w]); the
    wrapper negates it (tile_wrapper.py:206).  Batch must be 1 (tile_wrapper.py:148-149).  No
    global guidance (the default).
    """
    b, _, height, width = left.shape
    assert b == 1
    if height <= tile_h and width <= tile_w:
        return -model(left, right, mono_l, mono_r)
    acc = torch.zeros(b, 1, height, width)
    wsum = torch.zeros_like(acc)
    for (y0, y1, x0, x1) in enumerate_tiles(height, width, tile_h, tile_w, overlap):
        crop = lambda t: t[:, :, y0:y1, x0:x1]
        p = pad32(y1 - y0, x1 - x0)
        args = [F.pad(crop(t), p, mode="replicate") for t in (left, right, mono_l, mono_r)]
        d = -model(*args)
        hd, wd = d.shape[-2:]
        d = d[..., p[2] : hd - p[3], p[0] : wd - p[1]]
        wgt = blend_weight(y1 - y0, x1 - x0)[None, None]
        acc[:, :, y0:y1, x0:x1] += d * wgt
        wsum[:, :, y0:y1, x0:x1] += wgt
    return torch.where(wsum > 0, acc / torch.clamp(wsum, min=1e-4), acc)
