"""Tile-sharded full-resolution inference (BASELINE config 4) and batch sharding (config 3).

Host-side plumbing only (torch + torch.distributed).  The geometry reproduces the reference's
MapReduce wrapper so that a tile-sharded run returns the reference's tiled result:

* tile enumeration  - mapreduce_v2/tile_wrapper.py:101-120 (`TileWrapper._enumerate_tiles`):
  stride = tile - overlap, border tiles shifted back inside the image; the reference emits the
  clamped last row / column more than once and accumulates the duplicates (harmless after
  normalisation, but they do change the blend where tiles overlap).  `unique=False` keeps that
  behaviour bit for bit; `unique=True` runs every distinct tile ONCE and weights it by its
  multiplicity - the same stitched image up to fp32 summation order, with 10 instead of 12 model
  runs for the `middlebury` preset on 1984x2880.  The multi-GPU path shards the distinct tiles.
* per-tile replicate pad to a multiple of 32 and un-pad - tile_wrapper.py:226-247.
* blend weight clamp(sin(pi y) sin(pi x), 1e-4) and weighted accumulation - tile_wrapper.py:36-49,
  328-362; final `stitched / clamp(weight, 1e-4)` - tile_wrapper.py:185.

Multi-GPU: tiles are independent units.  Every rank runs its share of tiles and accumulates
`disp * w` locally; ONE collective - `reduce(sum)` of that `[H,W]` fp32 tensor to rank 0 - stitches the
image.  The weight accumulator `sum w` of the reference depends on the tile geometry
only, so rank 0 forms it locally (`weight_sum`, same accumulation order as the reference) instead of
shipping a second `[H,W]` plane through the collective.  Batch sharding needs one `all_gather` of the
disparities.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tile = Tuple[int, int, int, int]  # y0, y1, x0, x1

#: the reference's presets that matter for the BASELINE configs (mapreduce_v2/tile_presets.py:37-127):
#: name -> (tile_height, tile_width, overlap)
PRESETS = {
    "default": (448, 448, 96),
    "middlebury": (1120, 672, 112),
    "kitti": (448, 1344, 128),
    "sceneflow": (448, 448, 112),
    "booster": (896, 1120, 224),
}


def _axis_starts(extent: int, tile: int, overlap: int) -> List[Tuple[int, int]]:
    step = tile - overlap
    if step <= 0:
        raise ValueError("overlap must be smaller than the tile")
    spans = []
    pos = 0
    while pos < extent:
        hi = min(pos + tile, extent)
        spans.append((max(0, hi - tile), hi))
        pos += step
    return spans


def enumerate_tiles(height: int, width: int, tile_h: int, tile_w: int, overlap: int, unique: bool = False) -> List[Tile]:
    """Row-major list of (y0, y1, x0, x1); `unique=True` drops repeated tiles (first occurrence kept)."""
    tiles = [(y0, y1, x0, x1) for (y0, y1) in _axis_starts(height, tile_h, overlap)
             for (x0, x1) in _axis_starts(width, tile_w, overlap)]
    return list(dict.fromkeys(tiles)) if unique else tiles


def tile_multiplicity(height: int, width: int, tile_h: int, tile_w: int, overlap: int):
    """Distinct tiles in first-occurrence order with the number of times the reference emits them."""
    counts = {}
    for t in enumerate_tiles(height, width, tile_h, tile_w, overlap):
        counts[t] = counts.get(t, 0) + 1
    return list(counts.items())


def blend_weight(h: int, w: int, device=None) -> torch.Tensor:
    """[h, w] cosine window of the reference (tile_wrapper.py:36-49)."""
    if h <= 0 or w <= 0:
        raise ValueError("Tile dimensions must be positive")
    y = torch.linspace(0, 1, h, device=device)
    x = torch.linspace(0, 1, w, device=device)
    wy = torch.sin(torch.pi * y).unsqueeze(1)
    wx = torch.sin(torch.pi * x).unsqueeze(0)
    return torch.clamp(wy * wx, min=1e-4)


def weight_sum(height: int, width: int, work, device=None) -> torch.Tensor:
    """`[H, W]` sum of the blend windows of all tiles in `work` = [(tile, multiplicity)], accumulated in
    `work` order exactly like the reference's `weight_sum += weight` (tile_wrapper.py:358-362)."""
    den = torch.zeros(height, width, dtype=torch.float32, device=device)
    cache = {}
    for (y0, y1, x0, x1), mult in work:
        key = (y1 - y0, x1 - x0)
        if key not in cache:
            cache[key] = blend_weight(key[0], key[1], device=device)
        for _ in range(mult):
            den[y0:y1, x0:x1] += cache[key]
    return den


def pad_to_32(h: int, w: int) -> List[int]:
    """[left, right, top, bottom] of the replicate pad (tile_wrapper.py:226-229, test.py:204-207)."""
    ph = (32 - h % 32) % 32
    pw = (32 - w % 32) % 32
    return [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]


def shard(items: Sequence, rank: int, world: int) -> List:
    """Round-robin share of `items` for `rank` (tiles differ in cost only at the borders)."""
    return [it for i, it in enumerate(items) if i % world == rank]


def run_tile(model: Callable[..., torch.Tensor], tensors: Sequence[Optional[torch.Tensor]], tile: Tile) -> torch.Tensor:
    """Crop, replicate-pad to /32, run `model`, negate, un-pad (tile_wrapper.py:208-247)."""
    y0, y1, x0, x1 = tile
    pad = pad_to_32(y1 - y0, x1 - x0)
    args = [None if t is None else F.pad(t[:, :, y0:y1, x0:x1], pad, mode="replicate") for t in tensors]
    out = model(*args)
    if isinstance(out, (tuple, list)):
        out = out[0]
    disp = -out
    hd, wd = disp.shape[-2:]
    return disp[..., pad[2]: hd - pad[3], pad[0]: wd - pad[1]]


def tiled_inference(
    model: Callable[..., torch.Tensor],
    left: torch.Tensor,
    right: torch.Tensor,
    mono_left: Optional[torch.Tensor],
    mono_right: Optional[torch.Tensor],
    tile_h: int,
    tile_w: int,
    overlap: int,
    *,
    group=None,
    unique: Optional[bool] = None,
    dst: int = 0,
) -> Optional[torch.Tensor]:
    """Tiled inference, tiles sharded over the ranks of `group` (single process when torch.distributed
    is not initialised).  Returns the stitched `[1,1,H,W]` disparity on rank `dst`, None elsewhere.

    `unique=None` picks the reference's duplicate-keeping enumeration for a single process (bit-exact
    parity with `TileWrapper.forward`) and the de-duplicated one when sharding."""
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    b, _, height, width = left.shape
    if b != 1:
        raise ValueError("tiled inference supports batch size == 1 (tile_wrapper.py:148-149)")
    if height <= tile_h and width <= tile_w:  # single-shot path of the reference (tile_wrapper.py:151-153)
        if rank != dst:
            return None
        out = model(left, right, mono_left, mono_right)
        out = out[0] if isinstance(out, (tuple, list)) else out
        return -out
    if unique is None:
        unique = world > 1
    if unique:
        work = tile_multiplicity(height, width, tile_h, tile_w, overlap)
    else:
        work = [(t, 1) for t in enumerate_tiles(height, width, tile_h, tile_w, overlap)]
    num = torch.zeros(height, width, dtype=torch.float32, device=left.device)  # sum of disp * w over my tiles
    for (y0, y1, x0, x1), mult in shard(work, rank, world):
        disp = run_tile(model, (left, right, mono_left, mono_right), (y0, y1, x0, x1)).to(num.device)
        wgt = blend_weight(y1 - y0, x1 - x0, device=num.device)
        for _ in range(mult):  # the reference accumulates a repeated tile once per repeat
            num[y0:y1, x0:x1] += disp[0, 0].float() * wgt
    if world > 1:
        dist.reduce(num, dst=dst, op=dist.ReduceOp.SUM, group=group)  # the path's only collective
        if rank != dst:
            return None
    den = weight_sum(height, width, work, device=num.device)  # geometry only: no need to communicate it
    out = torch.where(den > 0, num / torch.clamp(den, min=1e-4), num)
    return out.view(1, 1, height, width)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous slice of the batch for `rank` (config 3: 64 pairs -> 8 per GPU)."""
    b = t.shape[0]
    per = math.ceil(b / world)
    return t[rank * per: min(b, (rank + 1) * per)]


def gather_batch(local: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of per-rank `[B/N, ...]` results back into `[B, ...]` (equal shares)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=0)


class SlotStitcher:
    """Stitch of tile-sharded inference WITHOUT a collective (config 4; `csrc/stitch.cu`).

    Every (image, tile) unit owns a slot of `tile_h x tile_w` floats on the gathering rank.  The rank that ran a
    tile launches ONE kernel (`sa_stitch_tile`: un-pad, scale, blend window, multiplicity) whose stores go straight
    into that slot - a peer pointer over NVLink / NVSwitch when the tile ran on another GPU (torch symmetric
    memory).  The gathering rank then runs `sa_stitch_finish` once per step: sum of the covering slots in the
    reference's enumeration order, divided by the weight plane (which depends on the geometry only and is formed
    once, locally).  What an NCCL `reduce(sum)` of the `[images, H, W]` accumulator did in round 1 - every rank
    zeroing, accumulating and shipping the whole 91 MB plane - shrinks to each tile crossing the switch exactly once,
    as soon as it is finished, and the result no longer depends on how the tiles were sharded (fixed summation
    order): N ranks return bit for bit what one rank returns.

    Synchronisation is two signal-pad flags per parity (double-buffered slots): non-gathering ranks `put_signal`
    DONE after their last tile of a step and `wait_signal` FREE before they overwrite a parity two steps later; the
    gathering rank does the mirror image on a side stream, so its own next step computes meanwhile.
    """

    DONE, FREE = 0, 2   # signal channels (+ parity)

    def __init__(self, images: int, height: int, width: int, work, device, group=None, dst: int = 0,
                 timeout_ms: int = 60000, local: bool = False):
        """`local=True`: a single-process stitcher even when torch.distributed is initialised (every tile runs here;
        used to check an N-rank result against the one-rank result)."""
        import torch.distributed as dist

        from . import _lib

        self._lib = _lib.load()
        self.device = torch.device(device)
        distributed = dist.is_available() and dist.is_initialized() and not local
        if local:
            dst = 0
        self.group = group if group is not None else (dist.group.WORLD if distributed else None)
        self.rank = dist.get_rank(self.group) if distributed else 0
        self.world = dist.get_world_size(self.group) if distributed else 1
        self.dst, self.timeout_ms = dst, timeout_ms
        self.images, self.height, self.width = images, height, width
        if width % 4:
            raise ValueError("SlotStitcher needs an image width that is a multiple of 4")
        self.units = [(img, tile, mult) for img in range(images) for (tile, mult) in work]
        self.offsets, total, rows = [], 0, []
        for img, (y0, y1, x0, x1), _mult in self.units:
            if x0 % 4 or (x1 - x0) % 4:
                raise ValueError(f"tile columns must start and end on multiples of 4 (got x0={x0}, x1={x1})")
            self.offsets.append(total)
            rows.append([img, y0, y1, x0, x1, total // 4])
            total += (y1 - y0) * (x1 - x0)
        self.slot_floats = total
        self.hdl = None
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm

            self.slots = symm.empty(2 * total, dtype=torch.float32, device=self.device)
            self.hdl = symm.rendezvous(self.slots, self.group)
            self._slot_base = int(self.hdl.buffer_ptrs[dst])
        else:
            self.slots = torch.empty(2 * total, dtype=torch.float32, device=self.device)
            self._slot_base = self.slots.data_ptr()
        self._weights = {}
        self._step = 0
        self._parity = 0
        self._need_free = [False, False]
        if self.rank == dst:
            self.den = torch.clamp(weight_sum(height, width, work, device=self.device), min=1e-4)
            self.table = torch.tensor(rows, dtype=torch.int32, device=self.device)
            self.out = [torch.empty(images, height, width, dtype=torch.float32, device=self.device) for _ in range(2)]
            self.side = torch.cuda.Stream(device=self.device)
            self.tiles_done = [torch.cuda.Event() for _ in range(2)]
            self.finished = [torch.cuda.Event() for _ in range(2)]
            for e in self.finished:
                e.record(torch.cuda.current_stream(self.device))

    def my_units(self, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
        """Indices into `self.units` of the tiles `rank` runs (round-robin)."""
        rank = self.rank if rank is None else rank
        world = self.world if world is None else world
        return [i for i in range(len(self.units)) if i % world == rank]

    def begin(self) -> None:
        """Start a step: pick the slot parity and wait until the gathering rank has finished reading it."""
        i = self._parity = self._step & 1
        self._step += 1
        if self.rank == self.dst:
            torch.cuda.current_stream(self.device).wait_event(self.finished[i])
        elif self._need_free[i]:
            self.hdl.wait_signal(self.dst, self.FREE + i, self.timeout_ms)
            self._need_free[i] = False

    def add(self, unit: int, src: torch.Tensor, up: int = 1, scale: float = -1.0, pad_top: int = 0, pad_left: int = 0) -> None:
        """Weighted tile of unit `unit` -> its slot on the gathering rank.  `src`: the tile's result `[..., h, w]`
        (contiguous fp32): padded full-resolution model output (`up=1, scale=-1`: the reference negates it,
        tile_wrapper.py:206) or a quarter-resolution disparity (`up=4, scale=4`)."""
        _img, (y0, y1, x0, x1), mult = self.units[unit]
        th, tw = y1 - y0, x1 - x0
        if (th, tw) not in self._weights:
            self._weights[(th, tw)] = blend_weight(th, tw, device=self.device).contiguous()
        if src.dtype != torch.float32 or not src.is_contiguous():
            src = src.float().contiguous()
        slot = self._slot_base + 4 * (self._parity * self.slot_floats + self.offsets[unit])
        rc = self._lib.sa_stitch_tile(src.data_ptr(), src.shape[-2], src.shape[-1], up, float(scale), pad_top, pad_left,
                                      th, tw, self._weights[(th, tw)].data_ptr(), float(mult), slot,
                                      torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("sa_stitch_tile: " + self._lib.sa_last_error().decode())

    def end(self) -> None:
        """All tiles of this rank's share have been added: signal the gathering rank / finish the image there."""
        i = self._parity
        if self.rank != self.dst:
            self.hdl.put_signal(self.dst, self.DONE + i, self.timeout_ms)
            self._need_free[i] = True
            return
        main = torch.cuda.current_stream(self.device)
        self.tiles_done[i].record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.tiles_done[i])
            for r in range(self.world):
                if r != self.dst:
                    self.hdl.wait_signal(r, self.DONE + i, self.timeout_ms)
            rc = self._lib.sa_stitch_finish(self._slot_base + 4 * i * self.slot_floats, self.table.data_ptr(), len(self.units),
                                            self.den.data_ptr(), self.out[i].data_ptr(), self.images, self.height,
                                            self.width, self.side.cuda_stream)
            if rc != 0:
                raise RuntimeError("sa_stitch_finish: " + self._lib.sa_last_error().decode())
            for r in range(self.world):
                if r != self.dst:
                    self.hdl.put_signal(r, self.FREE + i, self.timeout_ms)
            self.finished[i].record(self.side)

    def drain(self) -> Optional[torch.Tensor]:
        """Complete everything outstanding; the stitched `[images, H, W]` of the last step on the gathering rank."""
        if self.rank == self.dst:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
            return self.out[self._parity] if self._step else None
        for i in range(2):
            if self._need_free[i]:
                self.hdl.wait_signal(self.dst, self.FREE + i, self.timeout_ms)
                self._need_free[i] = False
        return None


def tiled_inference_b200(model: Callable[..., torch.Tensor], left: torch.Tensor, right: torch.Tensor,
                         mono_left: Optional[torch.Tensor], mono_right: Optional[torch.Tensor], tile_h: int, tile_w: int,
                         overlap: int, *, stitcher: Optional[SlotStitcher] = None, group=None,
                         dst: int = 0) -> Optional[torch.Tensor]:
    """`tiled_inference` on CUDA with the collective-free stitch (`SlotStitcher`): the distinct tiles of the
    reference's enumeration are sharded over the ranks, each tile's (negated, un-padded, blended) result is stored
    straight into the gathering rank's memory, and that rank returns the stitched `[1,1,H,W]` disparity - the
    reference's `TileWrapper.forward` result (tile_wrapper.py:143-187) up to fp32 summation order of repeated
    tiles.  Pass a `stitcher` to reuse its buffers across calls (same image size and preset)."""
    b, _, height, width = left.shape
    if b != 1:
        raise ValueError("tiled inference supports batch size == 1 (tile_wrapper.py:148-149)")
    if height <= tile_h and width <= tile_w:
        return tiled_inference(model, left, right, mono_left, mono_right, tile_h, tile_w, overlap, group=group, dst=dst)
    if stitcher is None:
        work = tile_multiplicity(height, width, tile_h, tile_w, overlap)
        if width % 4 or any(x0 % 4 or (x1 - x0) % 4 for (_y0, _y1, x0, x1), _m in work):
            # the stitch kernels move float4 columns; odd geometries take the portable stitch (one `reduce`)
            return tiled_inference(model, left, right, mono_left, mono_right, tile_h, tile_w, overlap, group=group,
                                   unique=True, dst=dst)
        stitcher = SlotStitcher(1, height, width, work, left.device, group=group, dst=dst)
    stitcher.begin()
    for u in stitcher.my_units():
        _img, (y0, y1, x0, x1), _m = stitcher.units[u]
        pad = pad_to_32(y1 - y0, x1 - x0)
        args = [None if t is None else F.pad(t[:, :, y0:y1, x0:x1], pad, mode="replicate")
                for t in (left, right, mono_left, mono_right)]
        out = model(*args)
        if isinstance(out, (tuple, list)):
            out = out[0]
        stitcher.add(u, out, up=1, scale=-1.0, pad_top=pad[2], pad_left=pad[0])
    stitcher.end()
    res = stitcher.drain()
    return None if res is None else res.view(1, 1, height, width).clone()
